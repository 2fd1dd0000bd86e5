#!/bin/bash
# Build A/B variants of libdrtk_b200.so that differ in compile-time knobs of ONE translation unit.
# usage: tools/build_variants.sh <unit.cu> name1:"-DFOO=1 -DBAR=2" name2:"..."
# Result: drtk_b200/variants/lib_<name>.so (select with DRTK_B200_LIB=<path>).  Not part of the product build.
set -e
cd "$(dirname "$0")/../drtk_b200/csrc"
UNIT=$1; shift
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 $ARCH -lineinfo --use_fast_math -Xcompiler -fPIC"
mkdir -p ../variants
make -s -j8 > /dev/null
OTHERS=$(ls *.o | grep -v "^${UNIT%.cu}.o$")
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  $NVCC $FLAGS $defs -ccbin /usr/bin/g++ -c $UNIT -o ../variants/${UNIT%.cu}_$name.o
  $NVCC -shared $ARCH -ccbin /usr/bin/g++ -o ../variants/lib_$name.so ../variants/${UNIT%.cu}_$name.o $OTHERS -lcudart_static -lrt -lpthread -ldl
  rm ../variants/${UNIT%.cu}_$name.o
  echo "built variants/lib_$name.so ($defs)"
done
