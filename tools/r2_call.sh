#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2m}
V=drtk_b200/variants
{
for cfg in 4 3 5; do
  echo "== config $cfg"
  DRTK_B200_LIB=$V/lib_wl_c2.so python tools/opbench.py --config $cfg --ops wireframe,wireframe_ref --iters 5 | sed "s/^/wl_c2 /"
  DRTK_B200_LIB=$V/lib_wl_c3.so python tools/opbench.py --config $cfg --ops wireframe --iters 5 | sed "s/^/wl_c3 /"
  DRTK_B200_LIB=$V/lib_eg_b8.so python tools/opbench.py --config $cfg --ops edge_fused --dump /tmp/eg_$cfg.pt | sed "s/^/eg_b8 /"
  DRTK_B200_LIB=$V/lib_eg_b16.so python tools/opbench.py --config $cfg --ops edge_fused --cmp /tmp/eg_$cfg.pt | sed "s/^/eg_b16 /"
done
echo "== config 4 overdraw"
DRTK_B200_LIB=$V/lib_eg_b8.so python tools/opbench.py --config 4 --overdraw --ops edge_fused --dump /tmp/eg_o.pt | sed "s/^/eg_b8 /"
DRTK_B200_LIB=$V/lib_eg_b16.so python tools/opbench.py --config 4 --overdraw --ops edge_fused --cmp /tmp/eg_o.pt | sed "s/^/eg_b16 /"
} > gpurun_out/${T}_opbench.txt 2>&1
cat gpurun_out/${T}_opbench.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
