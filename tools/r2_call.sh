#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2r}
V=drtk_b200/variants
python -m pytest tests -m gpu -x -q -k "edge or pipeline or raster" 2>&1 | tail -3
{
for cfg in 4 3 5; do
  echo "== config $cfg"
  DRTK_B200_LIB=$V/lib_eg_single.so python tools/opbench.py --config $cfg --ops edge_fused --dump /tmp/eg_$cfg.pt | sed "s/^/eg_single /"
  for x in eg_split_a5 eg_split_a4 eg_split_a6 eg_split_a4_b4 eg_split_a4_b2; do
    DRTK_B200_LIB=$V/lib_$x.so python tools/opbench.py --config $cfg --ops edge_fused --cmp /tmp/eg_$cfg.pt | sed "s/^/$x /"
  done
done
echo "== config 4 overdraw"
DRTK_B200_LIB=$V/lib_eg_single.so python tools/opbench.py --config 4 --overdraw --ops edge_fused --dump /tmp/eg_o.pt | sed "s/^/eg_single /"
for x in eg_split_a5 eg_split_a4 eg_split_a6 eg_split_a4_b4 eg_split_a4_b2; do
  DRTK_B200_LIB=$V/lib_$x.so python tools/opbench.py --config 4 --overdraw --ops edge_fused --cmp /tmp/eg_o.pt | sed "s/^/$x /"
done
} > gpurun_out/${T}_opbench.txt 2>&1
cat gpurun_out/${T}_opbench.txt
