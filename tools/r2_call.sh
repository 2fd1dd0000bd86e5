#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2l}
V=drtk_b200/variants
{
for cfg in 4 5 3; do
  echo "== config $cfg"
  DRTK_B200_LIB=$V/lib_rb_base.so python tools/opbench.py --config $cfg --ops render_bwd --dump /tmp/rb_$cfg.pt | sed "s/^/rb_base /"
  DRTK_B200_LIB=$V/lib_rb_wmerge.so python tools/opbench.py --config $cfg --ops render_bwd --cmp /tmp/rb_$cfg.pt | sed "s/^/rb_wmerge /"
  DRTK_B200_LIB=$V/lib_rf_pipe0.so python tools/opbench.py --config $cfg --ops render_fwd --dump /tmp/rf_$cfg.pt | sed "s/^/rf_pipe0 /"
  DRTK_B200_LIB=$V/lib_rf_pipe1.so python tools/opbench.py --config $cfg --ops render_fwd --cmp /tmp/rf_$cfg.pt | sed "s/^/rf_pipe1 /"
  DRTK_B200_LIB=$V/lib_eg_b1.so python tools/opbench.py --config $cfg --ops edge_fused --dump /tmp/eg_$cfg.pt | sed "s/^/eg_b1 /"
  for x in eg_b4 eg_b8; do DRTK_B200_LIB=$V/lib_$x.so python tools/opbench.py --config $cfg --ops edge_fused --cmp /tmp/eg_$cfg.pt | sed "s/^/$x /"; done
done
echo "== config 4 overdraw"
DRTK_B200_LIB=$V/lib_eg_b1.so python tools/opbench.py --config 4 --overdraw --ops edge_fused --dump /tmp/eg_o.pt | sed "s/^/eg_b1 /"
for x in eg_b4 eg_b8; do DRTK_B200_LIB=$V/lib_$x.so python tools/opbench.py --config 4 --overdraw --ops edge_fused --cmp /tmp/eg_o.pt | sed "s/^/$x /"; done
DRTK_B200_LIB=$V/lib_rb_base.so python tools/opbench.py --config 4 --overdraw --ops render_bwd --dump /tmp/rb_o.pt | sed "s/^/rb_base /"
DRTK_B200_LIB=$V/lib_rb_wmerge.so python tools/opbench.py --config 4 --overdraw --ops render_bwd --cmp /tmp/rb_o.pt | sed "s/^/rb_wmerge /"
} > gpurun_out/${T}_opbench.txt 2>&1
cat gpurun_out/${T}_opbench.txt
