#!/bin/bash
# One GPU visit of round 2 (content is whatever the repo holds when the call is accepted).
mkdir -p gpurun_out; T=${TAG:-r2f}
timeout 240 python tools/raster_sanity.py > gpurun_out/${T}_sanity.txt 2>&1; rc=$?
if [ $rc -ne 0 ]; then export DRTK_B200_RASTER_V1=1; echo "SANITY rc=$rc -> falling back to DRTK_B200_RASTER_V1" >> gpurun_out/${T}_sanity.txt; fi
tail -2 gpurun_out/${T}_sanity.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_pytest.txt
tail -3 gpurun_out/${T}_pytest.txt
O=gpurun_out/${T}_opbench.txt; : > $O
for cfg in 4 3 5; do
  echo "config $cfg" >> $O
  timeout 200 python tools/opbench.py --config $cfg --ops rasterize,edge_fused,render_bwd --dump /tmp/c.pt >> $O 2>&1
  DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --config $cfg --ops rasterize --cmp /tmp/c.pt >> $O 2>&1
  DRTK_B200_LIB=$PWD/drtk_b200/variants/lib_nopf.so timeout 200 python tools/opbench.py --config $cfg --ops edge_fused --cmp /tmp/c.pt >> $O 2>&1
done
echo "config 4 overdraw" >> $O
timeout 200 python tools/opbench.py --ops rasterize,edge_fused,render_bwd --overdraw --dump /tmp/b.pt >> $O 2>&1
DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --ops rasterize --overdraw --cmp /tmp/b.pt >> $O 2>&1
DRTK_B200_LIB=$PWD/drtk_b200/variants/lib_nopf.so timeout 200 python tools/opbench.py --ops edge_fused --overdraw --cmp /tmp/b.pt >> $O 2>&1
cat $O | sed 's/DRTK_B200_LIB=.*variants\/lib_/V:/' | cut -c1-200
for c in "--config 2" "--config 3" "--config 4 --overdraw"; do
  n=$(echo $c | tr -d ' -'); timeout 300 python bench.py $c --no-cpu-baseline > gpurun_out/${T}_bench_$n.json 2> gpurun_out/${T}_bench_$n.err; cut -c1-330 gpurun_out/${T}_bench_$n.json
done
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench.json
