#!/bin/bash
# One GPU visit of round 2 (content is whatever the repo holds when the call is accepted).
mkdir -p gpurun_out; T=${TAG:-r2d}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt
timeout 240 python tools/raster_sanity.py > gpurun_out/${T}_sanity.txt 2>&1; rc=$?
if [ $rc -ne 0 ]; then export DRTK_B200_RASTER_V1=1; echo "SANITY rc=$rc -> falling back to DRTK_B200_RASTER_V1" >> gpurun_out/${T}_sanity.txt; fi
tail -22 gpurun_out/${T}_sanity.txt
echo "--- the same scenes with the round-1 kernel (timing reference)" >> gpurun_out/${T}_sanity.txt
DRTK_B200_RASTER_V1=1 timeout 240 python tools/raster_sanity.py >> gpurun_out/${T}_sanity.txt 2>&1
tail -8 gpurun_out/${T}_sanity.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_pytest.txt
tail -6 gpurun_out/${T}_pytest.txt
DRTK_B200_DISPATCH=torch timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "pipeline or autograd or autocast or drop_in or graph or corner or empty_face" 2>&1 | tail -30 > gpurun_out/${T}_pytest_torchops.txt
tail -4 gpurun_out/${T}_pytest_torchops.txt
O=gpurun_out/${T}_opbench.txt; : > $O
for cfg in 4 3 5; do
  echo "config $cfg" >> $O
  timeout 200 python tools/opbench.py --config $cfg --ops rasterize,edge_fused --dump /tmp/c.pt >> $O 2>&1
  DRTK_B200_LIB=$PWD/drtk_b200/variants/lib_p104.so timeout 200 python tools/opbench.py --config $cfg --ops rasterize --cmp /tmp/c.pt >> $O 2>&1
  DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --config $cfg --ops rasterize --cmp /tmp/c.pt >> $O 2>&1
  DRTK_B200_LIB=$PWD/drtk_b200/variants/lib_edgeold.so timeout 200 python tools/opbench.py --config $cfg --ops edge_fused --cmp /tmp/c.pt >> $O 2>&1
done
echo "config 4 overdraw" >> $O
timeout 200 python tools/opbench.py --ops rasterize,edge_fused --overdraw --dump /tmp/b.pt >> $O 2>&1
DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --ops rasterize --overdraw --cmp /tmp/b.pt >> $O 2>&1
DRTK_B200_LIB=$PWD/drtk_b200/variants/lib_edgeold.so timeout 200 python tools/opbench.py --ops edge_fused --overdraw --cmp /tmp/b.pt >> $O 2>&1
cat $O | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"raster_tiles|edge_grad_strip" -s 2 -c 2 -f -o gpurun_out/${T}_prof python tools/profile_step.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench.json
