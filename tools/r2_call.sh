#!/bin/bash
# One GPU visit of round 2 (content is whatever the repo holds when the call is accepted).
mkdir -p gpurun_out; T=${TAG:-r2h}
timeout 240 python tools/raster_sanity.py > gpurun_out/${T}_sanity.txt 2>&1; rc=$?
if [ $rc -ne 0 ]; then export DRTK_B200_RASTER_V1=1; echo "SANITY rc=$rc -> falling back to DRTK_B200_RASTER_V1" >> gpurun_out/${T}_sanity.txt; fi
tail -3 gpurun_out/${T}_sanity.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_pytest.txt
tail -3 gpurun_out/${T}_pytest.txt
O=gpurun_out/${T}_opbench.txt; : > $O
for cfg in 4 3 5; do
  echo "config $cfg" >> $O
  timeout 300 python tools/opbench.py --config $cfg --ops rasterize,rasterize_ref,wireframe,wireframe_ref --dump /tmp/c.pt >> $O 2>&1
  DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --config $cfg --ops rasterize --cmp /tmp/c.pt >> $O 2>&1
done
cat $O | cut -c1-200
timeout 500 compute-sanitizer --tool memcheck python tools/raster_sanity.py --small > gpurun_out/${T}_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/${T}_sanitizer_memcheck.txt
timeout 500 compute-sanitizer --tool racecheck python tools/raster_sanity.py --small > gpurun_out/${T}_sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/${T}_sanitizer_racecheck.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"drtk" -s 13 -c 13 -f -o gpurun_out/${T}_prof_step python tools/profile_step.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench.json
