#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/raster_sanity.py 2>&1 | tail -13 | tee gpurun_out/${T}_raster_sanity.txt
python tools/raster_sweep.py 2>&1 | tee gpurun_out/${T}_raster_sweep.txt
python tools/raster_sweep.py --H 4096 --N 2 2>&1 | tee -a gpurun_out/${T}_raster_sweep.txt
python tools/raster_sweep.py --H 8192 --N 1 --iters 4 2>&1 | tee -a gpurun_out/${T}_raster_sweep.txt
for c in 4 3 5; do python tools/opbench.py --config $c --ops rasterize 2>&1 | tee -a gpurun_out/${T}_opbench.txt; done
timeout 600 compute-sanitizer --tool memcheck python tools/raster_sanity.py --small > gpurun_out/${T}_sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python tools/raster_sanity.py --small > gpurun_out/${T}_sanitizer_racecheck.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_racecheck.txt
