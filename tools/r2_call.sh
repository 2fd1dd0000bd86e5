#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2y}
python -m pytest tests -m gpu -x -q -k "rasterize or pipeline or smoke" 2>&1 | tail -2
python tools/raster_sanity.py 2>&1 | tail -4
python tools/raster_sweep.py 2>&1 | tee gpurun_out/${T}_raster_sweep.txt
python tools/opbench.py --config 4 --ops rasterize 2>&1 | tee gpurun_out/${T}_opbench.txt
python tools/opbench.py --config 3 --ops rasterize 2>&1 | tee -a gpurun_out/${T}_opbench.txt
python bench.py --no-extras --no-cpu-baseline --regions 3 2>/dev/null | cut -c1-330
