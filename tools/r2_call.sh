#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2w}
python -m pytest tests -m gpu -x -q -k "rasterize" 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_raster.txt
