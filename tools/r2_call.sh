#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2ab}
timeout 70 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(algorithms_agree and deep_300) or size_classes_agree" > gpurun_out/${T}_sanitizer_racecheck_deep.txt 2>&1; tail -3 gpurun_out/${T}_sanitizer_racecheck_deep.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${T}_pytest_gpu.txt
python tools/opbench.py --config 4 --ops rasterize --iters 10 2>&1 | tail -1 | tee gpurun_out/${T}_opbench.txt
