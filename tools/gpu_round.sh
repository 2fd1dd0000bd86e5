#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list, ncu --set full (+source) of chosen kernels.
# usage: tools/gpu_round.sh [kernel_regex] [tag]
K=${1:-interp_bwd_tile_kernel}; TAG=${2:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_$TAG.txt
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:$K -c ${NCU_COUNT:-8} -f -o gpurun_out/prof_$TAG python tools/profile_step.py > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.txt; cat gpurun_out/bench_$TAG.json | cut -c1-1500
