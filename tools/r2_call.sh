#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2v}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-420 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --regions 1 --no-extras --no-cpu-baseline --no-ref-cuda > gpurun_out/${T}_ncu_bench.log 2>&1; wc -l gpurun_out/${T}_launches.csv
