#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2x}
python bench.py --config 3 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; cut -c1-300 gpurun_out/${T}_bench_c3.json
python bench.py --config 2 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; cut -c1-300 gpurun_out/${T}_bench_c2.json
python bench.py --overdraw > gpurun_out/${T}_bench_c4o.json 2> gpurun_out/${T}_bench_c4o.err; cut -c1-300 gpurun_out/${T}_bench_c4o.json
python bench.py --config 5 --steps 5 --no-cpu-baseline > gpurun_out/${T}_bench_c5.json 2> gpurun_out/${T}_bench_c5.err; cut -c1-300 gpurun_out/${T}_bench_c5.json; tail -2 gpurun_out/${T}_bench_c5.err
