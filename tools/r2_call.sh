#!/bin/bash
# One GPU visit of round 2 (content is whatever the repo holds when the call is accepted).
mkdir -p gpurun_out; T=${TAG:-r2g}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${T}_pytest.txt
tail -4 gpurun_out/${T}_pytest.txt
DRTK_B200_DISPATCH=ctypes timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fp64.py -q -k "pipeline or autograd or autocast or drop_in or graph or corner or empty_face or f64" 2>&1 | tail -30 > gpurun_out/${T}_pytest_ctypes.txt
tail -3 gpurun_out/${T}_pytest_ctypes.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt
for c in "--config 2" "--config 3"; do
  n=$(echo $c | tr -d ' -'); timeout 300 python bench.py $c --no-cpu-baseline > gpurun_out/${T}_bench_$n.json 2> gpurun_out/${T}_bench_$n.err; cut -c1-330 gpurun_out/${T}_bench_$n.json
done
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
cut -c1-300 gpurun_out/${T}_bench_reference.json
