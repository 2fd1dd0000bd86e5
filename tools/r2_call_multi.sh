#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): reducer check, then bench.py under torchrun for each transport.
mkdir -p gpurun_out; T=${TAG:-r2m}; NG=${NG:-2}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/dist_check.py > gpurun_out/${T}_distcheck.txt 2>&1
tail -6 gpurun_out/${T}_distcheck.txt
for tr in nccl multimem; do
  timeout 400 $RUN --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 --transport $tr ${BENCH_ARGS} > gpurun_out/${T}_bench_${tr}.json 2> gpurun_out/${T}_bench_${tr}.err
  cut -c1-300 gpurun_out/${T}_bench_${tr}.json; tail -2 gpurun_out/${T}_bench_${tr}.err
done
