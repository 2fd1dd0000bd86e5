#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): reducer check, then bench.py under torchrun.
mkdir -p gpurun_out; T=${TAG:-r2m}; NG=${NG:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/dist_check.py > gpurun_out/${T}_distcheck.txt 2>&1
tail -3 gpurun_out/${T}_distcheck.txt
timeout 400 $RUN --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 --transport multimem > gpurun_out/${T}_bench_c4_multimem.json 2> gpurun_out/${T}_bench_c4_multimem.err
cut -c1-300 gpurun_out/${T}_bench_c4_multimem.json; tail -2 gpurun_out/${T}_bench_c4_multimem.err
timeout 400 $RUN --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 5 --transport nccl > gpurun_out/${T}_bench_c4_nccl.json 2> gpurun_out/${T}_bench_c4_nccl.err
cut -c1-300 gpurun_out/${T}_bench_c4_nccl.json
${EXTRA}
