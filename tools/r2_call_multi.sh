#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): reducer check, then bench.py under torchrun with the three exchange settings.
mkdir -p gpurun_out; T=${TAG:-r2n}; NG=${NG:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/dist_check.py > gpurun_out/${T}_distcheck.txt 2>&1
tail -4 gpurun_out/${T}_distcheck.txt
Q="--no-extras --no-cpu-baseline --no-ref-cuda"
timeout 300 $RUN --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 --transport multimem $Q > gpurun_out/${T}_bench_c4_multimem_bg.json 2> gpurun_out/${T}_bench_c4_multimem_bg.err
cut -c1-260 gpurun_out/${T}_bench_c4_multimem_bg.json; tail -2 gpurun_out/${T}_bench_c4_multimem_bg.err
timeout 300 $RUN --master-port 29513 bench.py --gpus $NG --steps 20 --warmup 5 --transport multimem --bg-ctas 0 $Q > gpurun_out/${T}_bench_c4_multimem_full.json 2> gpurun_out/${T}_bench_c4_multimem_full.err
cut -c1-260 gpurun_out/${T}_bench_c4_multimem_full.json
timeout 300 $RUN --master-port 29514 bench.py --gpus $NG --steps 20 --warmup 5 --transport multimem --bg-ctas 16 $Q > gpurun_out/${T}_bench_c4_multimem_bg16.json 2> gpurun_out/${T}_bench_c4_multimem_bg16.err
cut -c1-260 gpurun_out/${T}_bench_c4_multimem_bg16.json
timeout 300 $RUN --master-port 29515 bench.py --gpus $NG --steps 20 --warmup 5 --transport nccl $Q > gpurun_out/${T}_bench_c4_nccl.json 2> gpurun_out/${T}_bench_c4_nccl.err
cut -c1-260 gpurun_out/${T}_bench_c4_nccl.json
${EXTRA}
