python -m pytest tests/test_gpu_parity.py -x -q -k "interp" 2>&1 | tail -2
python tools/opbench.py --ops interp_bwd,interp_bwd_v,interp_bwd_b --env DRTK_B200_BWD_V4=1
ncu --set full --clock-control none --import-source on -k regex:interp_bwd_quad -c 2 -f -o gpurun_out/prof_q5 python tools/opbench.py --ops interp_bwd --iters 1 > gpurun_out/ncu_q5.log 2>&1
