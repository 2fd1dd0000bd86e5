python -m pytest tests/test_gpu_parity.py -x -q -k "interp" 2>&1 | tail -2
python tools/opbench.py --ops interp_bwd,interp_bwd_v,interp_bwd_b --env DRTK_B200_LINEAR=1,DRTK_B200_BWD_V4=1
