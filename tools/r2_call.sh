#!/bin/bash
# One GPU visit of round 2 (content is whatever the repo holds when the call is accepted).
mkdir -p gpurun_out; T=${TAG:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.txt
timeout 240 python tools/raster_sanity.py > gpurun_out/${T}_sanity.txt 2>&1; rc=$?
if [ $rc -ne 0 ]; then export DRTK_B200_RASTER_V1=1; echo "SANITY rc=$rc -> falling back to DRTK_B200_RASTER_V1" >> gpurun_out/${T}_sanity.txt; fi
tail -25 gpurun_out/${T}_sanity.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.txt
tail -5 gpurun_out/${T}_pytest.txt
DRTK_B200_DISPATCH=torch timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "pipeline or autograd or autocast or drop_in or graph or corner or empty_face" 2>&1 | tail -15 > gpurun_out/${T}_pytest_torchops.txt
tail -5 gpurun_out/${T}_pytest_torchops.txt
timeout 200 python tools/opbench.py --ops rasterize,interp_bwd,interp_bwd_b,interp_bwd_v,edge_fused,render_bwd --dump /tmp/a.pt > gpurun_out/${T}_opbench.txt 2>&1
DRTK_B200_MERGED=1 timeout 200 python tools/opbench.py --ops interp_bwd,interp_bwd_b --cmp /tmp/a.pt >> gpurun_out/${T}_opbench.txt 2>&1
DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --ops rasterize --cmp /tmp/a.pt >> gpurun_out/${T}_opbench.txt 2>&1
timeout 200 python tools/opbench.py --ops rasterize,render_fwd,interp_fwd,interp_bwd,render_bwd,edge_fused --overdraw --dump /tmp/b.pt >> gpurun_out/${T}_opbench.txt 2>&1
DRTK_B200_MERGED=1 timeout 200 python tools/opbench.py --ops interp_bwd --overdraw --cmp /tmp/b.pt >> gpurun_out/${T}_opbench.txt 2>&1
DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --ops rasterize --overdraw --cmp /tmp/b.pt >> gpurun_out/${T}_opbench.txt 2>&1
timeout 200 python tools/opbench.py --config 5 --ops rasterize --dump /tmp/c.pt >> gpurun_out/${T}_opbench.txt 2>&1
DRTK_B200_RASTER_V1=1 timeout 200 python tools/opbench.py --config 5 --ops rasterize --cmp /tmp/c.pt >> gpurun_out/${T}_opbench.txt 2>&1
cat gpurun_out/${T}_opbench.txt
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-700 gpurun_out/${T}_bench.json
