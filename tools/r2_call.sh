#!/bin/bash
mkdir -p gpurun_out; T=${TAG:-r2i}
K='regex:^(bin_kernel|scan_kernel|raster_tiles_kernel|render_fwd|interp_fwd|xy_table|edge_grad_strip|vi_table|interp_bwd_quad|tri_table|render_bwd_walk|unpad)'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 13 -c 13 -f -o gpurun_out/${T}_prof_step python tools/profile_step.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log; ls -la gpurun_out/${T}_prof_step.ncu-rep
