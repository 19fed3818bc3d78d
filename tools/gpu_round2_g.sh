#!/bin/bash
# GPU call G of round 2: ncu captures of the packed kernels (degree 1 at 128 bit, degree 2 at 96 bit) with source-level stall samples
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 1 -f -o gpurun_out/r2g_prof_d1_128 \
    python tools/profile_launch.py --kind filter --bits 128 --reps 1 > gpurun_out/r2g_ncu_d1.log 2>&1 || echo "ncu d1 failed"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 1 -f -o gpurun_out/r2g_prof_d2p_96 \
    python tools/profile_launch.py --kind filter --bits 96 --reps 1 > gpurun_out/r2g_ncu_d2p.log 2>&1 || echo "ncu d2p failed"
tail -2 gpurun_out/r2g_ncu_d1.log gpurun_out/r2g_ncu_d2p.log
