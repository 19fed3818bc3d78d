#!/bin/bash
# GPU call L of round 2: common-factor mode with two accumulators and 1024-guess batches
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_fuzz.py -m gpu -q -k "gcd or level_10 or common" > gpurun_out/r2l_pytest_gcd.log 2>&1; tail -2 gpurun_out/r2l_pytest_gcd.log
python tools/profile_launch.py --kind gcd > gpurun_out/r2l_launch_gcd.json 2>&1
python tools/profile_launch.py --kind gcd --bits 128 > gpurun_out/r2l_launch_gcd128.json 2>&1
cat gpurun_out/r2l_launch_gcd*.json | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_gcd_chain_kernel -c 1 -f -o gpurun_out/r2l_prof_gcd \
    python tools/profile_launch.py --kind gcd --reps 1 > gpurun_out/r2l_ncu_gcd.log 2>&1 || echo "ncu gcd failed"
