#!/bin/bash
# GPU call Q of round 2: 64-guess trips at degree 1
set -u
mkdir -p gpurun_out
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2q_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/r2q_launch_filter112.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 --level 6 > gpurun_out/r2q_launch_filter128_l6.json 2>&1
python tools/profile_launch.py --kind filter --force 1,0,0 > gpurun_out/r2q_launch_filter96_d1.json 2>&1
python tools/profile_launch.py --kind randomc --bits 128 > gpurun_out/r2q_launch_randomc.json 2>&1
cat gpurun_out/r2q_launch_*.json | cut -c1-330
python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -q -x -k "112 or 128 or degree1 or forced or chain_draw or campaign" > gpurun_out/r2q_pytest.log 2>&1; tail -2 gpurun_out/r2q_pytest.log
