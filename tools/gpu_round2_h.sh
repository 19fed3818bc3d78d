#!/bin/bash
# GPU call H of round 2: packed trips with warp-uniform replays: launch shapes, planner calibration, parity campaign
set -u
mkdir -p gpurun_out
python tools/profile_launch.py --kind filter > gpurun_out/r2h_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --option plan_packed=0 > gpurun_out/r2h_launch_filter_unpacked.json 2>&1
python tools/profile_launch.py --kind filter --force 1,0,0 > gpurun_out/r2h_launch_filter_d1.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/r2h_launch_filter112.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2h_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 --level 6 > gpurun_out/r2h_launch_filter128_l6.json 2>&1
cat gpurun_out/r2h_launch_*.json | cut -c1-420
python tools/plan_cost.py > gpurun_out/r2h_plan_cost.jsonl 2> gpurun_out/r2h_plan_cost.err; echo "plan_cost rc=$?"
python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
