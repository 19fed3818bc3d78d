#!/bin/bash
# GPU call F of round 2: why the packed degree-1 trips are slow -- instruction semantics, replay statistics, packed degree 2
set -u
mkdir -p gpurun_out
./tools/ubench/pipes 2>&1 | head -3 > gpurun_out/r2f_semantics.jsonl; cat gpurun_out/r2f_semantics.jsonl
python tools/profile_launch.py --kind filter --bits 128 --option debug_stats=1 --reps 1 > gpurun_out/r2f_stats128.json 2> gpurun_out/r2f_stats128.err; grep "filter stats" gpurun_out/r2f_stats128.err | head -4
python tools/profile_launch.py --kind filter --bits 112 --option debug_stats=1 --reps 1 > gpurun_out/r2f_stats112.json 2> gpurun_out/r2f_stats112.err; grep "filter stats" gpurun_out/r2f_stats112.err | head -4
python tools/profile_launch.py --kind filter --bits 96 --option debug_stats=1 --reps 1 > gpurun_out/r2f_stats96.json 2> gpurun_out/r2f_stats96.err; grep "filter stats" gpurun_out/r2f_stats96.err | head -4
python tools/profile_launch.py --kind filter > gpurun_out/r2f_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --option plan_packed=0 > gpurun_out/r2f_launch_filter_unpacked.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2f_launch_filter128.json 2>&1
cat gpurun_out/r2f_launch_*.json gpurun_out/r2f_stats*.json | cut -c1-420
python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
