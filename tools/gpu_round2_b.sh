#!/bin/bash
# GPU call B of round 2: drift lines + per-chain offset correction in the filter kernel, IMAD.WIDE / linear reciprocal in
# the exact chain kernel: GPU suite, bench, calibration of the planner's cost model, per-step times down slice 7 of 8, ncu.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python tools/plan_cost.py > gpurun_out/r2b_plan_cost.jsonl 2> gpurun_out/r2b_plan_cost.err; echo "plan_cost rc=$?"
python tools/slice_profile.py 7 8 896 25 > gpurun_out/r2b_slice7.txt 2>&1
python tools/slice_profile.py 0 8 896 6 > gpurun_out/r2b_slice0.txt 2>&1
for K in exact exact1 gcd; do python tools/profile_launch.py --kind $K > gpurun_out/r2b_launch_$K.json 2>&1; done
python tools/profile_launch.py --kind random --bits 128 > gpurun_out/r2b_launch_random.json 2>&1
python tools/profile_launch.py --kind filter > gpurun_out/r2b_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/r2b_launch_filter112.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2b_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 --level 7 > gpurun_out/r2b_launch_filter128_l7.json 2>&1
python tools/profile_launch.py --kind filter --depth 64 > gpurun_out/r2b_launch_deep64.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 8 -f -o gpurun_out/r2b_prof_filter \
    python tools/profile_launch.py --kind filter --reps 2 > gpurun_out/r2b_ncu_filter.log 2>&1 || echo "ncu filter failed"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_exact_chain_kernel -s 1 -c 1 -f -o gpurun_out/r2b_prof_exact \
    python tools/profile_launch.py --kind exact --reps 2 > gpurun_out/r2b_ncu_exact.log 2>&1 || echo "ncu exact failed"
cat gpurun_out/r2b_launch_*.json | cut -c1-330
