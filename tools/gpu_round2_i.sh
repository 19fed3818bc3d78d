#!/bin/bash
# GPU call I of round 2: two-stage replay of flagged packed trips: launch shapes, planner calibration
set -u
mkdir -p gpurun_out
python tools/profile_launch.py --kind filter > gpurun_out/r2i_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --option plan_packed=0 > gpurun_out/r2i_launch_filter_unpacked.json 2>&1
python tools/profile_launch.py --kind filter --force 1,0,0 > gpurun_out/r2i_launch_filter_d1.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/r2i_launch_filter112.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2i_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 --level 6 > gpurun_out/r2i_launch_filter128_l6.json 2>&1
cat gpurun_out/r2i_launch_*.json | cut -c1-330
python tools/plan_cost.py > gpurun_out/r2i_plan_cost.jsonl 2> gpurun_out/r2i_plan_cost.err; echo "plan_cost rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 1 -f -o gpurun_out/r2i_prof_d2p_96 \
    python tools/profile_launch.py --kind filter --bits 96 --reps 1 > gpurun_out/r2i_ncu_d2p.log 2>&1 || echo "ncu d2p failed"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 1 -f -o gpurun_out/r2i_prof_d1_128 \
    python tools/profile_launch.py --kind filter --bits 128 --reps 1 > gpurun_out/r2i_ncu_d1.log 2>&1 || echo "ncu d1 failed"
