#!/bin/bash
# GPU call P of round 2: degree-1 trip with a single min chain and the step counter's add on the multiplier pipe
set -u
mkdir -p gpurun_out
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2p_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/r2p_launch_filter112.json 2>&1
cat gpurun_out/r2p_launch_*.json | cut -c1-330
