#!/bin/bash
# GPU call O of round 2: the packed degree-2 offsets' running adds forced onto the multiplier pipe (FilterParams::one): launch shapes
set -u
mkdir -p gpurun_out
python tools/profile_launch.py --kind filter > gpurun_out/r2o_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --depth 64 > gpurun_out/r2o_launch_deep64.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2o_launch_filter128.json 2>&1
python tools/profile_launch.py --kind randomc --bits 128 > gpurun_out/r2o_launch_randomc.json 2>&1
cat gpurun_out/r2o_launch_*.json | cut -c1-330
python -m pytest tests/test_gpu_bench_shape.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; tail -2 gpurun_out/r2o_pytest.log
