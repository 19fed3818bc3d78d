#!/bin/bash
# GPU call A of round 2: the whole GPU suite at HEAD (incl. the bench-shape campaign), one bench line, and ncu captures
# of the exact chain kernel's two-stage loop (never profiled so far) and of the filter kernel at the bench shape.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
for K in exact filter; do
  python tools/profile_launch.py --kind $K > gpurun_out/r2a_launch_$K.json 2> gpurun_out/r2a_launch_$K.err || echo "launch $K failed"
  REGEX=$([ $K = exact ] && echo qcf_exact_chain_kernel || echo qcf_filter_kernel)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 1 -c 1 -f -o gpurun_out/r2a_prof_$K \
      python tools/profile_launch.py --kind $K --reps 2 > gpurun_out/r2a_ncu_$K.log 2>&1 || echo "ncu $K failed"
done
python tools/profile_launch.py --kind filter --depth 64 > gpurun_out/r2a_launch_deep64.json 2>&1
python tools/profile_launch.py --kind exact1 > gpurun_out/r2a_launch_exact1.json 2>&1
python tools/profile_launch.py --kind gcd > gpurun_out/r2a_launch_gcd.json 2>&1
python tools/profile_launch.py --kind random --bits 128 > gpurun_out/r2a_launch_random.json 2>&1
cat gpurun_out/r2a_launch_*.json | cut -c1-400
