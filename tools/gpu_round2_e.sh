#!/bin/bash
# GPU call E of round 2: packed 16-bit trips at degree 1, survivors tested on the spot when the queue is full, chain-draw
# applicability: whole GPU suite, bench lines, launch shapes, planner calibration.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -6 gpurun_out/r2e_pytest.log
for B in 96 112 128; do python bench.py --bits $B --steps 6 --warmup 3 --no-cpu --no-ttf > gpurun_out/r2e_bench_bits$B.json 2> gpurun_out/r2e_bench_bits$B.err; done
python tools/profile_launch.py --kind filter > gpurun_out/r2e_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --force 1,0,0 > gpurun_out/r2e_launch_filter_d1.json 2>&1
python tools/profile_launch.py --kind filter --force 1,0,8 > gpurun_out/r2e_launch_filter_d1j8.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/r2e_launch_filter112.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2e_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 --level 6 > gpurun_out/r2e_launch_filter128_l6.json 2>&1
python tools/profile_launch.py --kind randomc --bits 128 > gpurun_out/r2e_launch_randomc.json 2>&1
python tools/plan_cost.py > gpurun_out/r2e_plan_cost.jsonl 2> gpurun_out/r2e_plan_cost.err; echo "plan_cost rc=$?"
cat gpurun_out/r2e_launch_*.json | cut -c1-420
for f in gpurun_out/r2e_bench*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d["config"].get("level"), d["config"].get("batches_per_step"), (d.get("roofline") or {}).get("plan"), d["config"].get("tuner_cost_s"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
