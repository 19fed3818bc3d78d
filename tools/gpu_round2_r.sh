#!/bin/bash
# GPU call R of round 2: the biased group offsets and the class-restricted replay of the packed trips -- whole GPU suite, bench
# line, launch shapes, per-step times down slice 7 of 8 at the new step size, ncu captures for profiles/current_kernels.json
set -u
mkdir -p gpurun_out
T=${1:-r2r}
python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
for B in 112 128; do python bench.py --bits $B --steps 6 --warmup 3 --no-cpu --no-ttf > gpurun_out/${T}_bench_bits$B.json 2> gpurun_out/${T}_bench_bits$B.err; done
for S in "7 8" "3 4" "1 2"; do python bench.py --steps 20 --warmup 5 --slice-of $S --no-cpu --no-ttf > gpurun_out/${T}_bench_slice_${S// /of}.json 2> gpurun_out/${T}_bench_slice_${S// /of}.err; done
python tools/slice_profile.py 7 8 448 25 > gpurun_out/${T}_slice7.txt 2>&1
python tools/slice_profile.py 0 8 448 6 > gpurun_out/${T}_slice0.txt 2>&1
python tools/profile_launch.py --kind filter > gpurun_out/${T}_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/${T}_launch_filter112.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/${T}_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --depth 64 > gpurun_out/${T}_launch_deep64.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 8 -f -o gpurun_out/${T}_prof_filter \
    python tools/profile_launch.py --kind filter --reps 2 > gpurun_out/${T}_ncu_filter.log 2>&1 || echo "ncu filter failed"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 8 -f -o gpurun_out/${T}_prof_filter128 \
    python tools/profile_launch.py --kind filter --bits 128 --reps 2 > gpurun_out/${T}_ncu_filter128.log 2>&1 || echo "ncu filter128 failed"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-ttf > gpurun_out/${T}_ncu_bench.log 2>&1 || echo "ncu bench launches failed"
timeout 400 python tools/plan_cost.py > gpurun_out/${T}_plan_cost.jsonl 2> gpurun_out/${T}_plan_cost.err || echo "plan_cost failed"
cat gpurun_out/${T}_launch_*.json | cut -c1-300
cat gpurun_out/${T}_slice7.txt | tail -26
for f in gpurun_out/${T}_bench*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d["config"].get("level"), d["config"].get("batches_per_step"), (d.get("roofline") or {}).get("plan"), d["config"].get("tuner_cost_s"), (d.get("whole_range_sweep") or {}).get("seconds"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
