#!/bin/bash
# GPU call J of round 2: the validated state -- whole GPU suite, bench lines (sweep at 96/112/128 bit, random mode), launch
# shapes, per-step times down slice 7 of 8, ncu captures for profiles/current_kernels.json
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python bench.py --mode random --steps 8 --warmup 3 > gpurun_out/r2j_bench_random.json 2> gpurun_out/r2j_bench_random.err; echo "bench random rc=$?"
for B in 112 128; do python bench.py --bits $B --steps 6 --warmup 3 --no-cpu --no-ttf > gpurun_out/r2j_bench_bits$B.json 2> gpurun_out/r2j_bench_bits$B.err; done
python tools/slice_profile.py 7 8 896 25 > gpurun_out/r2j_slice7.txt 2>&1
python tools/slice_profile.py 0 8 896 6 > gpurun_out/r2j_slice0.txt 2>&1
for K in exact exact1 gcd; do python tools/profile_launch.py --kind $K > gpurun_out/r2j_launch_$K.json 2>&1; done
python tools/profile_launch.py --kind randomc --bits 128 > gpurun_out/r2j_launch_randomc.json 2>&1
python tools/profile_launch.py --kind filter > gpurun_out/r2j_launch_filter.json 2>&1
python tools/profile_launch.py --kind filter --bits 112 > gpurun_out/r2j_launch_filter112.json 2>&1
python tools/profile_launch.py --kind filter --bits 128 > gpurun_out/r2j_launch_filter128.json 2>&1
python tools/profile_launch.py --kind filter --depth 64 > gpurun_out/r2j_launch_deep64.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 8 -f -o gpurun_out/r2j_prof_filter \
    python tools/profile_launch.py --kind filter --reps 2 > gpurun_out/r2j_ncu_filter.log 2>&1 || echo "ncu filter failed"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 8 -f -o gpurun_out/r2j_prof_filter128 \
    python tools/profile_launch.py --kind filter --bits 128 --reps 2 > gpurun_out/r2j_ncu_filter128.log 2>&1 || echo "ncu filter128 failed"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_filter_kernel -c 8 -f -o gpurun_out/r2j_prof_deep64 \
    python tools/profile_launch.py --kind filter --depth 64 --reps 2 > gpurun_out/r2j_ncu_deep64.log 2>&1 || echo "ncu deep64 failed"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qcf_exact_chain_kernel -s 1 -c 1 -f -o gpurun_out/r2j_prof_exact \
    python tools/profile_launch.py --kind exact --reps 2 > gpurun_out/r2j_ncu_exact.log 2>&1 || echo "ncu exact failed"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-ttf > gpurun_out/r2j_ncu_bench.log 2>&1 || echo "ncu bench launches failed"
cat gpurun_out/r2j_launch_*.json | cut -c1-300
for f in gpurun_out/r2j_bench*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d["config"].get("level"), d["config"].get("batches_per_step"), (d.get("roofline") or {}).get("plan"), d["config"].get("tuner_cost_s"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
