#!/bin/bash
# GPU call K of round 2: common-factor mode with one carry chain per Montgomery row; bench line with the pinned roofline
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -k "gcd or level_10 or common" > gpurun_out/r2k_pytest_gcd.log 2>&1; tail -2 gpurun_out/r2k_pytest_gcd.log
python tools/profile_launch.py --kind gcd > gpurun_out/r2k_launch_gcd.json 2>&1
python tools/profile_launch.py --kind gcd --bits 128 > gpurun_out/r2k_launch_gcd128.json 2>&1
cat gpurun_out/r2k_launch_gcd*.json | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python bench.py --mode random --steps 8 --warmup 3 > gpurun_out/r2k_bench_random.json 2> gpurun_out/r2k_bench_random.err; echo "bench random rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err; echo "bench reference rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2k_bench.json", "gpurun_out/r2k_bench_random.json", "gpurun_out/r2k_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        r = d.get("roofline") or {}
        print(f, "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "roofline", r.get("bound"), r.get("frac"), r.get("issue_slot_frac"), (r.get("profile") or {}).get("status"), d.get("common_factor_mode"))
    except Exception as e:
        print(f, "FAILED", e)
PY
