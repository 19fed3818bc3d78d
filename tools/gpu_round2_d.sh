#!/bin/bash
# GPU call D of round 2: survivors that find the queue full are tested on the spot (no host fallback): whole GPU suite (no -x),
# pipe micro-benchmarks for the packed 16-bit trip, bench lines at 96/112/128 bit.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -6 gpurun_out/r2d_pytest.log
./tools/ubench/pipes > gpurun_out/r2d_pipes.jsonl 2>&1; cat gpurun_out/r2d_pipes.jsonl
for B in 96 112 128; do python bench.py --bits $B --steps 6 --warmup 3 --no-cpu --no-ttf > gpurun_out/r2d_bench_bits$B.json 2> gpurun_out/r2d_bench_bits$B.err; done
for f in gpurun_out/r2d_bench*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d["config"].get("level"), d["config"].get("batches_per_step"), (d.get("roofline") or {}).get("plan"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
