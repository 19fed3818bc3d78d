#!/bin/bash
# Evidence run for the BASELINE.json configs that are not the bench headline (run under gpurun, 1 GPU).
# Writes gpurun_out/cfg*.json|ssv; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
N100=771819207259660166073186517813   # 100-bit semiprime, bench.random_semiprime(100, 1004)
# config 4: wheel-level tuner sweep (levels 5..9) on a 100-bit semiprime -> qimcifa_calibration.ssv
( cd gpurun_out && printf "$N100\n1\n0\n" | ../qimcifa_b200/bin/qimcifa_tuner --batches 4096 > cfg4_tuner_stdout.txt 2>&1; cp qimcifa_calibration.ssv cfg4_tuner_100bit.ssv )
# config 2 at other sizes (device-resident guesses/s, tuner-selected level)
for B in 100 112 128; do
  python bench.py --bits $B --steps 10 --warmup 3 --no-cpu --no-ttf > gpurun_out/cfg_bits$B.json 2> gpurun_out/cfg_bits$B.err
done
# config 5: random-guess Monte-Carlo mode on a 128-bit semiprime
python bench.py --mode random --bits 128 --steps 8 --warmup 2 > gpurun_out/cfg5_random128.json 2> gpurun_out/cfg5.err
# exact kernel alone (the per-guess test the BASELINE roofline describes)
python bench.py --kernel 1 --level 5 --batches 1024 --steps 5 --warmup 3 --no-cpu --no-ttf > gpurun_out/cfg_exact_kernel.json 2> gpurun_out/cfg_exact.err
for f in gpurun_out/cfg_bits*.json gpurun_out/cfg5_random128.json gpurun_out/cfg_exact_kernel.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print(sys.argv[1], "%.4e" % d["value"], d["config"].get("level"), d["config"].get("batches_per_step"), d.get("roofline", {}).get("frac"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
cat gpurun_out/cfg4_tuner_100bit.ssv; tail -2 gpurun_out/cfg4_tuner_stdout.txt
