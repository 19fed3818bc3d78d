#!/bin/bash
# Random-guess mode (BASELINE config 5) at N GPUs on the final kernels.
set -u
mkdir -p gpurun_out
N=${1:-8}
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $N --mode random --steps 6 --warmup 3 \
    > gpurun_out/r2g_bench_random_n$N.json 2> gpurun_out/r2g_bench_random_n$N.err; echo "bench random N=$N rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2g_bench_random_n$N.json").read().strip().split("\n")[-1])
print("random", d["n_gpus"], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "ms/step %.2f" % d["ms_per_step"])
PY
