#!/bin/bash
# Second multi-GPU call of round 2 (gpurun --gpus 8): sweep bench at 4 and 8 GPUs after the self-check fix (its test number's
# factor sat at the top of rank 1's slice at nodeCount >= 4), random-mode bench with 16-block steps at 1, 2, 4, 8 GPUs.
set -u
mkdir -p gpurun_out
P=29600
for N in 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 \
      > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err; echo "bench N=$N rc=$?"
done
python bench.py --mode random --steps 6 --warmup 3 > gpurun_out/r2n_bench_random_n1.json 2> gpurun_out/r2n_bench_random_n1.err; echo "bench random N=1 rc=$?"
for N in 2 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --mode random --steps 6 --warmup 3 \
      > gpurun_out/r2n_bench_random_n$N.json 2> gpurun_out/r2n_bench_random_n$N.err; echo "bench random N=$N rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["n_gpus"], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "ms/step %.2f" % d["ms_per_step"], d["config"].get("kernel_ms_per_step_by_rank"), [(t.get("bits"), t.get("split", "")[:12], round(t.get("seconds"), 4), t.get("correct")) for t in d.get("time_to_factor", []) if t.get("bits", 0) > 100], (d.get("multi_gpu_checks") or {}).get("peer_flag_stops_running_sweeps", {}).get("ok"))
    except Exception as e:
        print(f, "FAILED", e)
PY
