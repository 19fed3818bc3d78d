#!/bin/bash
# Multi-GPU call at the end of round 2 (gpurun --gpus 8): the driver's command at 8 GPUs on the final kernels.
set -u
mkdir -p gpurun_out
N=${1:-8}
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2g_bench_n*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["n_gpus"], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "ms/step %.2f" % d["ms_per_step"], d["config"].get("kernel_ms_per_step_by_rank"), [(t.get("bits"), t.get("split", "")[:12], round(t.get("seconds"), 4), t.get("correct")) for t in d.get("time_to_factor", []) if t.get("bits", 0) > 100], (d.get("multi_gpu_checks") or {}).get("peer_flag_stops_running_sweeps", {}).get("ok"), (d.get("whole_range_sweep") or {}).get("seconds"), (d.get("whole_range_sweep") or {}).get("seconds_by_rank"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/r2g_bench_n$N.err
