#!/bin/bash
# Multi-GPU call of round 2 (gpurun --gpus 8): the 2-GPU tests (peer-memory found flag, shared dispenser), then the sweep bench and
# the random-mode bench (BASELINE config 5) at 2, 4 and 8 GPUs, one rank per GPU.
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2m_gpus.txt
python -m pytest tests/test_gpu_peer.py tests/test_gpu_cli.py -m gpu -q -k "peer or two_gpus or dispenser" > gpurun_out/r2m_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r2m_pytest_2gpu.log
P=29500
for N in 2 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 \
      > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err; echo "bench N=$N rc=$?"
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --mode random --steps 6 --warmup 3 \
      > gpurun_out/r2m_bench_random_n$N.json 2> gpurun_out/r2m_bench_random_n$N.err; echo "bench random N=$N rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, d["n_gpus"], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d["config"].get("kernel_ms_per_step_by_rank"), [(t.get("bits"), t.get("split"), t.get("seconds"), t.get("correct")) for t in d.get("time_to_factor", []) if t.get("bits", 0) > 100], d.get("self_checks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
