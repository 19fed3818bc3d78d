"""One 960-batch step with a forced filter plan (diagnostic for ncu): python tools/one_plan.py D,TP,J [DIV] [REPS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qimcifa_b200 import abi

force = sys.argv[1]
div = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
p, q, n = bench.synthetic_semiprime(96, 1002)
batches = 960
abi.set_option("plan_force", force)
with abi.Context(0) as ctx:
    _, _, r, bc = abi.node_split(n, 1)
    b_hi = bc // div - (3 * batches if div == 1 else 0)
    for _ in range(reps):
        res = ctx.sweep(n, 5, b_hi - batches, b_hi, batches_per_launch=batches)
        print(force, div, "%.2f ms" % (res.device_seconds * 1e3), "%.3e g/s" % (res.guesses_tested / res.device_seconds), res.kernel_launches, flush=True)
