"""Per-step device time down one node slice (diagnostic): python tools/slice_profile.py NODE COUNT [BATCHES] [STEPS]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qimcifa_b200 import abi
import bench

node, count = int(sys.argv[1]), int(sys.argv[2])
batches = int(sys.argv[3]) if len(sys.argv) > 3 else 960
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 23
p, q, n = bench.synthetic_semiprime(96, 1002)
with abi.Context(0) as ctx:
    _, _, r, bc = abi.node_split(n, count)
    lo, hi = abi.node_batches(bc, r, node)
    ctx.sweep(n, 5, hi - batches, hi, batches_per_launch=batches)
    for k in range(steps):
        b_hi = hi - k * batches
        if b_hi - batches < lo:
            break
        res = ctx.sweep(n, 5, b_hi - batches, b_hi, batches_per_launch=batches)
        print(k, b_hi, "%.2f ms" % (res.device_seconds * 1e3), "%.3e g/s" % (res.guesses_tested / res.device_seconds), res.kernel_launches, flush=True)
