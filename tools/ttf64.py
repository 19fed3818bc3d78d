"""Time-to-factor of one 64-bit semiprime through the C ABI (diagnostic): python tools/ttf64.py [LEVEL] [REPS]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qimcifa_b200 import abi

level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p, q, n = bench.synthetic_semiprime(64, 1001)
with abi.Context(0) as ctx:
    _, _, r, bc = abi.node_split(n, 1)
    ctx.build_tables(level)
    for _ in range(reps):
        t0 = time.perf_counter()
        res = ctx.sweep(n, level, 0, bc, batches_per_launch=bc)
        dt = time.perf_counter() - t0
        print(level, bc, res.found, res.factor, "host %.3f ms" % (dt * 1e3), "device %.3f ms" % (res.device_seconds * 1e3), res.kernel_launches, res.guesses_tested, flush=True)
