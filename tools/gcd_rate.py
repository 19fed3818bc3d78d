"""Throughput of the common-factor mode (QCF_MODE_GCD) near sqrt(N): python tools/gcd_rate.py [BITS] [BATCHES]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qimcifa_b200 import abi

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 96
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 64
p, q, n = bench.synthetic_semiprime(bits, 1002)
with abi.Context(0) as ctx:
    _, _, r, bc = abi.node_split(n, 1)
    hi = bc - 8
    ctx.sweep(n, 5, hi - batches, hi, batches_per_launch=batches, flags=abi.MODE_GCD)
    for _ in range(3):
        res = ctx.sweep(n, 5, hi - batches, hi, batches_per_launch=batches, flags=abi.MODE_GCD)
        print(bits, batches, "found", res.found, "%.2f ms" % (res.device_seconds * 1e3), "%.3e guesses/s" % (res.guesses_tested / res.device_seconds), res.kernel_launches, flush=True)
