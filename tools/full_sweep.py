"""Whole-range sweep of a 96-bit number with no divisor below sqrt(N) (a prime): the true full-job time on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qimcifa_b200 import abi
import bench

n = bench._prev_prime((1 << 96) - 12345)
with abi.Context(0) as ctx:
    _, _, r, bc = abi.node_split(n, 1)
    ctx.sweep(n, 5, bc - 2048, bc, batches_per_launch=2048)
    t0 = time.perf_counter()
    tot = 0
    hi = bc
    while hi > 0:
        lo = max(0, hi - 4096)
        res = ctx.sweep(n, 5, lo, hi, batches_per_launch=4096)
        assert not res.found
        tot += res.guesses_tested
        hi = lo
    dt = time.perf_counter() - t0
    print({"n": str(n), "batches": bc, "guesses": tot, "seconds": dt, "guesses_per_s": tot / dt})
