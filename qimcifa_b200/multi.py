"""Multi-GPU host logic: one process per GPU, rank k = node k of the reference's nodeCount/nodeId split
(reference src/qimcifa.cpp:155-158, intent semantics: node k owns the k-th slice from the top).

The path shards with ZERO data-path communication; the only exchange is the found flag (the reference's
finish(), include/qimcifa.hpp:102-105, which a human relays between nodes): one 1-word all_reduce(MAX) after
every run of batches.  `dist` is torch.distributed (NCCL on GPUs, gloo in the CPU tests) or None.
"""
import time

from . import abi


def rank_slice(n, world, rank):
    """-> (batch_lo, batch_hi): the batches rank `rank` of `world` owns, to be swept DESCENDING."""
    _, _, node_range, batch_count = abi.node_split(n, world)
    return abi.node_batches(batch_count, node_range, rank)


def exchange_found(found, dist=None, flag=None):
    """All ranks learn whether ANY rank found a divisor in the run just finished."""
    if dist is None:
        return bool(found)
    flag[0] = 1 if found else 0
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    return bool(int(flag.item()))


def sweep_slice_until_found(sweep_fn, lo, hi, chunk, dist=None, flag=None, max_seconds=None):
    """Sweeps batches [lo, hi) from the top in runs of `chunk` until some rank reports a divisor.

    sweep_fn(b_lo, b_hi) -> object with .found, .factor, .guesses_tested, .kernel_launches
    Every rank executes the same number of rounds (a rank whose slice is exhausted keeps taking part in
    the flag exchange), so the collective never deadlocks.  -> dict
    """
    t0 = time.perf_counter()
    guesses = launches = rounds = 0
    factor = None
    stop = False
    while not stop:
        found = False
        if hi > lo:
            b_lo = max(lo, hi - chunk)
            res = sweep_fn(b_lo, hi)
            guesses += res.guesses_tested
            launches += res.kernel_launches
            hi = b_lo
            if res.found:
                found, factor = True, res.factor
        rounds += 1
        found_any = exchange_found(found, dist, flag)
        # stop when a divisor was found anywhere, when EVERY rank is exhausted (MAX of "still has work" is 0), or on timeout
        all_exhausted = not exchange_found(hi > lo, dist, flag)
        timed_out = (max_seconds is not None) and exchange_found((time.perf_counter() - t0) > max_seconds, dist, flag)
        stop = found_any or all_exhausted or timed_out
    return {"seconds": time.perf_counter() - t0, "guesses": guesses, "factor": factor, "found": found_any,
            "launches": launches, "rounds": rounds, "next_batch_hi": hi}


def reduce_max_sum(dist, maxima, sums, device):
    """Device-side reduction of timings (max over ranks) and counters (sum over ranks)."""
    import torch

    mx = torch.tensor(list(maxima), dtype=torch.float64, device=device)
    sm = torch.tensor(list(sums), dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return [float(x) for x in mx.tolist()], [float(x) for x in sm.tolist()]
