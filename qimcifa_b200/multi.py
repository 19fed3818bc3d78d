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


def connect_peers(ctx, dist, rank, world, device):
    """Exchanges the CUDA-IPC handles of every rank's found-flag word (one all_gather of 64 bytes per rank) and maps
    the peers' words, so that a sweep which finds a divisor stops the other GPUs by writing their flags over NVLink
    peer memory -- no collective on the path afterwards.  Returns False (and leaves the NCCL exchange in place) if
    peer mapping is not available."""
    import torch

    mine = torch.tensor(list(ctx.peer_export()), dtype=torch.uint8, device=device)
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    handles = [bytes(t.cpu().tolist()) for t in allh]
    ok = 1
    try:
        ctx.peer_connect(handles, rank)
    except abi.QcfError:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not int(flag.item()):
        ctx.peer_disconnect()
        return False
    return True


def sweep_slice_with_peer_flag(sweep_fn, lo, hi, chunk, max_seconds=None):
    """Peer-flag variant of sweep_slice_until_found: no collective per run.  The rank that finds a divisor has already
    raised every peer's flag inside qcf_sweep; a rank whose flag was raised sees res.aborted and stops.  The caller
    agrees on the result with ONE collective afterwards.  -> dict"""
    t0 = time.perf_counter()
    guesses = launches = rounds = 0
    factor = None
    stopped = False
    while hi > lo:
        b_lo = max(lo, hi - chunk)
        res = sweep_fn(b_lo, hi)
        guesses += res.guesses_tested
        launches += res.kernel_launches
        rounds += 1
        if res.found:
            factor = res.factor
            break
        if res.aborted:
            stopped = True
            break
        hi = b_lo
        if (max_seconds is not None) and ((time.perf_counter() - t0) > max_seconds):
            break
    return {"seconds": time.perf_counter() - t0, "guesses": guesses, "factor": factor, "found": factor is not None,
            "stopped_by_peer": stopped, "launches": launches, "rounds": rounds, "next_batch_hi": hi}


def sweep_shared_dispenser(sweep_fn, dispense_fn, lo, hi, chunk, max_seconds=None):
    """The reference's THREAD model across GPUs: every rank pulls the next run of `chunk` batches of the SAME slice
    [lo, hi) from one shared descending counter (dispense_fn(chunk) -> batches handed out before the call; peer-memory
    fetch-add, abi.Context.peer_dispense) until a divisor is found (the finder raises every peer's flag inside
    qcf_sweep), the slice is exhausted, or its own flag was raised.  No collective on the path.  -> dict"""
    t0 = time.perf_counter()
    guesses = launches = rounds = 0
    factor = None
    stopped = False
    while True:
        start = dispense_fn(chunk)
        b_hi = hi - start
        if b_hi <= lo:
            break
        b_lo = max(lo, b_hi - chunk)
        res = sweep_fn(b_lo, b_hi)
        guesses += res.guesses_tested
        launches += res.kernel_launches
        rounds += 1
        if res.found:
            factor = res.factor
            break
        if res.aborted:
            stopped = True
            break
        if (max_seconds is not None) and ((time.perf_counter() - t0) > max_seconds):
            break
    return {"seconds": time.perf_counter() - t0, "guesses": guesses, "factor": factor, "found": factor is not None,
            "stopped_by_peer": stopped, "launches": launches, "rounds": rounds}


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


# ---- self-checks of the cross-GPU pieces (world >= 2): run by tests/test_gpu_peer.py and by `bench.py --gpus N`, so that the
# peer-memory flag and the shared dispenser have a pass/fail wherever a multi-GPU box runs the bench ----------------------
def _gather_objects(dist, obj, world):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def check_peer_flag(ctx, dist, rank, world, n, level=5, signal_after=0.5):
    """Rank 0 raises every peer's found flag (system-scope atomics into the peers' HBM, qcf_peer_signal) `signal_after`
    seconds into their long sweeps; every other rank's running kernel must stop at its next poll -- seconds, not the
    30 s the sweep would take -- and the flag must be clearable afterwards.  -> dict with "ok" (identical on all ranks)."""
    lo, hi = rank_slice(n, world, rank)
    ctx.set_found(0)
    ctx.sweep(n, level, hi - 64, hi, batches_per_launch=64)  # warm-up: tables, first launch
    dist.barrier()
    mine = {"rank": rank}
    if rank == 0:
        time.sleep(signal_after)
        ctx.peer_signal()
        mine.update(role="signalled")
    else:
        t0 = time.perf_counter()
        r = sweep_slice_with_peer_flag(lambda a, b: ctx.sweep(n, level, a, b, batches_per_launch=4096), lo, hi - 64, 4096,
                                       max_seconds=30.0)
        dt = time.perf_counter() - t0
        polled = ctx.poll_found()
        ctx.set_found(0)
        res = ctx.sweep(n, level, hi - 64, hi, batches_per_launch=64)  # the flag is per context and can be cleared
        mine.update(role="swept", stopped_by_peer=bool(r["stopped_by_peer"]), seconds=dt, rounds=r["rounds"], polled=polled,
                    cleared=(res.aborted == 0 and res.batches_done == 64))
    dist.barrier()
    allr = _gather_objects(dist, mine, world)
    ok = all(x["stopped_by_peer"] and x["polled"] == 1 and x["cleared"] and (0.6 * signal_after < x["seconds"] < signal_after + 4.5)
             for x in allr if x["role"] == "swept")
    return {"ok": bool(ok), "ranks": allr}


def check_shared_dispenser(ctx, dist, rank, world, n, p, level=5, chunk=128):
    """All ranks pull runs of `chunk` batches of ONE node's slice from a counter in rank 0's memory (system-scope
    fetch-adds, over NVLink for the peers): every run is handed out exactly once, the rank whose run holds the factor
    `p` finds it and stops the others through their flags.  -> dict with "ok" (identical on all ranks)."""
    lo, hi = rank_slice(n, 1, 0)
    ctx.set_found(0)
    ctx.sweep(n, level, lo, lo + 8, batches_per_launch=8)  # warm-up at the bottom (no divisor there)
    if rank == 0:
        ctx.dispenser_reset(0)
    dist.barrier()
    r = sweep_shared_dispenser(lambda a, b: ctx.sweep(n, level, a, b, batches_per_launch=chunk),
                               lambda take: ctx.peer_dispense(0, take), lo, hi, chunk, max_seconds=60.0)
    dist.barrier()
    handed_out = ctx.peer_dispense(0, 0)
    ctx.set_found(0)
    mine = {"rank": rank, "found": bool(r["found"]), "right_factor": (r["factor"] == p) if r["found"] else None,
            "rounds": r["rounds"], "stopped_by_peer": bool(r["stopped_by_peer"]), "handed_out": handed_out,
            "seconds": r["seconds"]}
    dist.barrier()
    allr = _gather_objects(dist, mine, world)
    finders = [x for x in allr if x["found"]]
    rounds = sum(x["rounds"] for x in allr)
    ok = (len(finders) >= 1 and all(x["right_factor"] for x in finders)
          and all(x["stopped_by_peer"] for x in allr if not x["found"])  # the others were stopped, they did not run out of work
          and all(x["handed_out"] == chunk * rounds for x in allr))       # every pull counted exactly once
    return {"ok": bool(ok), "rounds": rounds, "ranks": allr}
