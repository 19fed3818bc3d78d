"""CPU-only, world_size 2 over gloo: the N>1 host logic (rank = node slice, found-flag exchange, reductions)."""
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, factor_batch, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qimcifa_b200 import multi

    lo, hi = multi.rank_slice(n, world, rank)
    calls = []

    def fake_sweep(b_lo, b_hi):  # stands in for Context.sweep: "finds" the divisor iff its batch is in the run
        calls.append((b_lo, b_hi))
        hit = b_lo <= factor_batch < b_hi
        return types.SimpleNamespace(found=hit, factor=12345 if hit else None, guesses_tested=(b_hi - b_lo) * 10, kernel_launches=1)

    flag = torch.zeros(1, dtype=torch.int32)
    r = multi.sweep_slice_until_found(fake_sweep, lo, hi, 4, dist=dist, flag=flag)
    mx, sm = multi.reduce_max_sum(dist, [r["seconds"], float(rank)], [r["guesses"], 1.0], "cpu")
    out_q.put((rank, lo, hi, calls, r["found"], r["factor"], r["rounds"], mx[1], sm))
    dist.barrier()
    dist.destroy_process_group()


def _run(n, factor_batch, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, factor_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_rank_slices_and_found_flag():
    from qimcifa_b200 import abi

    n = (2**40 - 87) * (2**40 - 195)  # 80-bit: 2 x 2^40/2310*480/BW ~ 1024 batches
    _, _, node_range, bc = abi.node_split(n, 2)
    assert node_range >= 16
    # divisor in rank 1's (bottom) slice, 9 batches below its top: rank 1 finds it in its 3rd run of 4
    fb = bc - node_range - 10
    (r0, lo0, hi0, calls0, found0, fac0, rounds0, mxrank, sm), (r1, lo1, hi1, calls1, found1, fac1, rounds1, _, _) = _run(n, fb, 29631)
    assert (lo0, hi0) == (bc - node_range, bc) and (lo1, hi1) == (0, bc - node_range)  # rank 0 = top slice
    assert found0 and found1 and fac0 is None and fac1 == 12345  # both learn it, only the owner has the factor
    assert rounds0 == rounds1 == 3  # both stop after the same round
    assert calls0 == [(hi0 - 4, hi0), (hi0 - 8, hi0 - 4), (hi0 - 12, hi0 - 8)]  # descending runs
    assert calls1[-1][0] <= fb < calls1[-1][1]
    assert mxrank == 1.0 and sm == [240.0, 2.0]  # max over ranks, sum over ranks


def test_two_ranks_exhaust_without_a_hit():
    from qimcifa_b200 import abi

    n = 1000003 * 1000033  # one batch per node
    _, _, node_range, bc = abi.node_split(n, 2)
    res = _run(n, 10**9, 29633)
    for rank, lo, hi, calls, found, fac, rounds, _, _ in res:
        assert not found and fac is None and calls == [(lo, hi)] and rounds == 1


def test_shared_dispenser_hands_every_run_out_once():
    """Host logic of the shared dispenser (qimcifa_b200/multi.py): several workers pull runs from one counter; the runs
    are disjoint, descend from the top of the slice, and the worker whose run holds the target stops the rest."""
    import threading
    import types

    from qimcifa_b200 import multi

    lo, hi, chunk, target = 100, 5000, 128, 2222
    lock = threading.Lock()
    state = {"handed": 0, "stop": False}
    runs = []

    def dispense(take):
        with lock:
            start = state["handed"]
            state["handed"] += take
            return start

    def sweep(b_lo, b_hi):
        with lock:
            runs.append((b_lo, b_hi))
            if state["stop"]:
                return types.SimpleNamespace(found=0, aborted=1, factor=None, guesses_tested=0, kernel_launches=0)
        found = b_lo <= target < b_hi
        if found:
            with lock:
                state["stop"] = True
        return types.SimpleNamespace(found=int(found), aborted=0, factor=target if found else None,
                                     guesses_tested=b_hi - b_lo, kernel_launches=1)

    results = []
    threads = [threading.Thread(target=lambda: results.append(multi.sweep_shared_dispenser(sweep, dispense, lo, hi, chunk)))
               for _ in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert sum(r["found"] for r in results) == 1 and [r["factor"] for r in results if r["found"]] == [target]
    runs.sort(reverse=True)
    assert runs[0][1] == hi and all(a[0] == b[1] for a, b in zip(runs, runs[1:]))  # contiguous, descending, no overlap
    assert all(b - a == chunk for a, b in runs) and runs[-1][0] <= target
    # exhaustion: nothing to find -> every batch of the slice is handed out exactly once, the last run is clipped at lo
    state.update(handed=0, stop=False)
    runs.clear()
    target = -1
    r = multi.sweep_shared_dispenser(sweep, dispense, lo, hi, chunk)
    assert not r["found"] and sum(b - a for a, b in runs) == hi - lo and min(a for a, _ in runs) == lo
