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


class _StoreCtx:
    """Stands in for abi.Context in the cross-GPU self-checks (multi.check_*): flags and the dispenser word live in the
    process group's key-value store instead of peer-mapped HBM; a sweep "finds" the divisor iff its batch is in the run."""

    def __init__(self, rank, world, store, factor_batch):
        self.rank, self.world, self.store, self.fb = rank, world, store, factor_batch

    def set_found(self, v=1):
        self.store.set("flag%d" % self.rank, str(int(v)))

    def poll_found(self):
        return int(self.store.get("flag%d" % self.rank))

    def peer_signal(self):
        for r in range(self.world):
            if r != self.rank:
                self.store.set("flag%d" % r, "1")

    def dispenser_reset(self, v=0):
        self.store.set("disp", str(int(v)))

    def peer_dispense(self, owner, take):
        return self.store.add("disp", take) - take

    def sweep(self, n, level, lo, hi, batches_per_launch=0, flags=0):
        import time

        if self.poll_found():
            return types.SimpleNamespace(found=0, aborted=1, factor=None, batches_done=0, guesses_tested=0, kernel_launches=0)
        time.sleep(0.005)
        hit = lo <= self.fb < hi
        if hit:
            self.peer_signal()  # what qcf_sweep does on a connected context
        return types.SimpleNamespace(found=int(hit), aborted=0, factor=777 if hit else None, batches_done=hi - lo,
                                     guesses_tested=(hi - lo) * 10, kernel_launches=1)


def _check_worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qimcifa_b200 import abi, multi

    store = dist.distributed_c10d._get_default_store()
    n = (2**56 - 5) * (2**56 - 3)  # 112-bit: slices of millions of batches
    _, _, _, bc = abi.node_split(n, 1)
    ctx = _StoreCtx(rank, world, store, factor_batch=10**18)
    ctx.set_found(0)
    dist.barrier()
    flag = multi.check_peer_flag(ctx, dist, rank, world, n, signal_after=0.4)
    ctx.fb = bc - 700  # ~6 runs of 128 below the top
    disp = multi.check_shared_dispenser(ctx, dist, rank, world, n, 777, chunk=128)
    out_q.put((rank, flag, disp))
    dist.barrier()
    dist.destroy_process_group()


def test_cross_gpu_self_checks_over_gloo():
    """multi.check_peer_flag / check_shared_dispenser (what `bench.py --gpus N` asserts on a multi-GPU box and what
    tests/test_gpu_peer.py runs on 2 GPUs), world size 2 over gloo with a store-backed stand-in for the context."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_check_worker, args=(r, 2, 29637, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, flag0, disp0), (_, flag1, disp1) = res
    assert flag0 == flag1 and flag0["ok"], flag0
    swept = [x for x in flag0["ranks"] if x["role"] == "swept"]
    assert len(swept) == 1 and swept[0]["stopped_by_peer"] and swept[0]["cleared"]
    assert disp0 == disp1 and disp0["ok"], disp0
    assert sum(1 for x in disp0["ranks"] if x["found"]) == 1 and 3 <= disp0["rounds"] <= 16


def test_flag_check_number_keeps_its_factor_in_slice_0_at_every_node_count():
    """bench.py's cross-GPU found-flag self-check sweeps the slices of ranks 1.. of a number whose factor must not be met:
    the smaller factor lies inside rank 0's slice (rank 0 only signals) at nodeCount 2, 4 and 8, well away from its bottom.
    (Round 2's first version planted it at 75 % of sqrt(N) -- the top of rank 1's slice at nodeCount 4 -- and the bench failed
    at 4 and 8 GPUs; that version had only ever run on 2.)"""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle_c as oc
    from qimcifa_b200 import abi, multi

    n, p = bench.flag_check_number()
    assert n % p == 0 and n.bit_length() == 112 and p < n // p
    pb = (oc.backward(p) - 1 - 2) // abi.BIGGEST_WHEEL  # the batch that holds p's wheel-11 index
    for world in (2, 4, 8):
        lo0, hi0 = multi.rank_slice(n, world, 0)
        assert lo0 + (hi0 - lo0) // 8 < pb < hi0, (world, lo0, pb, hi0)
        for rank in range(1, world):
            lo, hi = multi.rank_slice(n, world, rank)
            assert not (lo <= pb < hi)


def test_bench_steps_stay_in_the_divisor_free_top_half_of_every_slice():
    """bench.py's step size is the same at every N and its warm-up + timed steps, walking down from the top of each rank's slice,
    (a) never reach the batch that holds the planted factor and (b) stay inside the top half -- the first octave of guess
    values -- of every slice of a 1, 2, 4 and 8-way split (the bottom octaves are timed by whole_range_sweep instead)."""
    sys.path.insert(0, ROOT)
    import bench
    from qimcifa_b200 import abi

    for bits in (96, 112, 128):
        p, q, n = bench.synthetic_semiprime(bits, 1002)
        pb = (bench._index_of(p) - 2) // abi.BIGGEST_WHEEL
        sizes = {w: bench.fixed_batches_per_step(abi, n, p, 20, 5, w) for w in (1, 2, 4, 8)}
        assert len(set(sizes.values())) == 1, sizes  # the same step at every N: a 1 -> 8 comparison measures scaling only
        B = sizes[1]
        assert B >= 64 and B % 64 == 0 and B <= 4096
        for world in (1, 2, 4, 8):
            _, _, node_range, batch_count = abi.node_split(n, world)
            for rank in range(world):
                lo, hi = abi.node_batches(batch_count, node_range, rank)
                bottom = hi - 25 * B  # the lowest batch the 25 steps touch
                assert bottom >= lo + (hi - lo) // 2, (bits, world, rank)
                assert not (bottom <= pb < hi), (bits, world, rank, pb)
