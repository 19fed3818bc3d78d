"""GPU (needs 2): the found flag over peer memory.  Two processes, one GPU each; rank 0 raises rank 1's flag through
the CUDA-IPC-mapped word while rank 1 is inside a long sweep; rank 1's kernel stops at its next poll."""
import os
import sys
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from qimcifa_b200 import abi, multi

    ctx = abi.Context(rank)
    assert multi.connect_peers(ctx, dist, rank, world, "cuda")
    import sympy

    pp = sympy.nextprime(int(2**55 * 1.07))
    n = pp * sympy.nextprime(pp * 1000 // 562)  # 112-bit; the smaller factor sits at 75 % of sqrt(N): nothing near the top
    lo, hi = multi.rank_slice(n, world, rank)
    ctx.sweep(n, 5, hi - 64, hi, batches_per_launch=64)  # warm-up (tables, first launch)
    dist.barrier()
    if rank == 0:
        time.sleep(0.5)
        ctx.peer_signal()
        out_q.put((0, "signalled", 0.0, 0))
    else:
        t0 = time.perf_counter()
        r = multi.sweep_slice_with_peer_flag(lambda a, b: ctx.sweep(n, 5, a, b, batches_per_launch=4096), lo, hi - 64, 4096,
                                             max_seconds=30.0)
        out_q.put((1, "stopped" if r["stopped_by_peer"] else "ran out", time.perf_counter() - t0, r["rounds"]))
        assert ctx.poll_found() == 1
        ctx.set_found(0)
        res = ctx.sweep(n, 5, hi - 64, hi, batches_per_launch=64)  # the flag is per context and can be cleared
        assert res.aborted == 0 and res.batches_done == 64
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_peer_flag_stops_a_running_sweep():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == "signalled"
    assert got[1][1] == "stopped" and 0.3 < got[1][2] < 5.0, got  # stopped shortly after the 0.5 s signal, not after 30 s


def _dispenser_worker(rank, world, port, out_q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from qimcifa_b200 import abi, multi

    ctx = abi.Context(rank)
    assert multi.connect_peers(ctx, dist, rank, world, "cuda")
    import sympy

    # 100-bit semiprime whose smaller factor sits ~1400 batches below the top of the range: 2 GPUs pulling runs of 128
    # batches from ONE counter (in rank 0's memory) reach it after ~6 pulls each
    q = sympy.prevprime(2**50 - 777)
    p = sympy.prevprime(q - 1400 * 223092870 * 2310 // 480)
    n = p * q
    lo, hi = multi.rank_slice(n, 1, 0)  # one node: the whole range, shared by both GPUs
    ctx.sweep(n, 5, lo, lo + 8, batches_per_launch=8)  # warm-up at the bottom (no divisor there)
    if rank == 0:
        ctx.dispenser_reset(0)
    dist.barrier()
    r = multi.sweep_shared_dispenser(lambda a, b: ctx.sweep(n, 5, a, b, batches_per_launch=128),
                                     lambda take: ctx.peer_dispense(0, take), lo, hi, 128, max_seconds=60.0)
    torch.cuda.synchronize()
    dist.barrier()
    handed_out = ctx.peer_dispense(0, 0)
    out_q.put((rank, r["found"], r["factor"] == p if r["found"] else None, r["rounds"], r["stopped_by_peer"], handed_out, p))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_shared_dispenser_over_peer_memory():
    """Both GPUs pull batch runs from one counter with system-scope fetch-adds (one of them over NVLink): every run is
    handed out exactly once, the GPU whose run holds the factor finds it and stops the other through its flag."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_dispenser_worker, args=(r, 2, 29643, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(1 for g in got if g[1]) == 1 and all(g[2] for g in got if g[1]), got  # exactly one finder, the right factor
    loser = [g for g in got if not g[1]][0]
    assert loser[4], got  # the other GPU was stopped by the peer flag, it did not run out of work
    rounds = got[0][3] + got[1][3]
    assert got[0][5] == got[1][5] == 128 * rounds, got  # every pull counted exactly once
    assert 3 <= rounds <= 16, got  # the factor is ~700 batches below sqrt(N): ~6 runs of 128, plus at most one in flight per GPU
