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
    out = multi.check_peer_flag(ctx, dist, rank, world, n)  # the same check `bench.py --gpus N` runs (N >= 2)
    out_q.put((rank, out))
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
    got = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] and got[0][1]["ok"], got  # stopped shortly after the 0.5 s signal, not after 30 s
    swept = [x for x in got[0][1]["ranks"] if x["role"] == "swept"]
    assert len(swept) == 1 and swept[0]["stopped_by_peer"] and 0.3 < swept[0]["seconds"] < 5.0, got


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
    out = multi.check_shared_dispenser(ctx, dist, rank, world, n, p, chunk=128)  # the same check `bench.py --gpus N` runs
    out_q.put((rank, out))
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
    got = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] and got[0][1]["ok"], got
    ranks = got[0][1]["ranks"]
    assert sum(1 for x in ranks if x["found"]) == 1, got  # exactly one finder
    assert 3 <= got[0][1]["rounds"] <= 16, got  # the factor is ~700 batches below sqrt(N): ~6 runs of 128, plus at most one in flight per GPU
