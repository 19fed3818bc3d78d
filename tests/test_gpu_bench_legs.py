"""GPU test of bench.py's whole_range_sweep leg: every rank sweeps its complete nodeCount slice of a prime; the slices tile the
range (reference src/qimcifa.cpp:149-158) and every guess of every batch is decided (include/qimcifa.hpp:369)."""
import pytest

pytestmark = pytest.mark.gpu


def test_whole_range_sweep_tiles_the_range_and_counts_every_guess():
    import torch

    import bench
    from qimcifa_b200 import abi

    with abi.Context(0) as ctx:
        bits, level = 64, 5
        n_full = bench._prev_prime((1 << bits) - 12345)
        for world in (1, 4):
            _, _, node_range, batch_count = abi.node_split(n_full, world)
            total = 0
            for rank in range(world):
                out = bench.whole_range_sweep(ctx, abi, torch, None, bits, level, world, rank, abi.KERNEL_AUTO, chunk=3)
                assert out["n"] == str(n_full) and out["n_gpus"] == world and out["batches"] == batch_count
                assert out["seconds"] > 0 and len(out["seconds_by_rank"]) == 1
                lo, hi = abi.node_batches(batch_count, node_range, rank)
                assert out["guesses"] == abi.count_guesses(level, lo, hi), (world, rank)
                total += out["guesses"]
            assert total == abi.count_guesses(level, 0, batch_count), world
