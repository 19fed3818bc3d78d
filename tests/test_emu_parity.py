"""CPU-only: the library's own kernels, compiled for the host by tests/emu (every kernel runs one block at a time, its
threads as fibers that meet at __syncthreads() and at warp shuffles), against the oracle -- the same test bodies the GPU
runs (tests/test_gpu_parity.py), on the cases small enough for a CPU.  TEST INFRASTRUCTURE: it checks the kernels' control
flow and arithmetic (claims, chains, segments, survivor queue, epochs, reports, counts) where no GPU is available, so that
a kernel change can be verified before GPU time is spent.  It is not a CPU path of the product: the product library
(libqimcifa_cuda.so) has no CPU fallback, and nothing outside this file loads tests/_build/libqimcifa_emu.so."""
import importlib.util
import os
import random
import sys

import pytest

from qimcifa_b200 import abi as product_abi
from tests import test_gpu_parity as gp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(900)  # a kernel emulated at the wrong size would otherwise run for hours


@pytest.fixture(scope="module")
def emu():
    """(abi module bound to the emulator library, context); test_gpu_parity's `abi` is pointed at it meanwhile."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    try:
        import build_emu
    finally:
        sys.path.pop(0)
    lib = build_emu.build()
    spec = importlib.util.spec_from_file_location("qcf_emu_abi", product_abi.__file__)
    eabi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(eabi)
    eabi.LIB_PATH = lib
    saved = gp.abi
    gp.abi = eabi
    ctx = eabi.Context(0)
    assert "emulator" in ctx.device_info()["name"]
    try:
        yield eabi, ctx
    finally:
        ctx.close()
        gp.abi = saved


@pytest.mark.parametrize("level", [5, 6, 7, 9])
def test_emu_device_tables(emu, level):
    gp.test_device_tables_enumerate_the_unit_group(emu[1], level)


@pytest.mark.parametrize("nbits,g_limbs", [(64, 1), (96, 2), (128, 2), (128, 3)])
def test_emu_divisible(emu, nbits, g_limbs):
    gp.test_divisible_matches_python_ints(emu[1], nbits, g_limbs)


@pytest.mark.parametrize("kind", gp.KINDS)
@pytest.mark.parametrize("level", [5, 6, 7, 8, 9])
def test_emu_sweep_indices_matches_oracle(emu, level, kind):
    gp.test_sweep_indices_matches_oracle(emu[1], level, kind)


@pytest.mark.parametrize("kind", gp.KINDS)
def test_emu_sweep_indices_edges(emu, kind):
    gp.test_sweep_indices_edges(emu[1], kind)


def test_emu_filter_kernel_matches_oracle(emu):
    gp.test_filter_kernel_matches_oracle(emu[1], 5)


@pytest.mark.parametrize("nbits", [96, 128])
def test_emu_filter_kernel_equals_exact_kernel(emu, nbits):
    gp.test_filter_kernel_equals_exact_kernel(emu[1], nbits, 5)


def test_emu_filter_kernel_many_divisors(emu):
    gp.test_filter_kernel_many_divisors(emu[1])


def test_emu_gcd_mode_short_intervals(emu):
    gp.test_gcd_mode_matches_oracle_on_short_intervals(emu[1], 5)


@pytest.mark.parametrize("nbits,width,force", [(80, 20_000_000, "1,0,0"), (80, 20_000_000, "1,0,4"), (96, 4_500_000, "1,0,0"), (128, 4_500_000, "1,0,0")])
def test_emu_degree1_packed_trips(emu, nbits, width, force):
    gp.test_degree1_packed_trips_match_oracle(emu[1], nbits, width, force)


def test_emu_survivor_queue_overflow_on_the_spot(emu, monkeypatch):
    """tests/test_gpu_parity.py::test_survivor_queue_overflow_is_tested_on_the_spot on a tenth of its interval."""
    eabi, ctx = emu
    from oracle import oracle_c as oc

    rnd = random.Random(606)
    p = gp._prime_below((1 << 40) - rnd.randrange(1 << 30), rnd)
    q = gp._prime_below(1 << 44, rnd)
    n = p * q
    ip = oc.backward(p) - 1
    lo, hi = ip - 3_300_000, ip + 10_000_000
    want_cnt, want_hits = oc.intent_sweep_indices(n, 5, lo, hi)
    assert want_hits == [p]
    with eabi.options(plan_force="3,9,1", plan_drift=0):
        with eabi.options(epoch_trips=60000):
            r = ctx.sweep_indices(n, 5, lo, hi, flags=eabi.KERNEL_FILTER | eabi.COUNT_ON_DEVICE)
            assert (r.guesses_tested, r.guesses_counted, ctx.last_hits(), r.factor) == (want_cnt, want_cnt, [p], p)
            assert r.kernel_kind == eabi.KERNEL_FILTER and ctx.last_queue_spills() > 0  # the on-the-spot path ran
        r = ctx.sweep_indices(n, 5, lo, hi, flags=eabi.KERNEL_FILTER | eabi.COUNT_ON_DEVICE)
        assert (r.guesses_counted, ctx.last_hits()) == (want_cnt, [p]) and r.kernel_kind == eabi.KERNEL_FILTER
        assert ctx.last_queue_spills() == 0  # epochs sized from the survivor rate: the queue holds them all


@pytest.mark.parametrize("nbits", [96, 128])
def test_emu_exact_chain_kernel_both_loops(emu, monkeypatch, nbits):
    """The exact chain kernel (one-stage and two-stage loop) on an interval a CPU can walk: the chain_min_log2 option lowers the
    threshold of the chain form to 2^16 periods.  Both loops and the row kernel (threshold raised again) report the
    same divisors and counts; a divisor sits in the last partial trip; a non-divisor built to pass stage 1 is rejected."""
    import math

    import sympy

    eabi, ctx = emu
    from oracle import oracle_c as oc

    rnd = random.Random(4000 + nbits)
    half = nbits // 2
    period = 2310
    periods = (1 << 17) + (1 << 16) + 4321  # 6 steps per chain at 2^15 periods per step: one trip of 4 + a tail of 2
    top = (1 << half) - rnd.randrange(1 << (half - 8))
    i_hi = oc.backward(top) - 1
    i_lo = oc.backward(top - periods * period)
    v_lo, v_hi = oc.forward(i_lo), oc.forward(i_hi)
    p = sympy.prevprime(v_lo + ((5 << 15) + 777) * period)  # step 5 of its chain: the tail
    q = sympy.prevprime((1 << (nbits - half)) - rnd.randrange(1 << 20))
    g = sympy.prevprime(rnd.randrange(v_lo + period, v_lo + (6 << 15) * period))
    cases = [(p * q, p), (gp._stage1_false_positive(g, (nbits + 31) // 32, rnd), None)]
    for n, planted in cases:
        out = {}
        for name, two, chain_min in (("one", "0", "16"), ("two", "1", "16"), ("rows", "0", "40")):
            with eabi.options(xchain_two_stage=two, chain_min_log2=chain_min):
                r = ctx.sweep_indices(n, 5, i_lo, i_hi, flags=eabi.KERNEL_EXACT | eabi.COUNT_ON_DEVICE)
            out[name] = (ctx.last_hits(), r.found, r.factor, r.guesses_tested, r.guesses_counted)
            launches = r.kernel_launches
            if name != "rows":
                assert launches >= 2  # a chain launch + the ragged ends
        assert out["one"] == out["two"] == out["rows"], (n, out)
        assert out["two"][3] == out["two"][4] == oc.intent_count(5, i_lo, i_hi)
        if planted is not None:
            assert planted in out["two"][0]
        else:
            assert all(n % h == 0 for h in out["two"][0])


@pytest.mark.parametrize("level", [5, 9, 10])
def test_emu_whole_range_of_a_40_bit_semiprime(emu, level):
    """__graft_entry__.smoke()'s sweep: the whole range of 1000003 * 1000033 with the automatic kernel choice (filter
    launches for the upper octaves, exact rows below), counts against the closed form -- incl. level 10's corrections."""
    eabi, ctx = emu
    from oracle import oracle_c as oc

    n = 1000003 * 1000033
    _, _, _, batch_count = eabi.node_split(n, 1)
    res = ctx.sweep(n, level, 0, batch_count, flags=eabi.COUNT_ON_DEVICE)
    assert res.found and res.factor == 1000003 and res.kernel_launches > 5
    assert res.guesses_tested == res.guesses_counted == oc.intent_count(level, 2, eabi.BIGGEST_WHEEL + 1)


def test_emu_random_mode(emu, monkeypatch):
    from tests import test_gpu_random as gr

    monkeypatch.setattr(gr, "abi", emu[0])
    gr.test_device_guesses_follow_the_documented_map(emu[1], 128, 7)
    gr.test_random_mode_finds_the_factor_deterministically(emu[1])


def test_emu_found_flag_and_errors(emu):
    eabi, ctx = emu
    n = (2**48 - 59) * (2**48 - 65)
    _, _, _, bc = eabi.node_split(n, 1)
    ctx.set_found(1)
    assert ctx.poll_found() == 1
    res = ctx.sweep(n, 5, bc - 64, bc - 60, batches_per_launch=2)
    assert res.aborted == 1 and res.found == 0 and res.batches_done == 0
    ctx.set_found(0)
    assert ctx.poll_found() == 0
    with pytest.raises(eabi.QcfError):
        ctx.sweep(35, 4, 0, 1)
    with pytest.raises(eabi.QcfError):
        ctx.divisible(15, [4], 1)


@pytest.mark.parametrize("seed", [int(s) for s in os.environ.get("QCF_EMU_FUZZ_SEEDS", "0,5,9,13").split(",")])
def test_emu_fuzz_random_intervals(emu, monkeypatch, seed):
    """tests/test_gpu_fuzz.py on the emulation: random N, level, depth and interval through all three kernel selections."""
    from tests import test_gpu_fuzz as gf

    monkeypatch.setattr(gf, "abi", emu[0])
    gf.test_random_intervals_match_oracle(emu[1], seed)


@pytest.fixture(scope="module")
def emu_bin(emu, tmp_path_factory):
    """The product's C++ host drivers (qimcifa_b200/host) linked against the emulation library."""
    import subprocess

    out = os.path.join(ROOT, "tests", "_build", "emu_bin")
    os.makedirs(out, exist_ok=True)
    host = os.path.join(ROOT, "qimcifa_b200", "host")
    lib_dir = os.path.join(ROOT, "tests", "_build")
    lib = os.path.join(lib_dir, "libqimcifa_emu.so")
    for name in ("qimcifa", "qimcifa_tuner"):
        exe, src = os.path.join(out, name), os.path.join(host, name + ".cpp")
        deps = [src, lib] + [os.path.join(host, h) for h in os.listdir(host) if h.endswith(".hpp")]
        if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", host, src,
                                   "-o", exe, "-L", lib_dir, "-lqimcifa_emu", "-Wl,-rpath," + lib_dir, "-ldl"])
    return out


def test_emu_host_driver_dialog_matches_reference_stdout(emu_bin, monkeypatch, golden_dir):
    """tests/test_gpu_cli.py on the emulation: the `qimcifa` dialog and result lines against the reference's recorded
    stdout (byte for byte at level 5, nodeCount 1), the node dialog, the level clamp and the pre-checks."""
    from tests import test_gpu_cli as gc

    monkeypatch.setattr(gc, "BIN", emu_bin)
    gc.test_qimcifa_dialog_matches_reference_output(golden_dir)
    gc.test_qimcifa_node_dialog_and_level_clamp()


def test_emu_tuner_calibration_round_trip(emu_bin, monkeypatch, tmp_path):
    from tests import test_gpu_cli as gc

    monkeypatch.setattr(gc, "BIN", emu_bin)
    gc.test_tuner_calibration_round_trip(tmp_path)


def test_emu_guesses_above_2_pow_64_and_whole_64_bit_ranges(emu):
    """Three-limb guesses in the padded top batches of a 128-bit N, the square of a 32-bit prime, a 64-bit prime's whole
    range and tiny N, with the automatic kernel choice (the exact-kernel legs of these tests are left to the GPU)."""
    gp.test_guesses_above_2_pow_64(emu[1], emu[0].KERNEL_AUTO)
    gp.test_square_prime_and_tiny_inputs(emu[1], emu[0].KERNEL_AUTO)


def test_emu_gcd_chain_kernel_finds_planted_multiples(emu, monkeypatch):
    """tests/test_gpu_parity.py::test_gcd_mode_chain_kernel_finds_planted_multiples on 2^17 periods (chain threshold lowered):
    the Montgomery-product chain kernel reports exactly the guesses that share the planted prime with N."""
    import math

    eabi, ctx = emu
    from oracle import oracle_c as oc

    rnd = random.Random(78)
    q = gp._prime_below(1 << 45, rnd)
    p2 = gp._prime_below(1 << 22, rnd)
    n2 = p2 * q
    i_hi = oc.backward(1 << 44) - 1
    i_lo = i_hi - 480 * ((1 << 17) + 999) - 12345
    v_lo, v_hi = oc.forward(i_lo), oc.forward(i_hi)
    want2 = sorted(m * p2 for m in range(v_lo // p2, v_hi // p2 + 2)
                   if v_lo <= m * p2 <= v_hi and all((m * p2) % s for s in gp.PRIMES[:5]))
    assert 5 < len(want2) <= 64
    with eabi.options(chain_min_log2=16):
        res2 = ctx.sweep_indices(n2, 5, i_lo, i_hi, flags=eabi.MODE_GCD | eabi.COUNT_ON_DEVICE)
    assert res2.kernel_launches >= 2  # the chain launch + ragged ends
    assert ctx.last_hits() == want2 and res2.found
    assert res2.guesses_counted == res2.guesses_tested == oc.intent_count(5, i_lo, i_hi)
    first = min(want2, key=lambda g: (-((oc.backward(g) - 1 - 2) // oc.BW), g))
    assert res2.factor == math.gcd(first, n2) == p2
    # the same planted prime in numbers whose top bits are all ones, as 3-limb and as 4-limb numbers: there the Montgomery
    # accumulator cannot stay unreduced (qcf_montmul_acc LAZY, csrc/qcf_gcd.cu) and the kernel subtracts after every product
    import sympy

    for top in (96, 128):
        n3 = p2 * sympy.prevprime((1 << top) // p2)
        assert n3.bit_length() == top and (n3 >> (top - 20)) == (1 << 20) - 1
        with eabi.options(chain_min_log2=16):
            res3 = ctx.sweep_indices(n3, 5, i_lo, i_hi, flags=eabi.MODE_GCD | eabi.COUNT_ON_DEVICE)
        assert ctx.last_hits() == want2 and res3.found and res3.factor == p2, top
        assert res3.guesses_counted == res3.guesses_tested


def test_emu_superset_mode(emu):
    """Superset mode: filter chains run over a lower wheel level than requested, so a divisor that is a multiple of one of
    the level's remaining primes survives the filter -- it must still never be reported at that level (levels 7, 9, 10)."""
    import sympy

    eabi, ctx = emu
    from oracle import oracle_c as oc

    rnd = random.Random(99)
    r = sympy.prevprime((1 << 40) - rnd.randrange(1 << 30))
    q = sympy.prevprime((1 << 44) - rnd.randrange(1 << 30))
    g = 17 * r
    n = g * q
    ig = oc.backward(g) - 1
    lo, hi = ig - 1_500_000, ig + 1_500_000
    for kind in (eabi.KERNEL_FILTER, eabi.KERNEL_AUTO):
        res = ctx.sweep_indices(n, 6, lo, hi, flags=kind | eabi.COUNT_ON_DEVICE)
        assert ctx.last_hits() == [g] and res.guesses_counted == res.guesses_tested == oc.intent_count(6, lo, hi)
        for lvl in (7, 9, 10):
            res = ctx.sweep_indices(n, lvl, lo, hi, flags=kind | eabi.COUNT_ON_DEVICE)
            assert res.found == 0 and ctx.last_hits() == []
            assert res.guesses_counted == res.guesses_tested == oc.intent_count(lvl, lo, hi)


def test_emu_signal_reaches_more_than_32_peers(emu):
    """qcf_peer_connect accepts up to 64 ranks = 63 peers: the signal kernel must raise EVERY peer's found flag, whatever
    its launch shape (it used to be one 32-thread block with one peer per thread).  Host logic + the kernel's own loop on
    the emulation, where a "peer's device memory" is host memory; the launcher is looked up by its C++ symbol."""
    import ctypes
    import subprocess

    eabi, _ = emu
    nm = subprocess.run(["nm", "-D", "--defined-only", eabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    sym = [ln.split()[-1] for ln in nm.splitlines() if "qcf_launch_signal_peers" in ln]
    assert len(sym) == 1, sym
    fn = getattr(ctypes.CDLL(eabi.LIB_PATH), sym[0])
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    for n_peers in (1, 31, 32, 33, 63):
        flags = (ctypes.c_uint32 * 64)()
        ptrs = (ctypes.c_void_p * 64)(*[ctypes.addressof(flags) + 4 * i for i in range(64)])
        assert fn(ctypes.cast(ptrs, ctypes.c_void_p), n_peers, None) == 0
        assert list(flags) == [1] * n_peers + [0] * (64 - n_peers), (n_peers, list(flags))


def test_emu_chain_draw_random_mode(emu):
    """tests/test_gpu_random.py::chain_draw_round_trip on the emulation: the chain-draw random mode (Philox draws row groups
    of a random window, the filter kernel walks them) tests exactly the guesses of the documented map."""
    from tests import test_gpu_random as gr

    eabi, ctx = emu
    gr.chain_draw_round_trip(ctx, eabi, 64, 5, seed=3, samples=2, force="2,14,0")
    gr.chain_draw_round_trip(ctx, eabi, 64, 7, seed=4, samples=2, force="2,14,0")
