"""CPU-only: the kernels' limb arithmetic (qimcifa_b200/csrc/qcf_device.cuh) compiled for the host from the SAME source
and checked against Python integers: exact divisibility by Hensel/Montgomery reduction in all its forms
(reference include/qimcifa.hpp:369, `toFactor % base == 0`), the carried reciprocal, and the two stages of the chain
kernel's test.  tests/device_math_host.cpp is the C interface; the carry-flag PTX of qcf_addn is device-only (the host
build takes its portable branch), everything else is the code the GPU runs."""
import ctypes
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = 1 << 32
SHAPES = [(2, 1, 1), (2, 1, 2), (2, 2, 1), (2, 2, 2), (3, 1, 2), (3, 1, 3), (3, 2, 1), (3, 2, 2), (3, 2, 3),
          (4, 1, 3), (4, 1, 4), (4, 2, 2), (4, 2, 3), (4, 2, 4), (4, 3, 1), (4, 3, 2)]  # the chain launcher's list


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "tests", "_build", "libdevice_math_host.so")
    src = os.path.join(ROOT, "tests", "device_math_host.cpp")
    hdr = os.path.join(ROOT, "qimcifa_b200", "csrc", "qcf_device.cuh")
    hdr2 = os.path.join(os.path.dirname(hdr), "qcf_filter_math.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in (src, hdr, hdr2)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-Wno-unknown-pragmas",
                               "-I", os.path.dirname(hdr), src, "-o", out])
    lb = ctypes.CDLL(out)
    u32, p32 = ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)
    lb.h_inv32.restype, lb.h_inv32.argtypes = u32, [u32]
    lb.h_ninv_step.restype, lb.h_ninv_step.argtypes = u32, [u32, u32]
    lb.h_divides.restype = ctypes.c_int
    lb.h_divides.argtypes = [ctypes.c_int] * 4 + [p32, p32, u32]
    lb.h_addn.argtypes = [ctypes.c_int, p32, p32]
    lb.h_add32.argtypes = [p32, p32, u32]
    lb.h_less.restype, lb.h_less.argtypes = ctypes.c_int, [p32, p32]
    return lb


def limbs(v, k):
    return (ctypes.c_uint32 * k)(*[(v >> (32 * i)) & (W - 1) for i in range(k)])


def value(arr):
    return sum(int(x) << (32 * i) for i, x in enumerate(arr))


def ninv_of(g):
    return (-pow(g, -1, W)) % W


def divides(lib, form, ln, lg, qn, n, g, ninv=None):
    r = lib.h_divides(form, ln, lg, qn, limbs(n, 4), limbs(g, 3), ninv_of(g) if ninv is None else ninv)
    assert r >= 0, (form, ln, lg, qn)
    return bool(r)


def operands(rnd, ln, lg, qn, divisor):
    """(N, g): g odd with exactly lg limbs' worth of room, N < 2^(32 ln), N // g < 2^(32 qn); g | N iff divisor."""
    while True:
        gb = rnd.randrange(32 * (lg - 1) + 2, 32 * lg + 1)
        g = rnd.randrange(1 << (gb - 1), 1 << gb) | 1
        cmax = min((1 << (32 * qn)) - 1, ((1 << (32 * ln)) - 1) // g)
        if cmax < 2:
            continue
        cb = rnd.randrange(1, cmax.bit_length() + 1)
        c = rnd.randrange(1 << (cb - 1), min(cmax, (1 << cb) - 1) + 1)
        n = g * c
        if not divisor:
            n += rnd.choice([1, 2, g - 1, g // 2 | 1, rnd.randrange(1, g)])
            if n % g == 0 or n >= 1 << (32 * ln) or n // g >= 1 << (32 * qn):
                continue
        return n, g


def test_inv32_and_the_carried_reciprocal(lib):
    rnd = random.Random(1)
    for g in [1, 3, W - 1, 0x10001, 0xFFFF0001] + [rnd.randrange(W) | 1 for _ in range(5000)]:
        x = lib.h_inv32(g)
        assert (x * g) % W == 1
        # one Newton step on the negated reciprocal carries it along a chain whose stride is a multiple of 2^16
        y = (-x) % W
        for _ in range(4):
            g2 = (g + (rnd.randrange(1 << 16) << 16)) % W
            y = lib.h_ninv_step(y, g2)
            assert (y * g2) % W == W - 1, (g, g2)
            g = g2


@pytest.mark.parametrize("shape", SHAPES)
def test_exact_divisibility_every_shape(lib, shape):
    ln, lg, qn = shape
    rnd = random.Random(hash(shape) & 0xFFFF)
    for it in range(1500):
        n, g = operands(rnd, ln, lg, qn, divisor=(it % 2 == 0))
        want = n % g == 0
        assert divides(lib, 0, ln, lg, qn, n, g) == want, (shape, n, g)
        assert divides(lib, 1, ln, lg, qn, n, g) == want, (shape, n, g)
        if lg == 2:
            assert divides(lib, 2, ln, lg, qn, n, g) == want, (shape, n, g)
        if lg == 2 and qn == 2 and ln >= 3:
            assert divides(lib, 3, ln, lg, qn, n, g) == want, (shape, n, g)


def t2_of(n, g):
    """The final t of the two-word reduction as an integer: (N + M g) / W^2, M = -N/g mod W^2."""
    m = (-n * pow(g, -1, W * W)) % (W * W)
    assert (n + m * g) % (W * W) == 0
    return (n + m * g) // (W * W)


def stage1_model(n, g):
    """What qcf_lowlimb_q2g2 may return: True when T2 mod W == g0 (so a true divisor always passes), and ALSO when the low
    limb of T1 is zero -- the one case in 2^32 where the loop's shortcut 'floor((L1 + q1 g0)/W) = hi(q1 g0) + 1' does not
    hold, which it hands to the full test instead of handling (qcf_device.cuh)."""
    g0, g1, n0 = g % W, g // W, n % W
    q0 = (-n0 * pow(g0, -1, W)) % W
    t1 = (q0 * g1 + n // W + (n0 + q0 * g0) // W) % (W * W)
    return t2_of(n, g) % W == g0


@pytest.mark.parametrize("ln", [3, 4])
def test_two_stage_test_low_limb(lib, ln):
    """Stage 1 equals 'T2 mod 2^32 == g0 or L1 == 0' for every operand pair, so a true divisor always passes; stage 1 AND
    stage 2 equals divisibility -- including operands built so that stage 1 passes for a non-divisor."""
    rnd = random.Random(7 + ln)
    for it in range(6000):
        n, g = operands(rnd, ln, 2, 2, divisor=(it % 3 == 0))
        s1 = divides(lib, 4, ln, 2, 2, n, g)
        assert s1 == stage1_model(n, g), (n, g)
        if n % g == 0:
            assert s1
        assert (s1 and divides(lib, 3, ln, 2, 2, n, g)) == (n % g == 0)
    # limb-boundary operands: all-ones limbs, g just below 2^64, cofactor just below 2^64, N just below 2^(32 ln)
    edge = [W - 1, W + 1, (1 << 64) - 1, (1 << 64) - 59, (1 << 63) + 1, (1 << 48) - 59, (1 << 33) + 1]
    for g in edge:
        for c in edge + [1, 2, 3]:
            for d in (0, 1, g - 1):
                n = g * c + d
                if n >= 1 << (32 * ln) or n // g >= 1 << 64:
                    continue
                s1 = divides(lib, 4, ln, 2, 2, n, g)
                assert s1 == stage1_model(n, g), (n, g)
                assert s1 or (t2_of(n, g) % W != g % W)
                assert (s1 and divides(lib, 3, ln, 2, 2, n, g)) == (d == 0), (n, g)
    # non-divisors that PASS stage 1: N = T2 W^2 - M g with T2 = g + j W (same low limb, different high limb)
    made = 0
    while made < 300:
        g = rnd.randrange(1 << 33, 1 << 64) | 1
        j = rnd.randrange(-(g >> 32), W)
        t2 = g + j * W
        if j == 0 or t2 <= 0:
            continue
        lo = max(0, t2 * W * W - (1 << (32 * ln)) + 1)  # M g must exceed this for N < 2^(32 ln)
        m_lo, m_hi = lo // g + 1, min(W * W - 1, (t2 * W * W - 1) // g)
        if m_lo > m_hi:
            continue
        m = rnd.randrange(m_lo, m_hi + 1)
        n = t2 * W * W - m * g
        if not (0 < n < 1 << (32 * ln)) or n // g >= 1 << 64 or n % g == 0:
            continue
        assert t2_of(n, g) == t2
        assert divides(lib, 4, ln, 2, 2, n, g), "stage 1 passes by construction"
        assert not divides(lib, 3, ln, 2, 2, n, g), "stage 2 rejects"
        made += 1


def tq_of(n, g, qn):
    """The final t of a qn-word reduction: (N + M g) / W^qn, M = -N/g mod W^qn."""
    wq = W ** qn
    m = (-n * pow(g, -1, wq)) % wq
    return (n + m * g) // wq


@pytest.mark.parametrize("shape", SHAPES + [(3, 3, 3), (4, 3, 4)])
def test_low_limb_stage_every_shape(lib, shape):
    """qcf_lowlimb == 'limb 0 of the final t equals g0' for every operand pair (so divisors always pass), and together
    with the full test it decides divisibility."""
    ln, lg, qn = shape
    rnd = random.Random(77 + hash(shape) % 1000)
    passed = 0
    for it in range(1500):
        n, g = operands(rnd, ln, lg, qn, divisor=(it % 3 == 0))
        s1 = divides(lib, 5, ln, lg, qn, n, g)
        assert s1 == (tq_of(n, g, qn) % W == g % W), (shape, n, g)
        if n % g == 0:
            assert s1
            passed += 1
        assert (s1 and divides(lib, 0, ln, lg, qn, n, g)) == (n % g == 0)
    assert passed >= 400
    # non-divisors that pass stage 1: final t = g + j W
    made = tries = 0
    while made < 40 and tries < 20000:
        tries += 1
        gb = rnd.randrange(32 * (lg - 1) + 2, 32 * lg + 1)
        g = rnd.randrange(1 << (gb - 1), 1 << gb) | 1
        wq = W ** qn
        j = rnd.randrange(-(g >> 32), W) if lg > 1 else rnd.randrange(1, W)
        t = g + j * W
        if j == 0 or t <= 0:
            continue
        lo = max(0, t * wq - (1 << (32 * ln)) + 1)
        m_lo, m_hi = lo // g + 1, min(wq - 1, (t * wq - 1) // g)
        if m_lo > m_hi:
            continue
        m = rnd.randrange(m_lo, m_hi + 1)
        n = t * wq - m * g
        if not (0 < n < 1 << (32 * ln)) or n // g >= wq or n % g == 0:
            continue
        assert divides(lib, 5, ln, lg, qn, n, g) and not divides(lib, 0, ln, lg, qn, n, g), (shape, n, g)
        made += 1


def test_limb_adds_and_compare(lib):
    rnd = random.Random(3)
    for _ in range(3000):
        lg = rnd.choice([1, 2, 3])
        a, b = rnd.randrange(1 << (32 * lg)), rnd.randrange(1 << (32 * lg))
        if rnd.random() < 0.3:
            a |= (1 << (32 * lg)) - (1 << 31)  # long carry runs
        aa, bb = limbs(a, 3), limbs(b, 3)
        lib.h_addn(lg, aa, bb)
        assert value(aa[:lg]) == (a + b) % (1 << (32 * lg))
        a3, s = rnd.randrange(1 << 96), rnd.randrange(W)
        r = limbs(0, 3)
        lib.h_add32(r, limbs(a3, 3), s)
        assert value(r) == (a3 + s) % (1 << 96)
        b3 = rnd.choice([a3, a3 + 1, a3 - 1, rnd.randrange(1 << 96)]) % (1 << 96)
        assert bool(lib.h_less(limbs(a3, 3), limbs(b3, 3))) == (a3 < b3)
