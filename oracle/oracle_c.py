"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY).

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

U64 = ctypes.c_uint64
U64x2 = U64 * 2
MASK64 = (1 << 64) - 1


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_mt_sweep.argtypes = [ctypes.c_void_p, ctypes.c_int, U64, U64, U64, ctypes.c_uint,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB.orc_literal_sweep.argtypes = [ctypes.c_void_p, ctypes.c_int, U64, U64, U64, U64, ctypes.c_void_p, U64,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB.orc_literal_sweep_mode.argtypes = [ctypes.c_void_p, ctypes.c_int, U64, U64, U64, U64, ctypes.c_int, ctypes.c_void_p,
                                                U64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _LIB.orc_literal_batches.argtypes = [U64, U64, U64, ctypes.c_void_p, ctypes.c_int]
        _LIB.orc_node_split.argtypes = [ctypes.c_void_p, U64, U64] + [ctypes.c_void_p] * 4
        _LIB.orc_tuner_batch_count.argtypes = [ctypes.c_void_p, U64, ctypes.c_void_p]
    return _LIB


def _p(v):
    return U64x2(v & MASK64, (v >> 64) & MASK64)


def _g(a):
    return int(a[0]) | (int(a[1]) << 64)


BW = 223092870


def wheel11():
    out = (ctypes.c_uint16 * 480)()
    lib().orc_wheel11(out)
    return list(out)


def forward(p):
    o = U64x2()
    lib().orc_forward(_p(p), o)
    return _g(o)


def backward(n):
    o = U64x2()
    lib().orc_backward(_p(n), o)
    return _g(o)


def isqrt(n):
    o = U64x2()
    lib().orc_isqrt(_p(n), o)
    return _g(o)


def node_split(n, node_count, bw=BW):
    s, fr, nr, bc = U64x2(), U64x2(), U64x2(), U64x2()
    lib().orc_node_split(_p(n), node_count, bw, s, fr, nr, bc)
    return _g(s), _g(fr), _g(nr), _g(bc)


def intent_sweep_indices(n, level, p_lo, p_hi, max_hits=64, gcd_mode=False):
    """gcd_mode: the common-factor build (IS_RSA_SEMIPRIME=OFF): a guess hits when gcd(guess, N) != 1."""
    tested = U64(0)
    hits = (U64 * (2 * max_hits))()
    nh = lib().orc_intent_sweep_indices_mode(_p(n), level, _p(p_lo), _p(p_hi), int(gcd_mode), ctypes.byref(tested), hits, max_hits)
    return tested.value, [int(hits[2 * i]) | (int(hits[2 * i + 1]) << 64) for i in range(min(nh, max_hits))]


def intent_count(level, p_lo, p_hi):
    o = U64x2()
    lib().orc_intent_count(level, _p(p_lo), _p(p_hi), o)
    return _g(o)


def tuner_cardinalities(n):
    out = (U64 * 10)()
    lib().orc_tuner_cardinalities(_p(n), out)
    return {5 + i: int(out[2 * i]) | (int(out[2 * i + 1]) << 64) for i in range(5)}


def tuner_batch_count(n, bw=BW):
    o = U64x2()
    lib().orc_tuner_batch_count(_p(n), bw, o)
    return _g(o)


def literal_batches(node_range, node_count, node_id, cap=4096):
    out = (U64 * cap)()
    k = lib().orc_literal_batches(node_range, node_count, node_id, out, cap)
    return [int(out[i]) for i in range(min(k, cap))]


def literal_sweep(n, level, node_count=1, node_id=0, bw=BW, max_guesses=0, trace_cap=64, gcd_mode=False):
    trace = (U64 * (2 * max(trace_cap, 1)))()
    tested, h = U64(0), U64(0)
    fac = U64x2()
    found = lib().orc_literal_sweep_mode(_p(n), level, node_count, node_id, bw, max_guesses, int(gcd_mode), trace, trace_cap,
                                         ctypes.byref(tested), fac, ctypes.byref(h))
    k = min(tested.value, trace_cap)
    first = [int(trace[2 * i]) | (int(trace[2 * i + 1]) << 64) for i in range(k)]
    return (_g(fac) if found else None), tested.value, first, "%016x" % h.value


def mt_sweep(n, level, batch_lo, batch_hi, threads, bw=BW):
    tested = U64(0)
    sec = ctypes.c_double(0)
    fac = U64x2()
    found = lib().orc_mt_sweep(_p(n), level, batch_lo, batch_hi, bw, threads, ctypes.byref(tested), ctypes.byref(sec), fac)
    return (_g(fac) if found else None), tested.value, sec.value
