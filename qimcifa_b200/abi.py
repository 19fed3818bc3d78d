"""ctypes binding of libqimcifa_cuda.so (include/qimcifa_cuda.h).

This is the same boundary the C++ host drivers (qimcifa_b200/host/) link against; the Python layer
exists for the parity tests and bench.py.  There is no CPU fallback: loading fails loudly when the
library has not been built, and creating a context fails loudly when there is no CUDA device.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libqimcifa_cuda.so")
if os.environ.get("QCF_LIB_VARIANT"):  # kernel A/B experiments (qimcifa_b200/build.py build_variant; tools/ only)
    LIB_PATH = os.path.join(_PKG, "libqimcifa_cuda_%s.so" % os.environ["QCF_LIB_VARIANT"])

BIGGEST_WHEEL = 223092870
MAX_HITS = 64
KERNEL_AUTO, KERNEL_EXACT, KERNEL_FILTER, COUNT_ON_DEVICE, MODE_GCD, RANDOM_CHAINS = 0, 1, 2, 4, 8, 16

# every symbol include/qimcifa_cuda.h declares
EXPORTS = (
    "qcf_create", "qcf_destroy", "qcf_set_stream", "qcf_last_error", "qcf_device_info",
    "qcf_build_tables", "qcf_get_tables", "qcf_sweep", "qcf_sweep_indices", "qcf_last_hits",
    "qcf_node_split", "qcf_count_guesses", "qcf_set_found", "qcf_poll_found", "qcf_time_level",
    "qcf_divisible", "qcf_microbench", "qcf_sweep_random", "qcf_random_guesses",
    "qcf_peer_export", "qcf_peer_connect", "qcf_peer_signal", "qcf_peer_disconnect", "qcf_plan_filter",
    "qcf_dispenser_reset", "qcf_peer_dispense", "qcf_last_plan", "qcf_set_option", "qcf_io_bytes", "qcf_random_chain_guesses",
    "qcf_last_queue_spills",
)


class QcfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"qcf error {code}: {msg}")
        self.code = code


class Result(ctypes.Structure):
    _fields_ = [
        ("found", ctypes.c_int32),
        ("aborted", ctypes.c_int32),
        ("factor_le", ctypes.c_uint32 * 4),
        ("n_hits", ctypes.c_uint32),
        ("kernel_kind", ctypes.c_uint32),
        ("guesses_tested", ctypes.c_uint64),
        ("guesses_counted", ctypes.c_uint64),
        ("batches_done", ctypes.c_uint64),
        ("next_batch_hi", ctypes.c_uint64),
        ("kernel_launches", ctypes.c_uint64),
        ("device_seconds", ctypes.c_double),
    ]

    @property
    def factor(self):
        return limbs_to_int(self.factor_le) if self.found else None


class FilterPlan(ctypes.Structure):
    """qcf_filter_plan: the filter-kernel launch the planner chooses (test hook, no GPU needed)."""
    _fields_ = [
        ("found", ctypes.c_int32), ("degree", ctypes.c_int32),
        ("tprime", ctypes.c_uint32), ("align", ctypes.c_uint32), ("steps", ctypes.c_uint32), ("trips", ctypes.c_uint32),
        ("unroll", ctypes.c_uint32), ("n_seg", ctypes.c_uint32), ("slack", ctypes.c_uint32), ("fbits", ctypes.c_uint32),
        ("seg_trips", ctypes.c_uint32 * 32), ("seg_shift", ctypes.c_uint32 * 32), ("seg_bound", ctypes.c_uint32 * 32),
        ("m_lo_le", ctypes.c_uint32 * 4), ("m_end_le", ctypes.c_uint32 * 4),
        ("c_lo", ctypes.c_uint64), ("period", ctypes.c_uint32), ("cost", ctypes.c_double),
        ("drift", ctypes.c_uint32), ("mu0", ctypes.c_uint64),
        ("seg_nu_lo", ctypes.c_uint32 * 32), ("seg_nu_hi", ctypes.c_uint32 * 32),
        ("seg_lambda", ctypes.c_uint64 * 32), ("seg_ka", ctypes.c_uint64 * 32), ("seg_window", ctypes.c_uint64 * 32),
        ("seg_beta_le", (ctypes.c_uint32 * 4) * 32),
        ("rho", ctypes.c_uint32), ("pinv32", ctypes.c_uint32), ("sigma0", ctypes.c_uint64),
        ("seg_sigma", ctypes.c_uint64 * 32), ("seg_kappa", ctypes.c_uint32 * 32),
        ("packed", ctypes.c_uint32), ("reserved_", ctypes.c_uint32),
    ]


_LIB = None


def load():
    """Loads the CUDA library; raises if it is missing (run `python -m qimcifa_b200.build`)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m qimcifa_b200.build` "
                "(qimcifa_b200 has no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.qcf_last_error.restype = ctypes.c_char_p
        u32p, u64p, vp = ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint64), ctypes.c_void_p
        lib.qcf_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        lib.qcf_destroy.argtypes = [vp]
        lib.qcf_set_stream.argtypes = [vp, vp]
        lib.qcf_device_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_int]
        lib.qcf_build_tables.argtypes = [vp, ctypes.c_int]
        lib.qcf_get_tables.argtypes = [vp, ctypes.c_int, u32p, u32p, u32p, u32p, ctypes.c_uint32, u32p, ctypes.c_uint32]
        lib.qcf_sweep.argtypes = [vp, u32p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                  ctypes.c_uint32, ctypes.POINTER(Result)]
        lib.qcf_sweep_indices.argtypes = [vp, u32p, ctypes.c_int, ctypes.c_int, u32p, u32p, ctypes.c_uint32, ctypes.POINTER(Result)]
        lib.qcf_last_hits.argtypes = [vp, u32p, ctypes.c_uint32, u32p]
        lib.qcf_node_split.argtypes = [u32p, ctypes.c_int, ctypes.c_uint64, u32p, u32p, u64p, u64p]
        lib.qcf_count_guesses.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, u32p]
        lib.qcf_set_found.argtypes = [vp, ctypes.c_int]
        lib.qcf_poll_found.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
        lib.qcf_time_level.argtypes = [vp, u32p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32,
                                       ctypes.POINTER(ctypes.c_double), u64p]
        lib.qcf_divisible.argtypes = [vp, u32p, ctypes.c_int, u32p, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint8)]
        lib.qcf_sweep_random.argtypes = [vp, u32p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.c_uint32, ctypes.POINTER(Result)]
        lib.qcf_random_guesses.argtypes = [vp, u32p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, u32p]
        u8p = ctypes.POINTER(ctypes.c_uint8)
        lib.qcf_random_chain_guesses.argtypes = [vp, u32p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, u32p]
        lib.qcf_peer_export.argtypes = [vp, u8p]
        lib.qcf_peer_connect.argtypes = [vp, u8p, ctypes.c_int, ctypes.c_int]
        lib.qcf_peer_signal.argtypes = [vp]
        lib.qcf_peer_disconnect.argtypes = [vp]
        lib.qcf_dispenser_reset.argtypes = [vp, ctypes.c_uint64]
        lib.qcf_peer_dispense.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, u64p]
        lib.qcf_microbench.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        lib.qcf_plan_filter.argtypes = [u32p, ctypes.c_int, ctypes.c_int, u32p, u32p, ctypes.c_int, ctypes.POINTER(FilterPlan)]
        lib.qcf_last_plan.argtypes = [vp, ctypes.POINTER(FilterPlan)]
        lib.qcf_last_queue_spills.argtypes = [vp, u64p]
        lib.qcf_set_option.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.qcf_io_bytes.argtypes = [u32p, u32p, u32p]
        _LIB = lib
    return _LIB


def io_bytes():
    """-> (H2D bytes per chain launch, H2D bytes per row launch, D2H bytes per sweep call): sizes of the parameter blocks / hit record."""
    a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    _check(load().qcf_io_bytes(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def set_option(name, value=None):
    """qcf_set_option: a test / calibration switch of the library (process-wide); value None restores the default."""
    _check(load().qcf_set_option(name.encode(), None if value is None else str(value).encode()))


class options:
    """with abi.options(plan_force="3,9,1", epoch_trips=60000): ...  -- switches set inside the block, defaults restored after."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        for k, v in self.kw.items():
            set_option(k, v)
        return self

    def __exit__(self, *a):
        for k in self.kw:
            set_option(k, None)


def int_to_limbs(v, n=4):
    if v < 0 or v >> (32 * n):
        raise OverflowError(f"{v} does not fit {n} 32-bit limbs")
    return (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def limbs_to_int(a):
    return sum(int(x) << (32 * i) for i, x in enumerate(a))


def _check(rc):
    if rc != 0:
        raise QcfError(rc, load().qcf_last_error().decode())


def node_split(n, node_count=1):
    """src/qimcifa.cpp:128,149,155-158 -> (floor sqrt N, fullRange, nodeRange, batchCount)."""
    s, fr = (ctypes.c_uint32 * 4)(), (ctypes.c_uint32 * 4)()
    nr, bc = ctypes.c_uint64(), ctypes.c_uint64()
    _check(load().qcf_node_split(int_to_limbs(n), 4, node_count, s, fr, ctypes.byref(nr), ctypes.byref(bc)))
    return limbs_to_int(s), limbs_to_int(fr), nr.value, bc.value


def count_guesses(level, batch_lo, batch_hi):
    c = (ctypes.c_uint32 * 4)()
    _check(load().qcf_count_guesses(level, batch_lo, batch_hi, c))
    return limbs_to_int(c)


def plan_filter(n, level, v_lo, v_hi, forced=False):
    """The filter-kernel launch planned for the whole periods inside the value interval [v_lo, v_hi] (None: no plan)."""
    pl = FilterPlan()
    _check(load().qcf_plan_filter(int_to_limbs(n), 4, level, int_to_limbs(v_lo), int_to_limbs(v_hi), int(forced), ctypes.byref(pl)))
    return pl if pl.found else None


def node_batches(batch_count, node_range, node_id):
    """Intent split: node k owns [batchCount-(k+1)R, batchCount-kR), swept descending."""
    return batch_count - (node_id + 1) * node_range, batch_count - node_id * node_range


class Context:
    """One GPU context (one per device / rank)."""

    def __init__(self, device=0, stream=None):
        self._h = ctypes.c_void_p()
        _check(load().qcf_create(device, ctypes.byref(self._h)))
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if self._h:
            load().qcf_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        _check(load().qcf_set_stream(self._h, ctypes.c_void_p(cuda_stream_handle)))

    def device_info(self):
        sm, khz = ctypes.c_int(), ctypes.c_int()
        name = ctypes.create_string_buffer(128)
        _check(load().qcf_device_info(self._h, ctypes.byref(sm), ctypes.byref(khz), name, 128))
        return {"sm_count": sm.value, "clock_khz": khz.value, "name": name.value.decode()}

    def build_tables(self, level):
        _check(load().qcf_build_tables(self._h, level))

    def get_tables(self, level):
        period, na, nb = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        _check(load().qcf_get_tables(self._h, level, ctypes.byref(period), ctypes.byref(na), ctypes.byref(nb), None, 0, None, 0))
        ta, tb = (ctypes.c_uint32 * na.value)(), (ctypes.c_uint32 * nb.value)()
        _check(load().qcf_get_tables(self._h, level, None, None, None, ta, na.value, tb, nb.value))
        return period.value, list(ta), list(tb)

    def sweep(self, n, level, batch_lo, batch_hi, batches_per_launch=0, flags=KERNEL_AUTO):
        res = Result()
        _check(load().qcf_sweep(self._h, int_to_limbs(n), 4, level, batch_lo, batch_hi, batches_per_launch, flags, ctypes.byref(res)))
        return res

    def sweep_indices(self, n, level, p_lo, p_hi, flags=KERNEL_AUTO):
        res = Result()
        _check(load().qcf_sweep_indices(self._h, int_to_limbs(n), 4, level, int_to_limbs(p_lo), int_to_limbs(p_hi), flags, ctypes.byref(res)))
        return res

    def last_hits(self):
        buf = (ctypes.c_uint32 * (4 * MAX_HITS))()
        nh = ctypes.c_uint32()
        _check(load().qcf_last_hits(self._h, buf, MAX_HITS, ctypes.byref(nh)))
        return sorted(limbs_to_int(buf[4 * i:4 * i + 4]) for i in range(nh.value))

    def last_plan(self):
        """The plan of the largest filter launch of the last sweep call on this context (None: no filter launch)."""
        pl = FilterPlan()
        _check(load().qcf_last_plan(self._h, ctypes.byref(pl)))
        return pl if pl.found else None

    def last_queue_spills(self):
        """Survivors the filter launches of the last sweep call tested on the spot because their block's queue was full."""
        n = ctypes.c_uint64(0)
        _check(load().qcf_last_queue_spills(self._h, ctypes.byref(n)))
        return n.value

    def set_found(self, value=1):
        _check(load().qcf_set_found(self._h, int(value)))

    def poll_found(self):
        v = ctypes.c_int()
        _check(load().qcf_poll_found(self._h, ctypes.byref(v)))
        return v.value

    def time_level(self, n, level, batches=1, flags=KERNEL_AUTO):
        sec, g = ctypes.c_double(), ctypes.c_uint64()
        _check(load().qcf_time_level(self._h, int_to_limbs(n), 4, level, batches, flags, ctypes.byref(sec), ctypes.byref(g)))
        return sec.value, g.value

    def divisible(self, n, guesses, g_limbs):
        cnt = len(guesses)
        flat = (ctypes.c_uint32 * (cnt * g_limbs))()
        for i, g in enumerate(guesses):
            for k in range(g_limbs):
                flat[i * g_limbs + k] = (g >> (32 * k)) & 0xFFFFFFFF
        out = (ctypes.c_uint8 * cnt)()
        _check(load().qcf_divisible(self._h, int_to_limbs(n), 4, flat, g_limbs, cnt, out))
        return [bool(x) for x in out]

    def sweep_random(self, n, level, seed, counter_lo, budget, flags=KERNEL_AUTO):
        res = Result()
        _check(load().qcf_sweep_random(self._h, int_to_limbs(n), 4, level, seed, counter_lo, budget, flags, ctypes.byref(res)))
        return res

    def random_guesses(self, n, level, seed, counter_lo, count):
        buf = (ctypes.c_uint32 * (4 * count))()
        _check(load().qcf_random_guesses(self._h, int_to_limbs(n), 4, level, seed, counter_lo, count, buf))
        return [limbs_to_int(buf[4 * i:4 * i + 4]) for i in range(count)]

    def random_chain_guesses(self, n, level, seed, counter_lo, count):
        """The guesses of counters [counter_lo, counter_lo + count) under the chain-draw map (0 = a skipped counter)."""
        buf = (ctypes.c_uint32 * (4 * count))()
        _check(load().qcf_random_chain_guesses(self._h, int_to_limbs(n), 4, level, seed, counter_lo, count, buf))
        return [limbs_to_int(buf[4 * i:4 * i + 4]) for i in range(count)]

    def peer_export(self):
        buf = (ctypes.c_uint8 * 64)()
        _check(load().qcf_peer_export(self._h, buf))
        return bytes(buf)

    def peer_connect(self, handles, self_rank):
        """handles: list of the 64-byte handles of ALL ranks, in rank order."""
        flat = (ctypes.c_uint8 * (64 * len(handles))).from_buffer_copy(b"".join(handles))
        _check(load().qcf_peer_connect(self._h, flat, len(handles), self_rank))

    def peer_signal(self):
        _check(load().qcf_peer_signal(self._h))

    def dispenser_reset(self, value=0):
        _check(load().qcf_dispenser_reset(self._h, value))

    def peer_dispense(self, owner_rank, take):
        """Batches handed out before this call from rank `owner_rank`'s dispenser; the caller owns the next `take`."""
        start = ctypes.c_uint64()
        _check(load().qcf_peer_dispense(self._h, owner_rank, take, ctypes.byref(start)))
        return start.value

    def peer_disconnect(self):
        _check(load().qcf_peer_disconnect(self._h))

    def microbench(self, kind, iters=1 << 14):
        ops, sec = ctypes.c_double(), ctypes.c_double()
        _check(load().qcf_microbench(self._h, kind, iters, ctypes.byref(ops), ctypes.byref(sec)))
        return ops.value, sec.value
