#!/usr/bin/env python3
"""bench.py -- wheel-filtered guesses/sec of the reverse-wheel trial-division sweep on B200.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation
                                                            (oracle/_ref = unmodified reference sources + shims,
                                                            else the oracle port) on the box's host cores
Under torchrun (N > 1) every rank is one GPU = one node of the reference's nodeCount/nodeId split
(src/qimcifa.cpp:155-158, intent semantics); rank 0 prints ONE JSON line.

A "step" is one pass of the hot path over one batch of synthetic input: a sweep of `--batches`
reference batches (223092870 wheel-11 indices each) of the rank's slice of a synthetic semiprime of
`--bits` bits, at wheel level `--level`, through qcf_sweep().  Successive steps walk DOWN the slice as
the reference's dispenser does, so no step repeats another's guesses.  The divisor is placed outside the
swept range, so every step covers its batches completely (whole-job throughput, no early exit).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Per-kernel profiler figures (warp instructions per guess, DRAM bytes per launch) are NOT literals here: they are read from
# profiles/current_kernels.json, each entry pinned to the SASS of the kernel it was captured from (sha256 of
# `cuobjdump -sass -fun <kernel>`), and they are printed only when the library loaded by this run carries the same SASS.
KERNEL_PROFILES = os.path.join(ROOT, "profiles", "current_kernels.json")

METRIC = "wheel_filtered_guesses_per_sec"
UNIT = "guesses/s"


# ---------------------------------------------------------------------------------------------------
# synthetic inputs: deterministic semiprimes (counter-based, Miller-Rabin), seeds recorded in the JSON
# ---------------------------------------------------------------------------------------------------
def _is_prime(n):
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:  # deterministic for n < 3.3e24; enough witnesses for 64-bit primes
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def _prev_prime(x):
    x -= 1 - (x & 1)
    while not _is_prime(x):
        x -= 2
    return x


def _splitmix(seed):
    z = (seed + 0x9E3779B97F4A7C15) & (2**64 - 1)
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return z ^ (z >> 31)


def synthetic_semiprime(bits, seed):
    """Two primes of bits/2 bits (top bit set), p < q, with p placed just ABOVE 75 % of sqrt(N): that is a slice
    boundary of the nodeCount = 2, 4 and 8 splits, so at every N the rank whose slice holds p meets it only at the very
    bottom of the slice and the timed steps (which walk down from the top of each slice) cover their batches completely."""
    h = bits // 2
    r1, r2 = _splitmix(seed), _splitmix(seed + 1)
    p = _prev_prime(int((1 << (h - 1)) * 1.07) + (r1 % (1 << (h - 8))))
    q = _prev_prime(p * 1000 // 562 * 995 // 1000 - (r2 % (1 << (h - 12))))
    assert p.bit_length() == h and q.bit_length() == h and p < q and (p * q).bit_length() == bits
    return p, q, p * q


_W11 = [v for v in range(1, 2310) if v % 2 and v % 3 and v % 5 and v % 7 and v % 11]


def _index_of(value):
    """0-based rank of `value` among the integers coprime to 2310 (inverse of forward(), include/qimcifa.hpp:256-263)."""
    import bisect

    return bisect.bisect_left(_W11, value % 2310) + 480 * (value // 2310)


def divisor_free_batches(abi, n, p, world):
    """Batches every rank can sweep from the top of its slice before any rank reaches the batch that holds p."""
    bp = (_index_of(p) - 2) // abi.BIGGEST_WHEEL
    _, _, node_range, batch_count = abi.node_split(n, world)
    avail = node_range
    for r in range(world):
        lo, hi = abi.node_batches(batch_count, node_range, r)
        if lo <= bp < hi:
            avail = min(avail, hi - bp - 1)
    return avail


def flag_check_number():
    """(N, p) for the cross-GPU found-flag self-check: 112-bit, smaller factor p at 97 % of sqrt(N) -- inside slice 0 at every
    nodeCount (rank 0 only signals), so no sweeping rank meets a divisor.  With the factor at 75 % (round 2's first version,
    never run above 2 GPUs) rank 1 of 4 found it within four launches and its own found flag stopped the others before rank 0
    signalled: the bench failed at 4 and 8 GPUs (tests/test_multi_gloo.py checks the placement for every nodeCount now)."""
    pp = _prev_prime(int(2 ** 55 * 1.9))
    return pp * _prev_prime(pp * 10628 // 10000), pp


def fixed_batches_per_step(abi, n, p, steps, warmup, world):
    """Batches per step, the SAME at every N: the largest multiple of 64 (at most 4096) such that warm-up + timed steps stay
    inside the divisor-free top of every rank's slice AND inside the top half of the shortest slice at nodeCount 1, 2, 4 and 8
    (and at this run's own world size) without repeating a batch -- so a 1 -> 8 GPU comparison measures scaling, not a change
    of launch size.  The top half of a slice is its first octave of guess values: every rank, like the single GPU of the N = 1
    run, is timed on the octave below the top of its own slice.  The bottom half of the LAST slice is every remaining octave of
    the range down to the largest wheel prime -- short chains, many small launches, a regime of its own --
    and is not left out of the record: `whole_range_sweep` in the same JSON line times every rank over its complete slice."""
    worlds = sorted({1, 2, 4, 8, world})
    free = min(divisor_free_batches(abi, n, p, w) for w in worlds)
    shortest = min(abi.node_split(n, w)[2] for w in worlds)
    per_step = min(free, shortest // 2) // max(1, steps + warmup)
    return max(1, min(4096, per_step // 64 * 64 if per_step >= 64 else per_step))


def whole_range_sweep(ctx, abi, torch, dist, bits, level, world, rank, kernel, chunk=4096):
    """Every rank sweeps its COMPLETE nodeCount = world slice of a number with no divisor below sqrt(N) (a prime of `bits`
    bits), top to bottom, all octaves: the job's wall time is the slowest rank's (the last slice holds the bottom of the
    range).  Outside the timed region; host clock around the calls, the kernels' device time beside it."""
    n_full = _prev_prime((1 << bits) - 12345)
    _, _, node_range, batch_count = abi.node_split(n_full, world)
    lo, hi = abi.node_batches(batch_count, node_range, rank)
    ctx.sweep(n_full, level, max(lo, hi - chunk), hi, batches_per_launch=chunk, flags=kernel)  # warm-up: tables, first launch
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    guesses = launches = 0
    dev_s = 0.0
    b = hi
    while b > lo:
        a = max(lo, b - chunk)
        res = ctx.sweep(n_full, level, a, b, batches_per_launch=chunk, flags=kernel)
        if res.found or res.aborted:
            raise SystemExit(f"bench.py: whole-range sweep of a prime reported found={res.found} aborted={res.aborted} on rank {rank}")
        guesses += res.guesses_tested
        launches += res.kernel_launches
        dev_s += res.device_seconds
        b = a
    torch.cuda.synchronize()
    mine_s = time.perf_counter() - t0
    by_rank, by_rank_dev, total = [mine_s], [dev_s], float(guesses)
    if dist is not None:
        t = torch.tensor([mine_s, dev_s, float(guesses)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        by_rank = [float(x[0].item()) for x in allr]
        by_rank_dev = [float(x[1].item()) for x in allr]
        total = sum(float(x[2].item()) for x in allr)
    job_s = max(by_rank)
    return {"n": str(n_full), "bits": bits, "level": level, "n_gpus": world, "batches": batch_count, "batches_per_call": chunk,
            "guesses": int(total), "seconds": job_s, "guesses_per_s": total / job_s,
            "seconds_by_rank": [round(x, 4) for x in by_rank], "device_seconds_by_rank": [round(x, 4) for x in by_rank_dev],
            "what": "every rank sweeps its whole nodeCount slice, top to bottom, every octave; seconds = the slowest rank "
                    "(host clock), i.e. the time the whole range of this number takes on n_gpus GPUs"}


# ---------------------------------------------------------------------------------------------------
# kernel identity: the profiler figures of profiles/current_kernels.json apply only to the binary they were captured from
# ---------------------------------------------------------------------------------------------------
def kernel_sass_sha256(lib_path, mangled):
    """sha256 of the kernel's SASS as `cuobjdump -sass -fun` prints it (None when cuobjdump or the kernel is missing)."""
    import hashlib

    try:
        out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, lib_path], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    i = out.find("Function : " + mangled)
    if i < 0:
        return None
    body = "\n".join(ln.strip() for ln in out[i:].splitlines() if ln.strip())
    return hashlib.sha256(body.encode()).hexdigest()


def kernel_profile(lib_path, key, mangled):
    """-> (entry or None, status).  The entry of profiles/current_kernels.json for `key`, only if its SASS hash is the
    loaded library's."""
    try:
        entry = json.load(open(KERNEL_PROFILES)).get(key)
    except (OSError, ValueError):
        return None, "profiles/current_kernels.json missing or unreadable"
    if entry is None:
        return None, f"no profiler capture recorded for {key}"
    have = kernel_sass_sha256(lib_path, mangled)
    if have is None:
        return None, "cuobjdump unavailable: the kernel binary could not be identified, profiler figures withheld"
    if have != entry.get("sass_sha256"):
        return None, f"stale capture: {key} in the loaded library differs from the profiled binary (commit {entry.get('commit')}); figures withheld"
    return entry, "SASS of the loaded kernel = SASS of the profiled kernel"


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ---------------------------------------------------------------------------------------------------
# work per guess for the roofline: A = Qn*(1+2*Lg)+2 IMAD-equivalents
# ---------------------------------------------------------------------------------------------------
def imad_eq_per_guess(bits):
    gbits = (bits + 1) // 2
    lg = (gbits + 31) // 32
    cmax_bits = bits - gbits + 1
    qn = (cmax_bits - 1 + 31) // 32
    return qn * (1 + 2 * lg) + 2


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------------
def _ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "qimcifa_ref")
    return p if os.path.exists(p) else None


def placed_semiprime_top_batch(bits, seed, frac=0.5):
    """A semiprime whose smaller factor sits `frac` of the way into the TOP batch of the single-node range, so
    that the unmodified reference (T threads) sweeps exactly the T top batches, one per thread, and exits."""
    from oracle import oracle_c as oc

    h = bits // 2
    q = _prev_prime((1 << h) - (_splitmix(seed) % (1 << (h - 4))))
    # fixed point: p just below sqrt(N) ~ p -> choose p from q so that p < q and p lies in the top batch
    p = _prev_prime(q - (1 << (h // 2)))
    for _ in range(8):
        n = p * q
        _, _, _, bc = oc.node_split(n, 1)
        lo_val = oc.forward((bc - 1) * oc.BW + 2)
        target = lo_val + int(frac * (oc.isqrt(n) - lo_val))
        p2 = _prev_prime(target)
        if p2 == p:
            break
        p = p2
    return p, q, p * q


def cpu_reference_step(bits, level, seed, threads):
    """One bounded sample of the workload on the CPU.  -> (guesses, seconds, kind, sample description)."""
    from oracle import oracle_c as oc

    exe = _ref_binary()
    p, q, n = placed_semiprime_top_batch(bits, seed)
    if exe:
        with tempfile.NamedTemporaryFile(suffix=".cnt", delete=False) as tf:
            cpath = tf.name
        env = dict(os.environ, QCF_SHIM_COUNT=cpath, QCF_REF_THREADS=str(threads))
        t0 = time.perf_counter()
        out = subprocess.run([exe], input=f"{n}\n1\n{level}\n", capture_output=True, text=True, env=env)
        dt = time.perf_counter() - t0
        guesses = int(open(cpath).read().strip() or 0)
        os.unlink(cpath)
        if f"{p} * {q}" not in out.stdout and f"{q} * {p}" not in out.stdout:
            raise RuntimeError("reference did not report the planted factor: " + out.stdout[-300:])
        return guesses, dt, "reference", (
            f"unmodified reference sources + compat shims (oracle/_ref), {bits}-bit semiprime with the factor placed mid-way "
            f"in the top batch, L{level}, {threads} threads: each thread sweeps one whole batch; guesses counted by the shim; "
            f"wall time of the process (includes wheel_gen setup and thread join)")
    _, _, _, bc = oc.node_split(n, 1)
    fac, tested, sec = oc.mt_sweep(n, level, bc - threads, bc, threads)
    return tested, sec, "port", (
        f"oracle port (oracle/oracle.cpp orc_mt_sweep, unsigned __int128), {bits}-bit semiprime, L{level}, top {threads} "
        f"batches, {threads} threads")


def cpu_reference_time_to_factor(n, level, threads):
    """Runs the unmodified reference (oracle/_ref) on N and returns its own "(Time elapsed: X seconds)"."""
    import re

    exe = _ref_binary()
    if not exe:
        return None
    env = dict(os.environ, QCF_REF_THREADS=str(threads))
    try:
        out = subprocess.run([exe], input=f"{n}\n1\n{level}\n", capture_output=True, text=True, env=env, timeout=120)
    except subprocess.TimeoutExpired:
        return None
    m = re.search(r"Time elapsed: ([0-9.eE+-]+) seconds", out.stdout)
    f = re.search(r"Found (\d+) \* (\d+) =", out.stdout)
    if not m or not f:
        return None
    return {"seconds": float(m.group(1)), "level": level, "threads": threads, "kind": "reference",
            "factors": [f.group(1), f.group(2)], "clock": "the reference's own, from after its dialog to printSuccess"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    level = args.level if args.level else 5
    # warm-up: tiny runs (page in the binary); the timed steps are whole-batch samples
    exe = _ref_binary()
    for _ in range(args.warmup):
        if exe:
            subprocess.run([exe], input="1000036000099\n1\n5\n", capture_output=True, text=True)
    tot_g, tot_s, kind, sample = 0, 0.0, None, None
    for k in range(args.steps):
        g, s, kind, sample = cpu_reference_step(args.bits, level, args.seed + 100 + k, threads)
        tot_g += g
        tot_s += s
    value = tot_g / tot_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32-limb integer", "data": "synthetic",
        "config": {"workload": f"{args.bits}-bit random semiprime, wheel level {level}, reference CPU threads on host cores",
                   "bits": args.bits, "level": level, "seed": args.seed},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# time-to-factor (the other half of BASELINE.json's metric): a real factoring run through the C ABI
# ---------------------------------------------------------------------------------------------------
def random_semiprime(bits, seed, balanced_log2=None):
    """p < q primes of bits/2 bits, top bit set.  balanced_log2=k places q/p - 1 <= 2^-k (config 3)."""
    h = bits // 2
    q = _prev_prime((1 << h) - (_splitmix(seed) % (1 << (h - 2))))
    if balanced_log2 is None:
        p = _prev_prime((1 << (h - 1)) + (_splitmix(seed + 7) % (1 << (h - 1))))
        if p >= q:
            p, q = _prev_prime(q - 2), q
    else:
        p = _prev_prime(q - (_splitmix(seed + 7) % (q >> balanced_log2)) - 2)
    return p, q, p * q


def time_to_factor(ctx, abi, n, level, world, rank, chunk, kernel, dist=None, flag=None, max_seconds=60.0):
    """Sweeps the rank's node slice from the top in runs of `chunk` batches until some rank finds a divisor
    (found flag all-reduced after every run).  -> dict (qimcifa_b200/multi.py)."""
    from qimcifa_b200 import multi

    lo, hi = multi.rank_slice(n, world, rank)
    return multi.sweep_slice_until_found(
        lambda b_lo, b_hi: ctx.sweep(n, level, b_lo, b_hi, batches_per_launch=chunk, flags=kernel),
        lo, hi, chunk, dist=dist, flag=flag, max_seconds=max_seconds)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_graft_arm(args):
    import torch
    import torch.distributed as dist

    from qimcifa_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; qimcifa_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    p, q, n = synthetic_semiprime(args.bits, args.seed)
    ctx = abi.Context(local_rank, stream=torch.cuda.current_stream().cuda_stream)
    info = ctx.device_info()

    if args.mode == "random":
        if args.bits == 96:
            args.bits = 128  # BASELINE config 5 is quoted on a 128-bit semiprime
        return run_random_mode(args, ctx, abi, torch, dist if world > 1 else None, world, rank, info)

    slice_world, slice_rank = (args.slice_of[1], args.slice_of[0]) if args.slice_of else (world, rank)
    _, _, node_range, batch_count = abi.node_split(n, slice_world)
    lo, hi = abi.node_batches(batch_count, node_range, slice_rank)
    free = divisor_free_batches(abi, n, p, slice_world)
    if not args.batches:
        args.batches = fixed_batches_per_step(abi, n, p, args.steps, args.warmup, slice_world)
    # level: the GPU tuner's choice (cost per guess x cardinality, src/qimcifa_tuner.cpp:137-151) unless given
    tuner = None
    if args.level:
        level = args.level
    else:
        tuner = {}
        for lv in range(5, 10):
            # time the launch shape the sweep will use (the filter kernel's plan depends on the interval length)
            # (best of three: the first launch of a kernel instantiation pays its lazy module load inside the events)
            best = None
            for _ in range(3):
                sec, g = ctx.time_level(n, lv, args.batches, args.kernel)
                best = sec / g if best is None else min(best, sec / g)
            tuner[lv] = best * abi.count_guesses(lv, 0, abi.node_split(n, 1)[3])
        # the levels cost the same to within the timing noise on the GPU (superset chains: a higher level sweeps the
        # same chains and discards more survivors): a higher level is taken only when it is at least 3 % cheaper,
        # otherwise run-to-run noise would pick the level -- and with it the number of guesses a batch counts -- at random
        level = 5
        for lv in range(6, 10):
            if tuner[lv] < 0.97 * tuner[level]:
                level = lv
        if world > 1:
            t = torch.tensor([level], device="cuda")
            dist.broadcast(t, 0)
            level = int(t.item())

    # node split: rank k = node k of nodeCount = world (intent semantics)
    if hi - lo < args.batches:
        raise SystemExit(f"slice of {hi - lo} batches is shorter than one step of {args.batches}; lower --batches")
    if free < args.batches:
        raise SystemExit(f"only {free} divisor-free batches at the top of the slices; lower --batches")
    span = free - args.batches + 1  # start offsets available; steps wrap around the divisor-free part if it is short
    wrapped = args.batches * (args.steps + args.warmup) > free
    found_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    # Found-flag exchange, the only cross-GPU traffic of the path.  Default: peer memory -- every rank maps the other
    # ranks' flag words (CUDA IPC) and a sweep that finds a divisor writes them with system-scope atomics over NVLink,
    # so nothing is exchanged while nothing is found.  Fallback / --flag nccl: a 1-word all_reduce(MAX) after every step.
    from qimcifa_b200 import multi

    peer_flag = False
    if world > 1 and args.flag == "peer":
        peer_flag = multi.connect_peers(ctx, dist, rank, world, "cuda")

    def step(k):
        b_hi = hi - (k * args.batches) % span
        res = ctx.sweep(n, level, b_hi - args.batches, b_hi, batches_per_launch=args.batches, flags=args.kernel)
        if world > 1 and not peer_flag:
            found_flag[0] = res.found
            dist.all_reduce(found_flag, op=dist.ReduceOp.MAX)
        return res

    for k in range(args.warmup):
        step(k)

    # One UNTIMED step with the guesses counted on the device (the kernel's own per-chain tally, QCF_COUNT_ON_DEVICE): the
    # numerator of `value` is the host's closed form, and this asserts that the device walked exactly that many guesses
    # for the launch shape the timed steps use (reference include/qimcifa.hpp:369: every guess of the batch is decided).
    b_chk = hi - ((args.warmup + args.steps) * args.batches) % span
    res_chk = ctx.sweep(n, level, b_chk - args.batches, b_chk, batches_per_launch=args.batches, flags=args.kernel | abi.COUNT_ON_DEVICE)
    plan_chk = ctx.last_plan()
    if res_chk.guesses_counted != res_chk.guesses_tested or res_chk.found or res_chk.aborted:
        raise SystemExit(f"bench.py: device-side guess count {res_chk.guesses_counted} != closed form {res_chk.guesses_tested} "
                         f"(found {res_chk.found}, aborted {res_chk.aborted}) on rank {rank}")
    device_count_check = {"guesses_counted_on_device": res_chk.guesses_counted, "guesses_closed_form": res_chk.guesses_tested,
                          "equal": True, "batches": args.batches, "where": "one untimed step of the timed launch shape, every rank"}

    # integer-pipe peak on this box, same process, before the timed region (the roofline denominator)
    imad_peak = max(ctx.microbench(0, 1 << 13)[0] for _ in range(3))
    wide_peak = max(ctx.microbench(1, 1 << 13)[0] for _ in range(3))
    iadd_peak = max(ctx.microbench(2, 1 << 13)[0] for _ in range(3))
    try:  # the high-half multiply-add of the chain kernel's two-stage loop (added after the round's last full GPU run)
        hi_peak = max(ctx.microbench(3, 1 << 13)[0] for _ in range(3))
    except Exception:
        hi_peak = None

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)

    # ---- timed region 1: device-resident (CUDA events on the launching stream) ----
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    guesses = launches = 0
    kernel_s = 0.0
    t_wall0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        res = step(args.warmup + k)
        assert not res.found and not res.aborted and res.batches_done == args.batches
        guesses += res.guesses_tested
        launches += res.kernel_launches
        kernel_s += res.device_seconds
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    clocks = sampler.stop() if rank == 0 else None

    per_rank = per_rank_kernel = None
    if world > 1:
        mine = torch.tensor([dev_s * 1e3 / args.steps, kernel_s * 1e3 / args.steps], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [round(float(x[0].item()), 3) for x in allr]
        per_rank_kernel = [round(float(x[1].item()), 3) for x in allr]
    stats = torch.tensor([dev_s, t_wall, kernel_s, float(guesses), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_s, t_wall, kernel_s = mx[0].item(), mx[1].item(), mx[2].item()
        total_guesses, total_launches = sm[3].item(), int(sm[4].item())
    else:
        total_guesses, total_launches = float(guesses), launches

    # value: whole-job guesses/s with N resident; every step's qcf_sweep is one launch whose only input is the
    # kernel parameter block, so the device-event clock over the K steps is the device-resident number.
    value = total_guesses / dev_s
    # e2e: the same K steps through the C ABI with HOST buffers (N limbs in, qcf_result + hit list out),
    # host clock around the calls: includes parameter upload, launch, the D2H read of the hit record, sync.
    e2e_value = total_guesses / t_wall

    # ---- common-factor mode (the reference's IS_RSA_SEMIPRIME=OFF build): throughput near sqrt(N), outside the timed region ----
    gcd_mode = None
    if (not args.no_ttf) and (world == 1):
        try:
            _, _, _, bc_g = abi.node_split(n, 1)
            ctx.sweep(n, level, bc_g - 72, bc_g - 8, batches_per_launch=64, flags=abi.MODE_GCD)
            rg = ctx.sweep(n, level, bc_g - 72, bc_g - 8, batches_per_launch=64, flags=abi.MODE_GCD)
            gcd_mode = {"guesses_per_s": rg.guesses_tested / rg.device_seconds, "batches": 64, "level": level,
                        "kernel": "qcf_gcd_chain_kernel (Montgomery product of 1024 guesses per gcd, two accumulators per thread)"}
        except Exception as e:  # a side leg must never take the bench line down
            gcd_mode = {"error": str(e)}

    # ---- the exact per-guess kernel (what BASELINE.md's IMAD roofline prices), both loops of its chain form, outside the
    # timed region: option xchain_two_stage = 0 is the full reduction per guess, 1 the low limb first ----
    exact_mode = None
    if (not args.no_ttf) and (world == 1):
        try:
            _, _, _, bc_x = abi.node_split(n, 1)
            exact_mode = {"batches": 256, "level": level, "kernel": "qcf_exact_chain_kernel", "unit": UNIT,
                          "imad_eq_per_guess": imad_eq_per_guess(args.bits)}
            for two, key in (("0", "one_stage"), ("1", "two_stage")):
                abi.set_option("xchain_two_stage", two)
                best = 0.0
                for _ in range(3):
                    rx = ctx.sweep(n, level, bc_x - 264, bc_x - 8, batches_per_launch=256, flags=abi.KERNEL_EXACT)
                    best = max(best, rx.guesses_tested / rx.device_seconds)
                exact_mode[key] = best
        except Exception as e:  # a side leg must never take the bench line down
            exact_mode = {"error": str(e)}
        finally:
            abi.set_option("xchain_two_stage", None)

    # ---- time-to-factor legs (outside the timed region) ----
    ttf_out = []
    if not args.no_ttf:
        if world == 1:
            # config 1: 64-bit semiprimes at the reference's default level 7 (src/qimcifa.cpp:90) and at the GPU tuner's
            # level; config 2: one 96-bit semiprime at the tuner's level.  Host clock around the whole factoring run.
            for bits_t, lv, seeds in ((64, 7, (1001, 1011, 1021)), (64, level, (1001, 1011, 1021)), (96, level, (1002,))):
                for sd in seeds:
                    p_t, q_t, n_t = random_semiprime(bits_t, sd)
                    chunk = 1 << 20 if bits_t == 64 else 4096
                    time_to_factor(ctx, abi, n_t, lv, 1, 0, chunk, args.kernel, max_seconds=0.0)  # warm-up: tables, first launch
                    r = time_to_factor(ctx, abi, n_t, lv, 1, 0, chunk, args.kernel)
                    ok = r["factor"] in (p_t, q_t)
                    ttf_out.append({"bits": bits_t, "level": lv, "n": str(n_t), "seed": sd, "seconds": r["seconds"],
                                    "guesses": r["guesses"], "correct": ok, "n_gpus": 1})
        else:
            # config 3: 112-bit balanced semiprime, nodeCount = world, found-flag early exit
            p_t, q_t, n_t = random_semiprime(112, 1003, balanced_log2=12)
            if peer_flag:
                lo_t, hi_t = multi.rank_slice(n_t, world, rank)
                dist.barrier()
                r = multi.sweep_slice_with_peer_flag(
                    lambda a, b: ctx.sweep(n_t, level, a, b, batches_per_launch=4096, flags=args.kernel), lo_t, hi_t, 4096,
                    max_seconds=60.0)
                torch.cuda.synchronize()
                dist.barrier()  # every rank has stopped: that is the time to factor
                tt = torch.tensor([r["seconds"]], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                r["seconds"] = float(tt.item())
                ctx.set_found(0)
            else:
                r = time_to_factor(ctx, abi, n_t, level, world, rank, 4096, args.kernel, dist=dist, flag=found_flag)
            g_all = torch.tensor([float(r["guesses"])], dtype=torch.float64, device="cuda")
            dist.all_reduce(g_all)
            f_all = torch.tensor([float(r["factor"] or 0) if r["factor"] in (p_t, q_t) else 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(f_all, op=dist.ReduceOp.MAX)
            ttf_out.append({"bits": 112, "level": level, "n": str(n_t), "seed": 1003, "balanced": "q/p-1 <= 2^-12",
                            "seconds": r["seconds"], "guesses": int(g_all.item()), "correct": f_all.item() > 0, "n_gpus": world,
                            "split": "nodeCount = n_gpus (rank k = node k)",
                            "note": "literal reference never finds it: nodes 0..3 of 8 sweep nothing (SURVEY.md D1)"})
            if peer_flag:
                # the same number as ONE node whose batches all GPUs pull from a shared descending counter (the
                # reference's thread model, include/qimcifa.hpp:107-129): a fetch-add on a word in rank 0's memory,
                # issued from each GPU over NVLink peer memory -- every GPU works at the top of the range
                ctx.set_found(0)
                if rank == 0:
                    ctx.dispenser_reset(0)
                lo_s, hi_s = multi.rank_slice(n_t, 1, 0)
                torch.cuda.synchronize()
                dist.barrier()
                r2 = multi.sweep_shared_dispenser(
                    lambda a, b: ctx.sweep(n_t, level, a, b, batches_per_launch=512, flags=args.kernel),
                    lambda take: ctx.peer_dispense(0, take), lo_s, hi_s, 512, max_seconds=60.0)
                torch.cuda.synchronize()
                dist.barrier()
                tt = torch.tensor([r2["seconds"]], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                g2 = torch.tensor([float(r2["guesses"])], dtype=torch.float64, device="cuda")
                dist.all_reduce(g2)
                f2 = torch.tensor([float(r2["factor"] or 0) if r2["factor"] in (p_t, q_t) else 0.0], dtype=torch.float64, device="cuda")
                dist.all_reduce(f2, op=dist.ReduceOp.MAX)
                ctx.set_found(0)
                ttf_out.append({"bits": 112, "level": level, "n": str(n_t), "seed": 1003, "balanced": "q/p-1 <= 2^-12",
                                "seconds": float(tt.item()), "guesses": int(g2.item()), "correct": f2.item() > 0, "n_gpus": world,
                                "split": "one node, shared dispenser over peer memory (runs of 512 batches)"})
    # ---- the whole range of a number of the same size, every rank over its complete slice (outside the timed region) ----
    whole = None
    if (not args.no_ttf) and (args.bits <= 96):
        whole = whole_range_sweep(ctx, abi, torch, dist if world > 1 else None, args.bits, level, world, rank, args.kernel)

    # ---- cross-GPU self-checks (world >= 2), outside the timed region: the peer-memory found flag stops RUNNING sweeps on the
    # other GPUs, and the shared dispenser hands every run of batches out exactly once (qimcifa_b200/multi.py; the same
    # functions tests/test_gpu_peer.py runs on 2 GPUs).  A failure fails the bench run. ----
    multi_checks = None
    if world > 1 and peer_flag and not args.no_ttf:
        n_flag = flag_check_number()[0]
        chk_flag = multi.check_peer_flag(ctx, dist, rank, world, n_flag)
        q_d = _prev_prime(2 ** 50 - 777)
        p_d = _prev_prime(q_d - 1400 * abi.BIGGEST_WHEEL * 2310 // 480)  # ~700 batches below sqrt(N) of a 100-bit semiprime
        chk_disp = multi.check_shared_dispenser(ctx, dist, rank, world, p_d * q_d, p_d, chunk=128)
        multi_checks = {"peer_flag_stops_running_sweeps": chk_flag, "shared_dispenser_hands_out_every_run_once": chk_disp}
        if not (chk_flag["ok"] and chk_disp["ok"]):
            raise SystemExit("bench.py: cross-GPU self-check failed: " + json.dumps(multi_checks))
    for t in ttf_out:
        if not t["correct"]:
            raise SystemExit("bench.py: a time-to-factor leg reported a wrong factor: " + json.dumps(t))

    if rank == 0:
        # roofline of the step's kernels (SURVEY.md 8d accounting): A IMAD-equivalents per guess x guesses / kernel time
        a_eq = imad_eq_per_guess(args.bits)
        achieved = guesses * a_eq / kernel_s
        cpu = None
        if world == 1 and not args.no_cpu:
            g, s, kind, sample = cpu_reference_step(args.bits, 5, args.seed + 100, os.cpu_count() or 1)
            cpu = {"value": g / s, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind, "sample": sample}
            # time-to-factor beside it: measured for the first 64-bit instance (at the CPU's better level 5),
            # extrapolated from the CPU guesses/s for the rest -- labelled so
            done64 = False
            for t in ttf_out:
                if t["bits"] == 64 and not done64:
                    t["cpu_reference"] = cpu_reference_time_to_factor(int(t["n"]), 5, os.cpu_count() or 1)
                    done64 = True
                else:
                    t["cpu_reference_extrapolated_seconds"] = t["guesses"] / cpu["value"]
        # ---- roofline of the dominant kernel: the SM issue port (one warp instruction per sub-partition and clock) ----
        # achieved = warp instructions issued per second by the step's kernel = guesses/s (this run, CUDA events) x warp
        # instructions per guess (ncu smsp__inst_executed.sum / guesses of the profiled launch of the SAME binary, see
        # kernel_profile); peak = SMs x 4 sub-partitions x the SM clock sampled during the timed region.  BASELINE.md's
        # A-accounting (what an exact per-guess test would cost on the IMAD pipe) is reported beside it under its own name.
        f_sm = (clocks.get("sm_mhz") or 1965.0) * 1e6
        issue_peak = info["sm_count"] * 4 * f_sm
        ln_n = max(2, (n.bit_length() + 31) // 32)
        if res.kernel_kind == 2 and plan_chk is not None:
            kkey = f"qcf_filter_kernel<{plan_chk.degree},{ln_n},2,{'true' if plan_chk.packed else 'false'}>"
            kmangled = f"_Z17qcf_filter_kernelILi{plan_chk.degree}ELi{ln_n}ELi2ELb{plan_chk.packed}EEv12FilterParams"
        else:
            kkey = f"qcf_exact_chain_kernel<{ln_n},2,2,true,true>"
            kmangled = f"_Z22qcf_exact_chain_kernelILi{ln_n}ELi2ELi2ELb1ELb1EEv12FilterParams"
        prof, prof_status = kernel_profile(abi.LIB_PATH, kkey, kmangled)
        rate_per_gpu = (guesses / kernel_s) if kernel_s > 0 else 0.0  # this rank's kernel-only rate
        wipg = prof.get("warp_inst_per_guess") if prof else None
        # The packed kernels are bound by the integer ALU pipe (16 lanes per sub-partition: one warp instruction per two clocks),
        # not by the issue port: when the pinned profile carries the pipe's busy SM-cycles per guess, the roofline is that pipe:
        #   achieved = guesses/s (live) x ALU-busy SM-cycles per guess (ncu, same binary) = busy SM-cycles per second,
        #   peak = SMs x the SM clock sampled live (the pipe can be busy every cycle).  The issue-port fraction stays beside it.
        alu_cpg = prof.get("alu_pipe_sm_cycles_per_guess") if prof else None
        issue_frac = (rate_per_gpu * wipg / issue_peak) if wipg else None
        alu_frac = (rate_per_gpu * alu_cpg / (info["sm_count"] * f_sm)) if alu_cpg else None
        alu_bound = alu_frac is not None and (issue_frac is None or alu_frac > issue_frac)
        roofline = {
            "bound": "alu_pipe" if alu_bound else "sm_issue", "unit": "busy SM-cycles/s" if alu_bound else "warp-inst/s",
            "achieved": (rate_per_gpu * alu_cpg) if alu_bound else (rate_per_gpu * wipg if wipg else None),
            "peak": (info["sm_count"] * f_sm) if alu_bound else issue_peak,
            "frac": alu_frac if alu_bound else issue_frac,
            "issue_slot_frac": issue_frac, "alu_pipe_frac": alu_frac,
            "traffic": prof.get("dram_bytes_per_launch") if prof else None,
            "kernel": kkey,
            "warp_inst_per_guess": wipg,
            "thread_inst_per_guess": prof.get("thread_inst_per_guess") if prof else None,
            "profile": {"status": prof_status, "file": (prof or {}).get("source"), "commit": (prof or {}).get("commit"),
                        "sass_sha256": (prof or {}).get("sass_sha256"),
                        "ncu": {k: prof[k] for k in ("issue_active_pct", "pipe_alu_pct", "pipe_alu_cycles_active_pct", "pipe_fma_pct", "avg_threads_per_inst", "guesses_of_profiled_launch",
                                                     "warp_inst_of_profiled_launch", "profiled_launch") if prof and k in prof}},
            "peak_how": (f"{info['sm_count']} SMs x {f_sm / 1e6:.0f} MHz: ALU-pipe busy cycles per SM and second" if alu_bound else
                         f"{info['sm_count']} SMs x 4 sub-partitions x {f_sm / 1e6:.0f} MHz") + " (median SM clock sampled by nvidia-smi during the timed region)",
            "why_not_hbm_or_tensor": "integer ALU path: 0 algorithmic HBM bytes per guess (tables staged once per block into shared "
                                     "memory), no contraction; the binding resource is the integer ALU pipe / the issue port",
            "guesses_per_sm_cycle": rate_per_gpu / (info["sm_count"] * f_sm),
            "plan": None if plan_chk is None else {"degree": plan_chk.degree, "tprime": plan_chk.tprime, "steps": plan_chk.steps,
                                                   "segments": plan_chk.n_seg, "filter_bits": plan_chk.fbits},
            # BASELINE.md section 3 accounting, kept under its own name: NOT a fraction of a peak for the filter kernel
            "imad_accounting": {
                "work_per_guess_imad_eq": a_eq, "imad_peak_per_s": imad_peak,
                "algorithmic_speedup_vs_A": achieved / imad_peak,
                "note": "A = Qn(1+2Lg)+2 IMAD-equivalents per guess prices an exact per-guess test (BASELINE.md sec. 3).  The filter "
                        "kernel is the disclosed cheaper scheme of SURVEY.md 8d (2-adic quotient stepped by finite differences, exact "
                        "test of survivors only): this ratio says how much work the scheme avoids, not how busy the hardware is.  "
                        "exact_kernel.* below is the kernel A does price."},
            "peaks_measured": {"imad_per_s": imad_peak, "imad_wide_per_s": wide_peak, "iadd_per_s": iadd_peak, "imad_hi_per_s": hi_peak,
                               "how": "qcf_microbench: 32 independent chains/thread, 256 thr x 8 blocks/SM, same process"},
            "hbm_bytes_per_guess": 0,
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32-limb integer", "data": "synthetic",
            "config": {
                "workload": f"{args.bits}-bit random semiprime (two {args.bits // 2}-bit primes), "
                            f"{'tuner-selected' if tuner else 'fixed'} wheel level {level}, nodeCount={world} split, "
                            f"{args.batches} batches (x223092870 indices) per step",
                "bits": args.bits, "level": level, "seed": args.seed, "n": str(n), "batches_per_step": args.batches,
                "guesses_per_step_per_gpu": guesses // args.steps, "kernel": {0: "auto", 1: "exact", 2: "filter"}[args.kernel],
                "l2": "not applicable: the kernel reads <= 50 KB of wheel tables into shared memory once per block and touches "
                      "no other global memory; successive steps sweep different batches",
                "tuner_cost_s": tuner,
                "steps_wrap_around_slice": wrapped,
                "steps_cover": "warm-up + timed steps walk down from the top of each rank's slice and stay inside the top half "
                               "(the first octave of guess values) of the shortest slice of a 1/2/4/8-way split; the complete "
                               "slices, bottom octaves included, are timed under whole_range_sweep",
                "ms_per_step_by_rank": per_rank,
                "kernel_ms_per_step_by_rank": per_rank_kernel,
                "found_flag_exchange": None if world == 1 else ("peer-memory atomics over NVLink (CUDA IPC)" if peer_flag
                                                                 else "NCCL all_reduce(MAX) of one word per step"),
                "slice": f"node {slice_rank} of {slice_world}" if args.slice_of else None,
            },
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": 16 + abi.io_bytes()[0] * int(round(total_launches / max(1, world * args.steps))),
                    "d2h_bytes_per_step": abi.io_bytes()[2],
                    "note": "qcf_sweep() per step with host buffers, host clock: N limbs + the kernel parameter blocks up (every "
                            "launch counted at the chain kernels' parameter-block size, qcf_io_bytes), the hit record down"},
            "gpu_launches": total_launches,
            "time_to_factor": ttf_out,
            "whole_range_sweep": whole,
            "common_factor_mode": gcd_mode,
            "exact_kernel": exact_mode,
            "roofline": roofline,
            "device_count_check": device_count_check,
            "multi_gpu_checks": multi_checks,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "device": info["name"],
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_random_mode(args, ctx, abi, torch, dist, world, rank, info):
    """BASELINE config 5: random-guess Monte-Carlo mode on a 128-bit semiprime, fixed guess budget per step and GPU,
    Philox4x32-10 through the wheel.  --random-map chains (default): Philox draws row groups of chains of a random window and
    the filter kernel walks them (QCF_RANDOM_CHAINS, include/qimcifa_cuda.h); independent: one Philox draw, one table lookup
    and one exact test per guess.  Rank r takes the counter blocks r, r + world, ...; nothing is exchanged between ranks but
    the found flag (one word all-reduced per step)."""
    bits = args.bits
    p, q, n = random_semiprime(bits, 1005)
    level = args.level or 5
    chains = args.random_map == "chains"
    # a step = 16 counter blocks per GPU (chain draws: 2^38 guess numbers, 16 randomly drawn windows, ~16 ms): one block is
    # one window at a random depth below sqrt(N), a millisecond whose cost varies 3x with the depth -- too short and too
    # uneven a step for a max-over-ranks clock (the first 8-GPU run of round 2 timed 6 such steps)
    blocks_per_step = 16
    block = blocks_per_step << (34 if chains else 28)
    flags = abi.RANDOM_CHAINS if chains else 0
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")

    def step(k):
        res = ctx.sweep_random(n, level, args.seed, (k * world + rank) * block, block, flags=flags)
        if dist is not None:
            flag[0] = res.found
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        return res

    for k in range(args.warmup):
        step(k)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    guesses = launches = 0
    kernel_s = 0.0
    for k in range(args.steps):
        res = step(args.warmup + k)
        assert not res.aborted
        guesses += res.guesses_tested
        launches += res.kernel_launches
        kernel_s += res.device_seconds
    ev1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    plan = ctx.last_plan() if chains else None
    from qimcifa_b200 import multi

    (dev_s, wall), (tot_g, tot_l) = multi.reduce_max_sum(dist, [ev0.elapsed_time(ev1) * 1e-3, wall], [guesses, launches], "cuda")
    if rank == 0:
        roofline = None
        if chains and plan is not None:
            f_sm = (clocks.get("sm_mhz") or 1965.0) * 1e6
            ln_n = max(2, (n.bit_length() + 31) // 32)
            kkey = f"qcf_filter_kernel<{plan.degree},{ln_n},2,{'true' if plan.packed else 'false'}>"
            prof, status = kernel_profile(abi.LIB_PATH, kkey, f"_Z17qcf_filter_kernelILi{plan.degree}ELi{ln_n}ELi2ELb{plan.packed}EEv12FilterParams")
            wipg = prof.get("warp_inst_per_guess") if prof else None
            rate = guesses / kernel_s if kernel_s > 0 else 0.0
            peak = info["sm_count"] * 4 * f_sm
            alu_cpg = prof.get("alu_pipe_sm_cycles_per_guess") if prof else None  # the packed kernels are bound by the integer ALU pipe
            issue_frac = rate * wipg / peak if wipg else None
            alu_frac = rate * alu_cpg / (info["sm_count"] * f_sm) if alu_cpg else None
            alu_bound = alu_frac is not None and (issue_frac is None or alu_frac > issue_frac)
            roofline = {"bound": "alu_pipe" if alu_bound else "sm_issue", "unit": "busy SM-cycles/s" if alu_bound else "warp-inst/s",
                        "achieved": (rate * alu_cpg) if alu_bound else (rate * wipg if wipg else None),
                        "peak": (info["sm_count"] * f_sm) if alu_bound else peak,
                        "frac": alu_frac if alu_bound else issue_frac, "issue_slot_frac": issue_frac, "alu_pipe_frac": alu_frac,
                        "traffic": prof.get("dram_bytes_per_launch") if prof else None,
                        "kernel": kkey, "warp_inst_per_guess": wipg, "profile": {"status": status, "file": (prof or {}).get("source")},
                        "plan": {"degree": plan.degree, "tprime": plan.tprime, "steps": plan.steps, "segments": plan.n_seg},
                        "guesses_per_sm_cycle": rate / (info["sm_count"] * f_sm), "hbm_bytes_per_guess": 0}
        h2d, _, d2h = abi.io_bytes()
        print(json.dumps({
            "metric": METRIC, "value": tot_g / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32-limb integer", "data": "synthetic",
            "config": {"workload": f"random-guess Monte-Carlo mode ({'chain draws: Philox picks row groups of chains, the filter kernel walks them' if chains else 'independent draws: one exact test per guess'}), "
                                   f"{bits}-bit semiprime, wheel level {level}, Philox4x32-10 seed {args.seed}, {blocks_per_step} counter blocks of 2^{34 if chains else 28} guess numbers per step "
                                   f"per GPU, disjoint counter blocks per rank",
                       "bits": bits, "level": level, "n": str(n), "mode": "random", "map": args.random_map,
                       "l2": "not applicable: no global-memory inputs",
                       "note": "no reference code exists for this mode (SURVEY.md section 0, reference README.md:25 only describes it)"},
            "e2e": {"value": tot_g / wall, "unit": UNIT, "h2d_bytes_per_step": 16 + (h2d if chains else 160) * int(round(tot_l / max(1, world * args.steps))),
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(tot_l), "roofline": roofline, "cpu_baseline": None, "clocks": clocks, "device": info["name"],
        }), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--bits", type=int, default=96, help="semiprime size (BASELINE configs: 64, 96, 100, 112, 128)")
    ap.add_argument("--level", type=int, default=0, help="wheel level 5..9; 0 = tuner-selected")
    ap.add_argument("--batches", type=int, default=0,
                    help="reference batches per step per GPU (0 = auto: the largest multiple of 64, at most 4096, such that "
                         "warm-up + timed steps fit the divisor-free top of every rank's slice without repeating a batch)")
    ap.add_argument("--seed", type=int, default=1002)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 exact, 2 filter")
    ap.add_argument("--flag", default="peer", choices=["peer", "nccl"], help="cross-GPU found-flag exchange (N > 1)")
    ap.add_argument("--mode", default="sweep", choices=["sweep", "random"], help="sweep (headline) or random-guess Monte-Carlo mode")
    ap.add_argument("--random-map", default="chains", choices=["chains", "independent"],
                    help="--mode random: chain draws through the filter kernel (default) or one exact test per independently drawn guess")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ttf", action="store_true", help="skip the time-to-factor legs")
    ap.add_argument("--slice-of", type=int, nargs=2, metavar=("NODE", "COUNT"), default=None,
                    help="single-GPU diagnostic: sweep the slice node NODE of COUNT would own")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_graft_arm(args)


if __name__ == "__main__":
    sys.exit(main())
