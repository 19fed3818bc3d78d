"""One launch shape, repeated -- the program ncu captures for profiles/current_kernels.json.  No torch (starts in a second).

  python tools/profile_launch.py --kind filter --bits 96 --level 5 --batches 896 --depth 1 --reps 3
  python tools/profile_launch.py --kind exact  --bits 96 --level 5 --batches 256
  python tools/profile_launch.py --kind filter --force 3,0,0 --depth 64          (a degree-3 launch at sqrt(N)/64)
  python tools/profile_launch.py --kind gcd | random

Prints ONE JSON line: guesses per launch, device ms (best of --reps), the plan of the filter launch, the kernel's key and
mangled name and the sha256 of its SASS in the loaded library (bench.kernel_sass_sha256): tools/kernel_profile.py joins that
line with the ncu report of the same command into an entry of profiles/current_kernels.json.
  --depth D sweeps at sqrt(N)/D (the top of slice D-1 of D when D is a nodeCount)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qimcifa_b200 import abi

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="filter", choices=["filter", "exact", "exact1", "gcd", "random", "randomc"])
ap.add_argument("--bits", type=int, default=96)
ap.add_argument("--level", type=int, default=5)
ap.add_argument("--batches", type=int, default=0)
ap.add_argument("--depth", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--force", default=None, help="plan_force option, e.g. 3,0,0")
ap.add_argument("--option", action="append", default=[], help="name=value for qcf_set_option")
args = ap.parse_args()

p, q, n = bench.synthetic_semiprime(args.bits, 1002)
ln = max(2, (n.bit_length() + 31) // 32)
out = {"kind": args.kind, "bits": args.bits, "level": args.level, "depth": args.depth, "n": str(n)}
if args.force:
    abi.set_option("plan_force", args.force)
for o in args.option:
    k, v = o.split("=", 1)
    abi.set_option(k, v)
with abi.Context(0) as ctx:
    out["sm_count"] = ctx.device_info()["sm_count"]
    _, _, _, bc = abi.node_split(n, 1)
    if args.kind in ("random", "randomc"):
        budget = 1 << (34 if args.kind == "randomc" else 28)
        best = None
        for r in range(args.reps):
            res = ctx.sweep_random(n, args.level, 1002, r * budget, budget, flags=abi.RANDOM_CHAINS if args.kind == "randomc" else 0)
            best = res.device_seconds if best is None else min(best, res.device_seconds)
        out.update(guesses=res.guesses_tested, ms=best * 1e3, launches=res.kernel_launches, key="qcf_random", mangled=None)
        pl = ctx.last_plan() if args.kind == "randomc" else None
        if pl is not None:  # chain draws run the sweep's filter kernel on the plan of the drawn window
            out["plan"] = {"degree": pl.degree, "tprime": pl.tprime, "steps": pl.steps, "trips": pl.trips, "segments": pl.n_seg,
                           "filter_bits": pl.fbits, "period": pl.period, "cost": pl.cost, "packed": pl.packed}
            out["key"] = f"qcf_filter_kernel<{pl.degree},{ln},2,{'true' if pl.packed else 'false'}>"
            out["mangled"] = f"_Z17qcf_filter_kernelILi{pl.degree}ELi{ln}ELi2ELb{pl.packed}EEv12FilterParams"
    else:
        B = args.batches or (bench.fixed_batches_per_step(abi, n, p, 20, 5, 1) if args.kind == "filter" else 256)
        b_hi = int(bc / args.depth) - (8 if args.depth == 1 else 0)
        flags = {"filter": abi.KERNEL_AUTO, "exact": abi.KERNEL_EXACT, "exact1": abi.KERNEL_EXACT, "gcd": abi.MODE_GCD}[args.kind]
        if args.kind == "exact1":
            abi.set_option("xchain_two_stage", "0")
        if args.force:
            flags = abi.KERNEL_FILTER if args.kind == "filter" else flags
        best = None
        for r in range(args.reps):
            res = ctx.sweep(n, args.level, b_hi - B, b_hi, batches_per_launch=B, flags=flags)
            best = res.device_seconds if best is None else min(best, res.device_seconds)
        pl = ctx.last_plan()
        out.update(batches=B, b_hi=b_hi, guesses=res.guesses_tested, ms=best * 1e3, launches=res.kernel_launches,
                   kernel_kind=res.kernel_kind, found=res.found)
        if pl is not None:
            out["plan"] = {"degree": pl.degree, "tprime": pl.tprime, "steps": pl.steps, "trips": pl.trips, "segments": pl.n_seg,
                           "filter_bits": pl.fbits, "period": pl.period, "cost": pl.cost}
            out["plan"]["packed"] = pl.packed
            out["key"] = f"qcf_filter_kernel<{pl.degree},{ln},2,{'true' if pl.packed else 'false'}>"
            out["mangled"] = f"_Z17qcf_filter_kernelILi{pl.degree}ELi{ln}ELi2ELb{pl.packed}EEv12FilterParams"
        elif args.kind in ("exact", "exact1"):
            two = "true" if args.kind == "exact" else "false"
            out["key"] = f"qcf_exact_chain_kernel<{ln},2,2,true,{two}>"
            out["mangled"] = f"_Z22qcf_exact_chain_kernelILi{ln}ELi2ELi2ELb1ELb{1 if args.kind == 'exact' else 0}EEv12FilterParams"
        elif args.kind == "gcd":
            out["key"] = f"qcf_gcd_chain_kernel<{ln},2>"
            out["mangled"] = f"_Z20qcf_gcd_chain_kernelILi{ln}ELi2EEv12FilterParams"
    out["guesses_per_s"] = out["guesses"] / (out["ms"] * 1e-3)
    if out.get("mangled"):
        out["sass_sha256"] = bench.kernel_sass_sha256(abi.LIB_PATH, out["mangled"])
print(json.dumps(out), flush=True)
