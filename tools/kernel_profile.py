"""Joins an ncu report with the JSON line of tools/profile_launch.py (the same command, run without ncu first) into an entry
of profiles/current_kernels.json -- the ONLY source of the per-kernel profiler figures bench.py prints (warp instructions per
guess, DRAM bytes per launch), each pinned to the SASS hash of the binary it was captured from.

  python tools/kernel_profile.py --rep gpurun_out/x.ncu-rep --launch gpurun_out/x.json --summary profiles/r02_x_ncu_full.txt
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import ncu_summary  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rep", required=True)
ap.add_argument("--launch", required=True, help="JSON line printed by tools/profile_launch.py for the same command")
ap.add_argument("--summary", required=True, help="text summary to write under profiles/")
ap.add_argument("--match", default=None, help="substring of the kernel name to pick (default: from the launch line's key)")
args = ap.parse_args()

launch = json.loads([ln for ln in open(args.launch).read().splitlines() if ln.startswith("{")][-1])
raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = args.match or launch["key"].split("<")[0]
cands = [r for r in rows[2:] if want in r[hdr.index("Kernel Name")]]
if not cands:
    sys.exit("no launch of %s in %s" % (want, args.rep))


def val(r, k):
    return float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] not in ("", "n/a") else None


r = max(cands, key=lambda x: val(x, "gpu__time_duration.sum") or 0)  # the dominant launch
unit = lambda k: units[hdr.index(k)] if k in hdr else ""  # noqa: E731
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram = sum((val(r, k) or 0) * scale.get(unit(k), 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
g = launch["guesses"]
sm_count = launch.get("sm_count") or int(val(r, "device__attribute_multiprocessor_count") or 0)
entry = {
    "kernel": r[hdr.index("Kernel Name")], "mangled": launch["mangled"], "sass_sha256": launch["sass_sha256"],
    "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip(),
    "source": os.path.relpath(args.summary, ROOT), "profiled_launch": {k: launch.get(k) for k in ("kind", "bits", "level", "depth", "batches", "plan")},
    "guesses_of_profiled_launch": g,
    "warp_inst_of_profiled_launch": val(r, "smsp__inst_executed.sum"),
    "warp_inst_per_guess": (val(r, "smsp__inst_executed.sum") or 0) / g,
    "thread_inst_per_guess": (val(r, "smsp__thread_inst_executed.sum") or 0) / g if val(r, "smsp__thread_inst_executed.sum") else None,
    "dram_bytes_per_launch": dram,
    "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    # ALU pipe (integer add / compare / min / permute: 16 lanes per sub-partition, one warp instruction per two clocks): the
    # SM-cycles the pipe was busy, per guess -- what bench.py's roofline multiplies its live rate by
    "pipe_alu_cycles_active_pct": val(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    "sm_cycles_active": val(r, "sm__cycles_active.avg"),
    "sm_count": sm_count,
    "alu_pipe_sm_cycles_per_guess": ((val(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") or 0) / 100.0
                                     * (val(r, "sm__cycles_active.avg") or 0) * sm_count / g) or None,
    "pipe_fmaheavy_cycles_active_pct": val(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "avg_threads_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
    "pipe_alu_pct": val(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_fma_pct": val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "registers": val(r, "launch__registers_per_thread"),
    "ncu_duration_ms": (val(r, "gpu__time_duration.sum") or 0) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1),
    "unprofiled_ms": launch["ms"],
}
ncu_summary.main(args.rep, args.summary)
path = os.path.join(ROOT, "profiles", "current_kernels.json")
cur = json.load(open(path)) if os.path.exists(path) else {}
cur[launch["key"]] = entry
json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps({launch["key"]: entry}, indent=1))
