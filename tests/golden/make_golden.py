#!/usr/bin/env python3
"""Regenerates the golden fixtures under tests/golden/ from the reference checkout.

Run HERE (the build container), where /root/reference exists; the fixtures are committed
because the GPU box has no reference checkout.

  wheel11.json      the 480-entry table embedded in reference include/qimcifa.hpp:224-254
  ref_traces.json   outputs + hot-line traces of the UNMODIFIED reference sources compiled
                    against the compat shims (oracle/Makefile -> oracle/_ref/qimcifa_ref),
                    run single-threaded (QCF_REF_THREADS=1) so the visiting order is defined.

The trace is taken by the shim's cpp_int::operator% (oracle/shim/boost/multiprecision/
cpp_int.hpp): the right operand of every `N % g` (reference include/qimcifa.hpp:369).
"""
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("QCF_REFERENCE", "/root/reference")


def wheel11():
    src = open(os.path.join(REF, "include", "qimcifa.hpp")).read()
    m = re.search(r"wheel11\[480U\]\s*=\s*\{(.*?)\};", src, re.S)
    vals = [int(x) for x in re.findall(r"(\d+)U", m.group(1))]
    assert len(vals) == 480
    consts = {
        "BIGGEST_WHEEL": int(re.search(r"BIGGEST_WHEEL\s*=\s*(\d+)", src).group(1)),
        "SMALLEST_WHEEL": int(re.search(r"SMALLEST_WHEEL\s*=\s*(\d+)", src).group(1)),
        "MIN_RTD_LEVEL": int(re.search(r"MIN_RTD_LEVEL\s*=\s*(\d+)", src).group(1)),
    }
    json.dump({"source": "include/qimcifa.hpp:86-89,224-254", "wheel11": vals, **consts},
              open(os.path.join(HERE, "wheel11.json"), "w"))


def run_ref(n, node_count, node_id, level, threads=1, timeout=600):
    exe = os.path.join(ROOT, "oracle", "_ref", "qimcifa_ref")
    with tempfile.NamedTemporaryFile(suffix=".json", delete=False) as tf:
        tpath = tf.name
    env = dict(os.environ, QCF_REF_THREADS=str(threads), QCF_SHIM_TRACE=tpath)
    stdin = f"{n}\n{node_count}\n" + (f"{node_id}\n" if node_count > 1 else "") + f"{level}\n"
    out = subprocess.run([exe], input=stdin, capture_output=True, text=True, env=env, timeout=timeout)
    trace = json.load(open(tpath))
    os.unlink(tpath)
    m = re.search(r"Exact factor: Found (\d+) \* (\d+) = (\d+)", out.stdout)
    return {
        "n": str(n), "node_count": node_count, "node_id": node_id, "level": level,
        "found": [m.group(1), m.group(2)] if m else None,
        "count": trace["count"], "fnv1a64": trace["fnv1a64"], "last": trace["last"], "first": trace["first"],
        "stdout_tail": re.sub(r"\(Time elapsed: [^)]*\)", "(Time elapsed: T seconds)", out.stdout)[-260:],
    }


CASES = [
    # (N, nodeCount, nodeId, level)  -- 64-bit semiprimes of two 32-bit primes unless noted
    (4294966243 * 4294917283, 1, 0, 5),   # factor just below sqrt(N): found in the first batch
    (4294966243 * 4294917283, 1, 0, 6),
    (4294966243 * 4294917283, 1, 0, 7),
    (4294966243 * 4294917283, 1, 0, 8),
    (2999999929 * 4294979653, 1, 0, 5),   # factor in the 3rd batch from the top: two full batches + part
    (66695159677, 1, 0, 6),               # 255313 * 261229: literal L6 skips 255313 (wheel-phase defect D2)
    (66695159677, 1, 0, 5),
    (4294966243 * 4294917283, 8, 0, 5),   # reversed-sentinel defect D1: nodes 0-3 sweep nothing
    (4294966243 * 4294917283, 8, 3, 5),
    (4294966243 * 4294917283, 8, 4, 5),   # node 4 sweeps batch 3 (indices, not the top slice)
    (1000003 * 1000033, 1, 0, 5),         # 40-bit: one batch
    (1000003 * 1000033, 2, 1, 5),
]

def run_ref_gcd(n, node_count, node_id, level, threads=1, timeout=900):
    """The common-factor build of the UNMODIFIED reference (oracle/_ref/qimcifa_ref_gcd: the same sources compiled with
    IS_RSA_SEMIPRIME off, oracle/shim/gcd_mode/config.h).  Only stdout is recorded: the shim's `%` trace would also see
    the operands inside the reference's Euclid loop (include/qimcifa.hpp:201-210)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "qimcifa_ref_gcd")
    env = dict(os.environ, QCF_REF_THREADS=str(threads))
    stdin = f"{n}\n{node_count}\n" + (f"{node_id}\n" if node_count > 1 else "") + f"{level}\n"
    out = subprocess.run([exe], input=stdin, capture_output=True, text=True, env=env, timeout=timeout)
    m = re.search(r"Has common factor: Found (\d+) \* (\d+) = (\d+)", out.stdout)
    return {
        "n": str(n), "node_count": node_count, "node_id": node_id, "level": level,
        "found": [m.group(1), m.group(2)] if m else None,
        "stdout_tail": re.sub(r"\(Time elapsed: [^)]*\)", "(Time elapsed: T seconds)", out.stdout)[-260:],
    }


GCD_CASES = [
    # (N, nodeCount, nodeId, level): composite N whose prime factors are not wheel primes
    (1000003 * 1000033, 1, 0, 5),                 # semiprime: the same factor as the exact-divisor build
    (1009 * 1000003 * 1000033, 1, 0, 5),          # three primes: the first guess sharing any of them wins
    (29 * 31 * 1000003 * 1000033, 1, 0, 5),       # small factors: thousands of guesses qualify, the first visited wins
    (29 * 31 * 1000003 * 1000033, 1, 0, 7),
    (1009 * 1000003 * 1000033, 2, 1, 5),
    (8 * 1009 * 1000003, 1, 0, 5),                # even N with level-5 pre-check: 2 divides -> "Factors:" line, no sweep
]


def tuner_file():
    """qimcifa_calibration.ssv as written by the UNMODIFIED reference tuner (oracle/_ref/qimcifa_tuner_ref) for a
    100-bit semiprime.  Only the header and the cardinality column are meaningful: its timing loop times nothing
    in this snapshot."""
    import shutil

    exe = os.path.join(ROOT, "oracle", "_ref", "qimcifa_tuner_ref")
    n = 771819207259660166073186517813
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([exe], input=f"{n}\n8\n0\n", capture_output=True, text=True, cwd=d, timeout=1800)
        shutil.copy(os.path.join(d, "qimcifa_calibration.ssv"), os.path.join(HERE, "ref_tuner_100bit.ssv"))


if __name__ == "__main__":
    if "--gcd-only" in sys.argv:  # add the common-factor cases to the existing file
        path = os.path.join(HERE, "ref_traces.json")
        doc = json.load(open(path))
        doc["gcd_cases"] = []
        for c in GCD_CASES:
            print("running (common-factor build)", c, file=sys.stderr)
            doc["gcd_cases"].append(run_ref_gcd(*c))
        json.dump(doc, open(path, "w"), indent=1)
        sys.exit(0)
    wheel11()
    tuner_file()
    out = []
    for c in CASES:
        print("running", c, file=sys.stderr)
        out.append(run_ref(*c))
    gcd_out = []
    for c in GCD_CASES:
        print("running (common-factor build)", c, file=sys.stderr)
        gcd_out.append(run_ref_gcd(*c))
    json.dump({"generator": "tests/golden/make_golden.py", "threads": 1, "cases": out, "gcd_cases": gcd_out},
              open(os.path.join(HERE, "ref_traces.json"), "w"), indent=1)
