"""CPU-only: the host-side C++ (BigInteger + the reference-surface helpers in qimcifa_b200/host) against Python ints
and the oracle."""
import math
import os
import random
import subprocess

import pytest

from oracle import oracle_py as op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver():
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "host_bigint_driver")
    src = os.path.join(ROOT, "tests", "host_bigint_driver.cpp")
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "qimcifa_b200", "host"), src, "-o", exe, "-L",
                           os.path.join(ROOT, "qimcifa_b200"), "-lqimcifa_cuda", "-Wl,-rpath," + os.path.join(ROOT, "qimcifa_b200")])

    def run(lines):
        r = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
        return r.stdout.strip("\n").split("\n")

    return run


def test_bigint_arithmetic(driver):
    rnd = random.Random(3)
    M = 1 << 256
    lines, want = [], []
    for _ in range(300):
        a = rnd.randrange(0, 1 << rnd.choice((1, 31, 64, 65, 127, 128, 200, 255)))
        b = rnd.randrange(1, 1 << rnd.choice((1, 31, 64, 65, 127, 128)))
        for o, f in (("add", lambda: (a + b) % M), ("sub", lambda: (a - b) % M), ("mul", lambda: (a * b) % M),
                     ("div", lambda: a // b), ("mod", lambda: a % b), ("and", lambda: a & b)):
            lines.append(f"{o} {a} {b}")
            want.append(str(f()))
        s = rnd.randrange(0, 200)
        lines.append(f"shr {a} {s}")
        want.append(str(a >> s))
        lines.append(f"shl {a} {s}")
        want.append(str((a << s) % M))
        lines.append(f"cmp {a} {b}")
        want.append("".join(str(int(x)) for x in (a < b, a <= b, a > b, a >= b, a == b, a != b)))
        lines.append(f"mixed {a % (1 << 128)} 0")
        want.append(str(((3 * (a % (1 << 128)) + 6) // 5) % 1000003))
    assert driver(lines) == want


def test_host_helpers_match_oracle(driver):
    rnd = random.Random(4)
    lines, want = [], []
    for bits in (8, 40, 64, 96, 112, 128):
        for _ in range(20):
            n = rnd.randrange(1 << (bits - 1), 1 << bits)
            lines += [f"sqrt {n} 0", f"bwd {n % (1 << 100)} 0", f"fwd {n % (1 << 100)} 0", f"pow2 {n} 0", f"pow2 {1 << (bits - 1)} 0"]
            want += [str(math.isqrt(n)), str(op.backward(n % (1 << 100))), str(op.forward(n % (1 << 100))), "0" if n & (n - 1) else "1", "1"]
            for nc in (1, 8):
                _, fr, nr, bc = op.node_split(n, nc)
                lines.append(f"split {n} {nc}")
                want.append(f"{fr} {nr} {bc}")
    s = 2**64 - 59
    lines += [f"sqrt {s * s} 0", f"sqrt {s * s - 1} 0", f"dbl {1 << 70} 0"]
    want += [str(s), str(s - 1), str(1 << 60)]
    assert driver(lines) == want


def test_dispenser_intent_order(driver):
    # node k owns the k-th slice from the top, handed out descending
    out = driver(["disp 5 %d" % (8 * 16 + k) for k in range(8)])
    for k, line in enumerate(out):
        spans = [tuple(int(x) for x in s.split(":")) for s in line.strip(",").split(",")]
        top = 40 - 5 * k
        assert spans == [(top - 2, top), (top - 4, top - 2), (top - 5, top - 4)]
        flat = [b for lo, hi in spans for b in range(hi - 1, lo - 1, -1)]
        assert flat == list(range(top - 1, top - 6, -1))
