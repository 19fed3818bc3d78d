"""GPU: the host drivers keep the reference's dialog and output lines (reference src/qimcifa.cpp:61-207,
src/qimcifa_tuner.cpp:100-181, include/qimcifa.hpp:212-222) while sweeping on the GPU."""
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "qimcifa_b200", "bin")


def run(exe, stdin, args=(), cwd=None):
    r = subprocess.run([os.path.join(BIN, exe), *args], input=stdin, capture_output=True, text=True, cwd=cwd, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


def norm(s):
    return re.sub(r"\(Time elapsed: [^)]*\)", "(Time elapsed: T seconds)", s)


def test_qimcifa_dialog_matches_reference_output(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "ref_traces.json")))["cases"]
    for c in cases:
        if c["node_count"] != 1 or c["level"] != 5:
            continue
        out = norm(run("qimcifa", f"{c['n']}\n1\n5\n", ("--gpus", "1")))
        # the reference's stdout (prompts + result lines), byte for byte after the elapsed time is masked
        assert out.endswith(c["stdout_tail"][-200:]), (out, c["stdout_tail"])
        assert out.startswith("Number to factor: Bits to factor: ")


def test_qimcifa_node_dialog_and_level_clamp():
    n = 1000003 * 1000033
    out = run("qimcifa", f"{n}\n0\n2\n5\n1\n3\n", ("--gpus", "1"))  # invalid node count, invalid id, level 3 -> 5
    assert "Invalid node count choice!" in out and "Invalid node ID choice!" in out
    assert "Which node is this? (0-1): " in out
    assert "Exact factor: Found 1000003 * 1000033 = 1000036000099" in out
    out = run("qimcifa", f"{n}\n1\n12\n", ("--gpus", "1", "--stats"))  # level 12 -> 9
    assert "Exact factor: Found 1000003 * 1000033 = 1000036000099" in out and "(Stats:" in out
    # pre-checks (src/qimcifa.cpp:128-142)
    assert "Number to factor is a perfect square: 1000003 * 1000003 = 1000006000009" in run("qimcifa", "1000006000009\n1\n5\n")
    assert "Factors: 2 * 500018000049" in run("qimcifa", "1000036000098\n1\n5\n")
    # node 1 of 2 owns the bottom slice and finds the small factor there; node 0 sweeps its slice without a hit
    big = 4294966243 * 4294917283
    assert "Exact factor: Found" in run("qimcifa", f"{big}\n2\n0\n5\n", ("--gpus", "1"))
    assert "Exact factor: Found" not in run("qimcifa", f"{big}\n2\n1\n5\n", ("--gpus", "1"))


def test_tuner_calibration_round_trip(tmp_path):
    n = 18446524746962277769
    out = run("qimcifa_tuner", f"{n}\n1\n0\n", ("--batches", "2"), cwd=str(tmp_path))
    assert "Calibrated reverse trial division level: " in out and "Estimated average time to exit: " in out
    rows = open(tmp_path / "qimcifa_calibration.ssv").read().strip().split("\n")
    assert rows[0] == "level, cardinality, batch time (s), cost (s)"
    assert [r.split()[0] for r in rows[1:]] == ["5", "6", "7", "8", "9"]
    from oracle import oracle_py as op

    card = op.tuner_cardinalities(n)
    for r in rows[1:]:
        lvl, c, t, cost = r.split()
        assert int(c) == card[int(lvl)] and float(t) > 0 and float(cost) > 0
    best = min(rows[1:], key=lambda r: float(r.split()[3])).split()[0]
    assert f"Calibrated reverse trial division level: {best}" in out
    out2 = run("qimcifa", f"{n}\n1\n-1\n", ("--gpus", "1"), cwd=str(tmp_path))
    assert f"Calibrated wheel factorization level: {best}" in out2
    assert "Exact factor: Found 4294917283 * 4294966243 = 18446524746962277769" in out2


def test_qimcifa_random_mode():
    import sympy

    p, q = sympy.prevprime(990000), sympy.nextprime(1010000)
    out = run("qimcifa", f"{p * q}\n1\n5\n", ("--gpus", "1", "--random", "42", "--random-budget", str(1 << 31), "--stats"))
    assert f"Exact factor: Found {p} * {q} = {p * q}" in out
    out2 = run("qimcifa", f"{p * q}\n1\n5\n", ("--gpus", "1", "--random", "42", "--random-budget", str(1 << 31)))
    assert norm(out2).split("(Stats")[0] == norm(out).split("(Stats")[0]  # deterministic in the seed


def test_qimcifa_random_mode_chain_draws():
    """64-bit golden semiprime: large enough for the filter kernel, so --random uses chain draws (QCF_RANDOM_CHAINS); a budget
    of 2^36 guess numbers covers the 8.9e8 guesses below sqrt(N) dozens of times over: the factor is drawn."""
    n = 18446524746962277769
    out = run("qimcifa", f"{n}\n1\n5\n", ("--gpus", "1", "--random", "7", "--random-budget", str(1 << 36), "--stats"))
    assert "Exact factor: Found 4294917283 * 4294966243 = 18446524746962277769" in out
    out2 = run("qimcifa", f"{n}\n1\n5\n", ("--gpus", "1", "--random", "7", "--random-budget", str(1 << 36), "--random-independent"))
    assert "Exact factor: Found 4294917283 * 4294966243 = 18446524746962277769" in out2


def _gpu_count():
    import torch

    return torch.cuda.device_count()


def test_qimcifa_two_gpus_share_the_dispenser():
    """One process, one worker thread per GPU pulling batch runs from the node's dispenser (the reference's thread
    model, src/qimcifa.cpp:160-178); the finder's finish() stops the other GPU through its found flag."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    import sympy

    p = sympy.prevprime(2**48 - 2**42)  # ~1.6 % below 2^48: a few launches down from the top
    q = sympy.prevprime(2**48 - 12345)
    n = p * q
    out = run("qimcifa", f"{n}\n1\n5\n", ("--gpus", "2", "--stats", "--batches-per-launch", "512"))
    assert f"Exact factor: Found {p} * {q} = {n}" in out
    assert out.count("Exact factor: Found") == 1 and "2 GPU(s)" in out
    # the node split still applies on top: node 1 of 2 (bottom half) does not hold this factor
    out = run("qimcifa", f"{1000003 * 1000033}\n2\n1\n5\n", ("--gpus", "2"))
    assert "Exact factor: Found 1000003 * 1000033" in out


def test_gpu_tuner_writes_the_reference_tuners_cardinalities(tmp_path, golden_dir):
    """Same header and same cardinality column as the file the unmodified reference tuner wrote for this number."""
    ref = open(os.path.join(golden_dir, "ref_tuner_100bit.ssv")).read().strip().split("\n")
    run("qimcifa_tuner", "771819207259660166073186517813\n8\n0\n", ("--batches", "64"), cwd=str(tmp_path))
    mine = open(tmp_path / "qimcifa_calibration.ssv").read().strip().split("\n")
    assert mine[0] == ref[0]
    assert [r.split()[:2] for r in mine[1:]] == [r.split()[:2] for r in ref[1:]]


def test_qimcifa_gcd_mode_matches_the_reference_common_factor_build(golden_dir):
    """`qimcifa --gcd` against stdout of the unmodified reference compiled with IS_RSA_SEMIPRIME off."""
    cases = json.load(open(os.path.join(golden_dir, "ref_traces.json")))["gcd_cases"]
    seen = 0
    for c in cases:
        if c["node_count"] != 1:
            continue
        out = norm(run("qimcifa", f"{c['n']}\n1\n{c['level']}\n", ("--gpus", "1", "--gcd")))
        assert out.endswith(c["stdout_tail"][-150:]), (out, c["stdout_tail"])
        seen += 1
    assert seen >= 4


def test_qimcifa_level_10_is_clamped_like_the_reference_unless_asked():
    n = 1000003 * 1000033
    out9 = run("qimcifa", f"{n}\n1\n10\n", ("--gpus", "1", "--stats"))
    out10 = run("qimcifa", f"{n}\n1\n10\n", ("--gpus", "1", "--stats", "--level-10"))
    for out in (out9, out10):
        assert "Exact factor: Found 1000003 * 1000033 = 1000036000099" in out
    g9 = int(re.search(r"\(Stats: (\d+) guesses", out9).group(1)) if re.search(r"\(Stats: (\d+) guesses", out9) else None
    g10 = int(re.search(r"\(Stats: (\d+) guesses", out10).group(1)) if re.search(r"\(Stats: (\d+) guesses", out10) else None
    if g9 and g10:
        assert g10 < g9  # level 10 tests fewer guesses over the same batches


def test_reference_program_with_gpu_sweep(golden_dir):
    """The UNMODIFIED reference src/qimcifa.cpp (its main(), dialog, pre-checks, node split, dispenser globals, worker
    threads and printSuccess) compiled with tests/dropin/qimcifa.hpp, which replaces only getSmoothNumbers()
    (reference include/qimcifa.hpp:384-399) by qcf_sweep: INTEGRATION.md section 2a.  The binary is built where the
    reference checkout exists and travels to the GPU box prebuilt."""
    exe = os.path.join(ROOT, "tests", "_build", "qimcifa_dropin")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/qimcifa_dropin not built (no reference checkout at build time)")
    cases = json.load(open(os.path.join(golden_dir, "ref_traces.json")))["cases"]
    seen = 0
    for c in cases:
        if c["node_count"] != 1 or not c["found"]:
            continue
        r = subprocess.run([exe], input=f"{c['n']}\n1\n{c['level']}\n", capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        out, tail = norm(r.stdout), norm(c["stdout_tail"])
        if c["level"] == 5:
            # intent == literal at nodeCount 1, level 5: the reference's own stdout, byte for byte (elapsed time masked)
            assert out.endswith(tail[-200:]), (out, tail)
        else:
            # above level 5 the snapshot's wheel is out of phase: same factorisation, either order
            m = re.search(r"Exact factor: Found (\d+) \* (\d+) = (\d+)", tail)
            m2 = re.search(r"Exact factor: Found (\d+) \* (\d+) = (\d+)", out)
            assert m and m2 and sorted(m.groups()) == sorted(m2.groups()), (out, tail)
        assert out.startswith("Number to factor: Bits to factor: ")
        seen += 1
    assert seen >= 2
