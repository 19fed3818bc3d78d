"""CPU-only: the C-ABI library loads, exports every symbol include/qimcifa_cuda.h declares, refuses to
run without a GPU (no CPU fallback), and its host-side range arithmetic agrees with the oracle."""
import os
import re
import random

import pytest

from oracle import oracle_c as oc
from oracle import oracle_py as op
from qimcifa_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "qimcifa_cuda.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(qcf_\w+)\s*\(", hdr, re.M))
    assert declared == set(abi.EXPORTS)
    lib = abi.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(abi.QcfError) as e:
        abi.Context(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_dropin_reference_program_refuses_without_gpu():
    """tests/dropin: the unmodified reference main() over the C ABI says why it finds nothing on a box without a GPU."""
    import subprocess

    import torch

    exe = os.path.join(ROOT, "tests", "_build", "qimcifa_dropin")
    if torch.cuda.is_available() or not os.path.exists(exe):
        pytest.skip("GPU present, or drop-in binary not built")
    r = subprocess.run([exe], input="1000036000099\n1\n5\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "no CPU fallback" in r.stderr and "Found" not in r.stdout


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qimcifa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f
                # nor the host emulation of the kernels that the CPU tests build (tests/emu): test infrastructure only
                assert "libqimcifa_emu" not in src and "emu_launch" not in src, f
    for f in ("bench.py", "__graft_entry__.py"):
        assert "libqimcifa_emu" not in open(os.path.join(ROOT, f)).read(), f


def test_node_split_matches_oracle():
    rnd = random.Random(5)
    for bits in (40, 64, 96, 100, 112, 128):
        for nc in (1, 2, 8):
            n = rnd.randrange(1 << (bits - 1), 1 << bits) | 1
            assert abi.node_split(n, nc) == oc.node_split(n, nc)
    n = 18446524746962277769
    assert abi.node_split(n, 1) == op.node_split(n, 1)


def test_count_guesses_matches_oracle():
    for level in range(5, 10):
        for lo, hi in ((0, 1), (0, 3), (5, 9), (262143, 262144), (17179869183, 17179869184)):
            plo, phi = op.batch_index_range(lo)[0], op.batch_index_range(hi - 1)[1]
            assert abi.count_guesses(level, lo, hi) == oc.intent_count(level, plo, phi)
        assert abi.count_guesses(level, 4, 4) == 0


def test_intent_node_slices_tile_batches():
    n = (2**56 - 5) * (2**56 - 3)
    _, _, r, bc = abi.node_split(n, 8)
    spans = [abi.node_batches(bc, r, k) for k in range(8)]
    assert spans[0][1] == bc and spans[-1][0] == 0
    for a, b in zip(spans, spans[1:]):
        assert a[0] == b[1]
    for k in range(8):
        lo, hi = spans[k]
        assert list(range(hi - 1, hi - 4, -1)) == op.intent_node_batches(n, 8, k)[:3]


def test_header_is_plain_c(tmp_path):
    """include/qimcifa_cuda.h compiles as C and links against the library; errors come back as codes."""
    import subprocess

    exe = str(tmp_path / "abi_c_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_c_smoke.c"), "-o", exe, "-L", os.path.join(ROOT, "qimcifa_b200"),
                           "-lqimcifa_cuda", "-Wl,-rpath," + os.path.join(ROOT, "qimcifa_b200")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    n = 18446524746962277769
    s, _, nr, bc = oc.node_split(n, 8)
    assert f"sqrt={s} nodeRange={nr} batchCount={bc}" in out
    assert f"count={oc.intent_count(5, 2, oc.BW + 1)} rc=0" in out
    assert "bad=-1 msg=n_limbs must be in [1,4]" in out
    import torch

    assert ("create=0" in out) if torch.cuda.is_available() else ("create=-4" in out)


def test_level_10_count_matches_oracle():
    """qcf_count_guesses at level 10 (host only): inclusion-exclusion over the ten wheel primes."""
    from oracle import oracle_c as oc
    from qimcifa_b200 import abi

    for lo, hi in ((0, 1), (0, 7), (123, 4567)):
        assert abi.count_guesses(10, lo, hi) == oc.intent_count(10, lo * abi.BIGGEST_WHEEL + 2, hi * abi.BIGGEST_WHEEL + 1)
        assert abi.count_guesses(10, lo, hi) < abi.count_guesses(9, lo, hi)


def test_host_only_entry_points_reject_bad_levels():
    import pytest as _pytest

    from qimcifa_b200 import abi

    with _pytest.raises(abi.QcfError):
        abi.count_guesses(11, 0, 1)
    with _pytest.raises(abi.QcfError):
        abi.count_guesses(4, 0, 1)
    with _pytest.raises(abi.QcfError):
        abi.plan_filter((1 << 96) - 17, 10, 1 << 40, 1 << 41)  # chains run over levels 5..9 only
    assert abi.plan_filter((1 << 96) - 17, 9, (1 << 47) + 1, 1 << 48) is not None
    assert abi.plan_filter((1 << 96) - 17, 5, 1000, 2000) is None  # too short for any chain
