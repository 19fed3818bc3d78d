import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "timeout: per-test limit (pytest-timeout; ignored where the plugin is absent)")
    # build the native pieces if a fresh checkout has not been built yet (nvcc cross-compiles without a GPU).  Without nvcc
    # (or after a failed build) only the tests that need libqimcifa_cuda.so are skipped -- with the reason -- and the pure
    # CPU ones (oracle, device arithmetic compiled by g++, the kernel emulation) still run.
    from qimcifa_b200 import build as qbuild

    config._qcf_build_error = None
    try:
        qbuild.build_all()
    except Exception as e:  # noqa: BLE001 -- whatever the toolchain raised is the skip reason
        config._qcf_build_error = "libqimcifa_cuda.so could not be built: %s" % (str(e).splitlines() or ["?"])[0]


NEEDS_CUDA_LIBRARY = ("test_abi", "test_filter_math", "test_host", "test_multi_gloo", "test_gpu_")


def pytest_collection_modifyitems(config, items):
    err = getattr(config, "_qcf_build_error", None)
    if not err:
        return
    skip = pytest.mark.skip(reason=err)
    for item in items:
        if any(key in item.nodeid for key in NEEDS_CUDA_LIBRARY):
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _qcf_default_options():
    """Every test starts and ends with the library's test / calibration switches (qcf_set_option) at their defaults."""
    yield
    from qimcifa_b200 import abi

    if abi._LIB is not None:
        for name in ("xchain_two_stage", "chain_min_log2", "plan_force", "plan_drift", "plan_packed", "cost_model", "epoch_trips", "debug", "debug_stats"):
            abi.set_option(name, None)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
