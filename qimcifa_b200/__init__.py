"""qimcifa_b200 -- B200-native (sm_100a) reverse-wheel trial-division sweep behind a C ABI.

The product is qimcifa_b200/libqimcifa_cuda.so (include/qimcifa_cuda.h) plus the C++ host drivers
under qimcifa_b200/host/.  `abi` is the ctypes view of the same boundary used by tests and bench.py.
Nothing here falls back to the CPU.
"""
from . import abi  # noqa: F401
from .abi import Context, QcfError  # noqa: F401
