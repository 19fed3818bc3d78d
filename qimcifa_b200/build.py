"""Builds the native pieces in-tree (no JIT cache, so the .so files travel with the repo snapshot):

  qimcifa_b200/libqimcifa_cuda.so   CUDA kernels + C ABI (include/qimcifa_cuda.h), sm_100a only
  qimcifa_b200/bin/qimcifa          host driver (same stdin dialog as the reference's `qimcifa`)
  qimcifa_b200/bin/qimcifa_tuner    host tuner  (same dialog / .ssv format as `qimcifa_tuner`)

Usage: python -m qimcifa_b200.build [--force]
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "libqimcifa_cuda.so")
BIN = os.path.join(PKG, "bin")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _newer(srcs, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def _run(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r


def build_cuda(force=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [
        os.path.join(ROOT, "include", "qimcifa_cuda.h")]
    if not force and not _newer(deps, LIB):
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB] + srcs
    _run(cmd, log=os.path.join(PKG, "build_cuda.log"))
    return LIB


def build_variant(tag, defines):
    """A/B builds for kernel experiments (tools/ only): qimcifa_b200/libqimcifa_cuda_<tag>.so with extra -D flags; a process
    picks it with QCF_LIB_VARIANT=<tag> (qimcifa_b200/abi.py).  Never used by the tests, the bench or the host drivers."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    out = os.path.join(PKG, "libqimcifa_cuda_%s.so" % tag)
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + ["-D" + d for d in defines] + [
        "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", out] + srcs
    _run(cmd, log=os.path.join(PKG, "build_cuda_%s.log" % tag))
    return out


def build_host(force=False):
    os.makedirs(BIN, exist_ok=True)
    outs = []
    hdrs = glob.glob(os.path.join(HOST, "*.hpp")) + [os.path.join(ROOT, "include", "qimcifa_cuda.h")]
    for name in ("qimcifa", "qimcifa_tuner"):
        src = os.path.join(HOST, name + ".cpp")
        if not os.path.exists(src):
            continue
        out = os.path.join(BIN, name)
        if force or _newer([src] + hdrs, out):
            cxx = os.environ.get("CXX", "g++")
            _run([cxx, "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), "-I", HOST, src,
                  "-o", out, "-L", PKG, "-lqimcifa_cuda", "-Wl,-rpath,$ORIGIN/..", "-ldl"])
        outs.append(out)
    return outs


def build_all(force=False):
    lib = build_cuda(force)
    return [lib] + build_host(force)


if __name__ == "__main__":
    for p in build_all("--force" in sys.argv):
        print(p)
