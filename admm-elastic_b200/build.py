"""Builds the native libraries of the package in-tree (nvcc / g++), so the .so files travel with a
repository snapshot:

  libadmm_b200.so        CUDA kernels + C-ABI (include/admm_b200.h), sm_100a only
  libadmm_b200_host.so   C++ host mirror of admm::Solver / EnergyTerm / LinearSolver + extern "C" view

Nothing here needs a GPU: nvcc cross-compiles.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB_CUDA = os.path.join(HERE, "libadmm_b200.so")
LIB_HOST = os.path.join(HERE, "libadmm_b200_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _find(tool, fallbacks):
    p = shutil.which(tool)
    if p:
        return p
    for f in fallbacks:
        if os.path.exists(f):
            return f
    raise RuntimeError("%s not found" % tool)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "admm_b200.h")]
    if not force and not _newer(LIB_CUDA, srcs):
        return LIB_CUDA
    nvcc = _find("nvcc", ["/usr/local/cuda/bin/nvcc"])
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "/usr/bin/g++", "-o", LIB_CUDA, os.path.join(CSRC, "admm_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_CUDA


def build_host(force=False):
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST))] + [os.path.join(ROOT, "include", "admm_b200.h"), LIB_CUDA]
    if not force and not _newer(LIB_HOST, srcs):
        return LIB_HOST
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-o", LIB_HOST,
           os.path.join(HOST, "c_api.cpp"), "-L" + HERE, "-ladmm_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd, cwd=HERE)
    return LIB_HOST


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_CUDA, LIB_HOST)
