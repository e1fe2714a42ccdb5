"""Build recipe of libpps_b200.so (in-tree, sm_100a only) and of the C++ driver.

    python -m parallelpoissonsolver_b200.build            # library + driver
nvcc cross-compiles without a GPU; the built files are git-ignored but travel with gpurun snapshots.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(CSRC, "libpps_b200.so")
DRIVER_DIR = os.path.join(PKG, "driver")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    out = []
    for root, _, files in os.walk(d):
        out += [os.path.join(root, f) for f in files if f.endswith(exts)]
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    deps = _sources(CSRC, (".cu", ".cuh", ".hpp", ".h")) + [os.path.join(ROOT, "include", "pps_b200.h")]
    if not force and not _newer(LIB, deps):
        return LIB
    cus = sorted(s for s in deps if s.endswith(".cu"))
    cmd = [NVCC] + ARCH + NVCC_FLAGS + ["--threads", "4", "-shared", "-cudart", "static", "-o", LIB] + cus + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    if verbose:
        print(r.stderr)
    return LIB


def build_driver(force: bool = False) -> str:
    """solverPoisson: the C++ host driver (same argv and stdout as the reference's main.cpp)."""
    src = os.path.join(DRIVER_DIR, "main.cpp")
    exe = os.path.join(DRIVER_DIR, "solverPoisson")
    if not os.path.exists(src):
        return ""
    deps = _sources(DRIVER_DIR, (".cpp", ".hpp", ".h")) + _sources(os.path.join(ROOT, "include"), (".h", ".hpp"))
    if not force and not _newer(exe, deps + [LIB]):
        build_driver_host()
        return exe
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-std=c++17", "-O3", "-DNDEBUG", "-pthread", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "include", "reference_compat"), "-I" + DRIVER_DIR, src, "-o", exe,
           "-L" + CSRC, "-lpps_b200", "-Wl,-rpath," + CSRC, "-Wl,-rpath,$ORIGIN/../csrc"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("driver build failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    build_driver_host(force=True)
    return exe


DRIVER_HOST_LIB = os.path.join(DRIVER_DIR, "libpps_driver_host.so")


def build_driver_host(force: bool = False) -> str:
    """libpps_driver_host.so: the driver's host-side setProblem() as a C-callable helper (driver/setproblem.cpp) for
    callers that are not C++ (bench.py, tools/).  No CUDA, no oracle code."""
    src = os.path.join(DRIVER_DIR, "setproblem.cpp")
    deps = [src, os.path.join(DRIVER_DIR, "solverSetup.hpp")]
    if not force and not _newer(DRIVER_HOST_LIB, deps):
        return DRIVER_HOST_LIB
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-std=c++17", "-O3", "-DNDEBUG", "-pthread", "-fPIC", "-shared", "-I" + DRIVER_DIR,
           "-I" + os.path.join(ROOT, "include", "reference_compat"), src, "-o", DRIVER_HOST_LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("driver host helper build failed:\n%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return DRIVER_HOST_LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_library(force=force, verbose="-v" in sys.argv))
    print(build_driver(force=force))
