"""Builds libd3q19b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libd3q19b200.so")
SOURCES = ["d3q19_api.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "d3q19_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA extension; returns the path of the shared library."""
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


DRIVER_SRC = os.path.join(HERE, "host", "channel_driver.cpp")
DRIVER = os.path.join(HERE, "host", "channel_driver")


def build_driver(force=False):
    """The compiled-language host above the C-ABI (host/channel_driver.cpp: main.f90 + para + initial in C++),
    linked against the in-tree library.  -ffp-contract=off: IEEE evaluation of the reference's expression order."""
    lib = build()
    hdr = os.path.join(HERE, "..", "include", "d3q19_b200.h")
    if not force and os.path.exists(DRIVER) and os.path.getmtime(DRIVER) >= max(
            os.path.getmtime(DRIVER_SRC), os.path.getmtime(hdr), os.path.getmtime(lib)):
        return DRIVER
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-I", os.path.join(HERE, "..", "include"),
           "-o", DRIVER, DRIVER_SRC, "-L", HERE, "-ld3q19b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/..", "-pthread"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout)
    return DRIVER


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_driver(force=True))
