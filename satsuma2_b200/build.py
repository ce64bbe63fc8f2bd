"""Builds libsatsuma_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m satsuma2_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsatsuma_b200.so")
SOURCES = ["sx_kernels.cu", "sx_engine.cu", "sx_multi.cu", "sx_kmatch.cu"]
HEADERS = ["sx_kernels.h", "sx_fft.cuh", "sx_scan.cuh", os.path.join("..", "..", "include", "satsuma_xcorr.h"),
           os.path.join("..", "..", "include", "satsuma_kmatch.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-Xlinker", "-Bsymbolic", "-Xlinker", "--exclude-libs,ALL",
    # double-precision scoring must not be FMA-contracted (it is compared operation by operation
    # with the reference's x86-64 arithmetic); FP32 FFT code may fuse.
    "-Xptxas", "-v",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libsatsuma_b200.so")
    log = os.path.join(HERE, "_build")
    os.makedirs(log, exist_ok=True)
    with open(os.path.join(log, "ptxas.log"), "w") as f:
        f.write(proc.stdout + proc.stderr)
    return LIB


HOST = os.path.join(HERE, "host")
BIN = os.path.join(HERE, "bin")
HOST_PROGRAMS = {"HomologyByXCorr": "homology_by_xcorr_main.cc", "HomologyByXCorrSlave": "homology_slave_main.cc",
                 "XCorrMatchTool": "match_tool_main.cc", "KMatch": "kmatch_main.cc"}


def build_host(force: bool = False) -> list:
    """C++ host programs above the C ABI (drop-ins for the reference's HomologyByXCorr[Slave])."""
    build(force=False)
    os.makedirs(BIN, exist_ok=True)
    outs = []
    common = [os.path.join(HOST, "sx_host.cc")]
    for name, main in HOST_PROGRAMS.items():
        exe = os.path.join(BIN, name)
        srcs = ([] if name == "KMatch" else common) + [os.path.join(HOST, main)]  # KMatch only needs the C ABI
        deps = srcs + [os.path.join(HOST, "sx_host.h"), LIB]
        if force or not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
            cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe] + srcs + [
                "-L" + HERE, "-lsatsuma_b200", "-Wl,-rpath,$ORIGIN/..", "-lpthread"]
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                sys.stderr.write(proc.stdout + proc.stderr)
                raise RuntimeError(f"g++ failed building {name}")
        outs.append(exe)
    return outs


if __name__ == "__main__":
    if "--host" in sys.argv:
        print("\n".join(build_host(force="--force" in sys.argv)))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
