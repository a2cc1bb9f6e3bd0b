"""Builds libc3poa_gpu.so (sm_100a) in-tree with nvcc.  No torch involved."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "c3poa_gpu.cu")
SRC_HOST = os.path.join(HERE, "csrc", "ingest.cpp")
OUT = os.path.join(HERE, "libc3poa_gpu.so")
DEPS = [SRC, SRC_HOST] + [os.path.join(HERE, "csrc", f) for f in ("common.cuh", "conk.cuh", "peaks.cuh", "poa.cuh", "poa_lane.cuh", "poa_grp.cuh", "poa_graph.cuh")] + [
    os.path.join(os.path.dirname(HERE), "include", "c3poa_gpu.h")]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.sep not in p or os.path.exists(p)):
            return p
    return "nvcc"


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-o", out, SRC, SRC_HOST, "-lz"]
    cmd += [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
