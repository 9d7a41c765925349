"""Builds the CUDA extension in-tree: csrc/fjsp_api.cu -> libfjsp_b200.so (sm_100a).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfjsp_b200.so")
SOURCES = [os.path.join(CSRC, f) for f in ("fjsp_api.cu", "fjsp_core.cuh", "fjsp_host.h", "fjsp_layout.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "fjsp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550"]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def stale():
    return not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SOURCES)


def build_cuda(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "fjsp_api.cu")]
    subprocess.check_call(cmd)
    return LIB


def build_variant(name, defines):
    """A tuning / diagnostic build next to the product library (select it with FJSP_B200_LIB),
    e.g. build_variant("trace", ["-DFJ_TRACE"]) for tools/cta_trace.py."""
    out = os.path.join(HERE, "libfjsp_b200_%s.so" % name)
    subprocess.check_call([nvcc_path()] + NVCC_FLAGS + list(defines) + ["-o", out, os.path.join(CSRC, "fjsp_api.cu")])
    return out


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
