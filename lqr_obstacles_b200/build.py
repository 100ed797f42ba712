"""Build liblqro.so (the product: CUDA kernels + C ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU.  -fmad=false is part of the arithmetic contract (csrc/lqro_common.cuh).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liblqro.so")
SRC = os.path.join(HERE, "csrc", "lqro_api.cu")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]


def _sources():
    d = os.path.join(HERE, "csrc")
    out = [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".cu", ".cuh"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "lqro.h"))
    return out


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
