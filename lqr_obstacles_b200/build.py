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


QCONVEX_SRC = os.path.join(os.path.dirname(HERE), "tools", "qconvex", "qconvex_gpu.cpp")
QCONVEX_EXE = os.path.join(os.path.dirname(HERE), "tools", "qconvex", "qconvex")


def build_qconvex(force=False):
    """tools/qconvex/qconvex: the `qconvex` executable the unmodified reference spawns (MAIN:769-770), backed by
    the GPU hull through the C ABI.  Host-only C++ (g++), linked against liblqro.so with an rpath."""
    so = build()
    if not force and os.path.exists(QCONVEX_EXE) and os.path.getmtime(QCONVEX_EXE) > max(os.path.getmtime(QCONVEX_SRC), os.path.getmtime(so)):
        return QCONVEX_EXE
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", QCONVEX_EXE, QCONVEX_SRC, "-L" + HERE, "-llqro",
                           "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../../lqr_obstacles_b200"])
    return QCONVEX_EXE


HEADLESS_SRC = os.path.join(os.path.dirname(HERE), "tools", "headless", "lqro_headless.cpp")
HEADLESS_EXE = os.path.join(os.path.dirname(HERE), "tools", "headless", "lqro_headless")


def build_headless(force=False):
    """tools/headless/lqro_headless: the reference's simulator main loop headless on Linux over the C ABI (host-only C++)."""
    so = build()
    if not force and os.path.exists(HEADLESS_EXE) and os.path.getmtime(HEADLESS_EXE) > max(os.path.getmtime(HEADLESS_SRC), os.path.getmtime(so)):
        return HEADLESS_EXE
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I" + os.path.join(os.path.dirname(HERE), "include"), "-o", HEADLESS_EXE,
                           HEADLESS_SRC, "-L" + HERE, "-llqro", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../../lqr_obstacles_b200"])
    return HEADLESS_EXE


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_qconvex(force=True))
    print(build_headless(force=True))
