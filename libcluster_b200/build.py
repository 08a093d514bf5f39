"""In-tree build of libcluster_b200's native library for sm_100a.

    python -m libcluster_b200.build        # or __graft_entry__.build()

Produces libcluster_b200/_lib/liblcb200.so (git-ignored, travels with gpurun).
nvcc cross-compiles here without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "liblcb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fopenmp,-Wall,-Wno-unused-function"]
SOURCES = ["kernels.cu", "tc_kernels.cu", "mstep.cu", "engine.cu", "engine_dev.cu", "host_model.cpp", "c_api.cpp"]


def _newer(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "libcluster_b200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    objs, rebuilt = [], False
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        obj = os.path.join(OUT_DIR, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _newer(src, obj):
            cmd = [NVCC] + ARCH + COMMON + ["-Xptxas", "-v"] * bool(verbose) + ["-x", "cu", "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            rebuilt = True
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lcudart_static", "-ldl", "-lpthread", "-lrt"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
