"""Build libhikari_b200.so in-tree: hand-written CUDA for sm_100a + the C++ host mirror, linked into one shared
library that exports the C ABI of include/hikari_b200.h and include/hikari_host.h.

nvcc cross-compiles without a GPU.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX fallback for other architectures
  -fmad=false / -ffp-contract=off           no implicit FMA contraction: every fused op is an explicit fmaf() in
                                            include/hk_math.h, which is what makes device results comparable bit-for-bit
                                            with the CPU oracle
  -lineinfo                                 ncu source page maps to these files
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INC = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libhikari_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("HK_CXX", "/usr/bin/g++")

CU = ["csrc/context.cu", "csrc/kernels_light.cu", "csrc/kernels_post.cu", "csrc/kernels_upscale.cu"]
CPP = ["host/hikari.cpp", "host/hikari_capi.cpp"]
HEADERS = ["csrc/hk_device.cuh", "csrc/hk_kernels.h", "host/hikari.hpp", "../include/hk_math.h", "../include/hk_layout.h",
           "../include/hikari_b200.h", "../include/hikari_host.h"]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-I", INC, "-I", os.path.join(HERE, "csrc")]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-I", INC, "-I", os.path.join(HERE, "host")]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    extra_env = os.environ.get("HK_NVCC_EXTRA", "").split()   # tuning experiments, e.g. -DHK_MINB_INDIRECT=5
    if extra_env:
        force = True
    headers = [os.path.join(HERE, h) for h in HEADERS]
    jobs = []
    objs = []
    for src in CU:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            extra = (["-Xptxas", "-v"] if ptxas_info else []) + extra_env
            jobs.append([NVCC] + NVCC_FLAGS + extra + ["-c", s, "-o", o])
    for src in CPP:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append([CXX] + CXX_FLAGS + ["-c", s, "-o", o])
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            logs = list(ex.map(_run, jobs))
    if jobs or not os.path.exists(LIB):
        logs.append(_run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                "-cudart", "static", "-Xcompiler", "-fPIC"]))
    if verbose or ptxas_info:
        print("\n".join(l for l in logs if l.strip()))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
