"""Builds igx_raytracing_b200/librtb200.so (sm_100a) in-tree with nvcc.

    python -m igx_raytracing_b200.build [--force] [--verbose]

The library is the product: there is no Python or CPU fallback.  `-fmad=false` is part of the numerical
contract (csrc/rtb_math.cuh): the shader arithmetic is evaluated as separate IEEE operations and fused
multiply-adds are written explicitly where they are wanted.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.environ.get("RTB_LIB_OUT", os.path.join(HERE, "librtb200.so"))   # RTB_LIB_OUT / RTB_NVCC_EXTRA: tuning variants
SOURCES = ["rtb_api.cu", "rtb_kernels.cu", "rtb_refit.cu", "rtb_sort.cu", "rtb_build.cu", "rtb_probe.cu", "rtb_bvh.cpp", "rtb_host.cpp"]
HEADERS = ["rtb_types.h", "rtb_math.cuh", "rtb_crmath.h", "rtb_kernels.cuh", "rtb_trace8.cuh", "rtb_trace8p.cuh", "rtb_trace8f.cuh", "rtb_trace8b.cuh", "rtb_trace8s.cuh", "rtb_path.cuh", "rtb_node8_encode.h", "rtb_bvh.h",
           os.path.join(ROOT, "include", "rtb200.h"), os.path.join(ROOT, "include", "igx_rt.hpp")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O3,-pthread,-ffp-contract=off", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    build_dir = os.path.join(HERE, "build", os.path.splitext(os.path.basename(LIB))[0])
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(build_dir, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("RTB_NVCC_EXTRA", "").split() + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-x", "cu" if s.endswith(".cu") else "c++",
                                        "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    link = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-Xcompiler", "-pthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(link)}\n{r.stdout}")
    if r.returncode:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(build_dir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
