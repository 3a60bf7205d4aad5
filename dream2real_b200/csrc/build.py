#!/usr/bin/env python3
"""Build libd2r_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python dream2real_b200/csrc/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libd2r_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]
# arithmetic half of the reference's --use_fast_math (approximate div/sqrt, flush-to-zero); the
# transcendental intrinsics are spelled out in the sources that mirror reference maths.
FAST = ["-ftz=true", "-prec-div=false", "-prec-sqrt=false"]
SOURCES = {
    "d2r_model.cu": FAST,
    "d2r_march.cu": FAST,
    "d2r_phys.cu": FAST,
    "d2r_post.cu": [],
    "d2r_gemm.cu": [],
    "d2r_clip.cu": [],
}


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(os.path.dirname(PKG), "include", "d2r_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        os.makedirs(os.path.dirname(o), exist_ok=True)
        cmd = [nvcc()] + ARCH + COMMON + SOURCES[os.path.basename(s)] + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"---- {os.path.basename(s)} ----\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc()] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
