"""Builds librsa_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

Run as `python rectified-spaattn_b200/build_native.py [-v] [--force]` or through __graft_entry__.build().
The .so lands in rectified-spaattn_b200/rsa_b200/ (git-ignored, shipped to the GPU box by gpurun)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "rsa_b200")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "librsa_b200.so")

SOURCES = ["rsa_api.cu", "gilbert.cc", "permute.cu", "pool_stats.cu", "block_scores.cu", "block_select.cu",
           "rect_c.cu", "attn_tc5.cu", "host_call.cu", "peer.cu"]
# test infrastructure, NOT part of the product library: the mma.sync cross-check implementation of kernel 4
XCHECK_DIR = os.path.join(REPO, "tests", "xcheck")
XCHECK_SOURCES = ["attn_mma.cu", "xcheck_api.cu"]
XCHECK_LIB = os.path.join(XCHECK_DIR, "librsa_xcheck.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(REPO, "include"),
                "-I", CSRC]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(REPO, "include", "rsa.h"))
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, os.path.splitext(s)[0] + ".o")
        if force or _newer(src, obj) or any(_newer(h, obj) for h in hdrs):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write("## " + cmd[-3] + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ_DIR, os.path.splitext(s)[0] + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        run([NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-lcudart"])
    return LIB


def build_xcheck(force=False):
    """tests/xcheck/librsa_xcheck.so (the tests' independent implementation of kernel 4; see tests/xcheck/__init__.py)."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, "rsa_common.cuh"), os.path.join(REPO, "include", "rsa.h")]
    objs, rebuilt = [], False
    for s in XCHECK_SOURCES:
        src = os.path.join(XCHECK_DIR, s)
        obj = os.path.join(OBJ_DIR, "xcheck_" + os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in hdrs):
            r = subprocess.run([NVCC] + FLAGS + ["-x", "cu", "-c", src, "-o", obj], capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on " + s)
            rebuilt = True
    if rebuilt or not os.path.exists(XCHECK_LIB):
        r = subprocess.run([NVCC, "-shared", "-o", XCHECK_LIB] + objs + ARCH + ["-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("linking librsa_xcheck.so failed")
    return XCHECK_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_xcheck(force="--force" in sys.argv))
