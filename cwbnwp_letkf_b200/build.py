"""Build libletkf_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the
repository snapshot to the GPU box)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libletkf_b200.so")
SOURCES = ["api.cu", "kdtree_host.cu", "kernels_obs.cu", "kernels_search.cu", "kernels_gram.cu",
           "kernels_eig.cu", "kernels_xform.cu", "kernels_eig32.cu", "kernels_k32.cu"]
HEADERS = [os.path.join(CSRC, "letkf_internal.cuh"),
           os.path.join(HERE, "..", "include", "letkf_b200.h"),
           os.path.join(HERE, "..", "include", "letkf_b200_math.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-O2", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> list:
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _stamp(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources_with_extra() -> list:
    extra = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and f not in SOURCES)
    return SOURCES + extra


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = sources_with_extra()
    incs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hdr_stamp = _stamp(sorted(set(HEADERS + incs)))
    nvcc = _nvcc()
    todo = []
    objs = []
    for src in srcs:
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp([path]) + hdr_stamp
        objs.append(obj)
        old = open(stamp_file).read() if os.path.exists(stamp_file) else ""
        if force or old != stamp or not os.path.exists(obj):
            todo.append((path, obj, stamp_file, stamp))

    def compile_one(item):
        path, obj, stamp_file, stamp = item
        cmd = [nvcc] + _host_cxx() + NVCC_FLAGS + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log = obj + ".log"
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (path, r.stdout[-6000:]))
        with open(stamp_file, "w") as f:
            f.write(stamp)
        return r.stdout

    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            outs = list(ex.map(compile_one, todo))
        if verbose:
            for o in outs:
                print(o)
    if todo or not os.path.exists(LIB):
        cmd = [nvcc] + _host_cxx() + ["-shared", "-o", LIB] + objs + \
              ["-Xcompiler", "-fopenmp", "-lgomp", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout[-4000:])
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose=True))
