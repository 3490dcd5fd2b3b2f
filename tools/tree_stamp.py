#!/usr/bin/env python
"""Stamp of the kernel sources (sha256 over csrc/*.cu, csrc/*.cuh and include/*.h): identifies the tree an ncu
capture was taken from.  The GPU box has no .git, so `git rev-parse` cannot be used there; this hash is the same
here and there.  bench.py prints capture-derived numbers only when the stamp stored with them equals this one."""
import glob
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tree_stamp() -> str:
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "cwbnwp_letkf_b200", "csrc", "*.cu")) +
                   glob.glob(os.path.join(ROOT, "cwbnwp_letkf_b200", "csrc", "*.cuh")) +
                   glob.glob(os.path.join(ROOT, "include", "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(tree_stamp())
