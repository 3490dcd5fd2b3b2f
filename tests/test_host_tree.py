"""CPU tier: the C-ABI library loads, exports every declared symbol, and its host-side k-d tree
builder + search walk (the same __host__ __device__ routine the CUDA kernel runs) agree with the
oracle's kdtree2 restatement bit for bit -- permutation, hit sets, hit ORDER and distances."""
import os
import re

import numpy as np
import pytest

from cwbnwp_letkf_b200 import host as H
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = H.load_library()
    hdr = open(os.path.join(ROOT, "include", "letkf_b200.h")).read()
    declared = set(re.findall(r"\b(letkf_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(H.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.letkf_b200_version() >= 100


def test_ctypes_bindings_match_the_header_arity():
    """Every prototype in include/letkf_b200.h has as many parameters as the ctypes binding in host.py
    (an ABI drift between the header, the library and the Python mirror shows up here, without a GPU), and
    the Fortran ISO_C_BINDING module declares an interface for every call the driver shim makes."""
    L = H.load_library()
    hdr = open(os.path.join(ROOT, "include", "letkf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b(letkf_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) >= 20
    checked = 0
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        fn = getattr(L, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, (name, n, len(fn.argtypes))
            checked += 1
    assert checked >= 15
    f90 = open(os.path.join(ROOT, "cwbnwp_letkf_b200", "fortran", "letkf_b200_mod.f90")).read().lower()
    for name in ("letkf_b200_init", "letkf_b200_finalize", "letkf_b200_set_obs", "letkf_b200_analyze",
                 "letkf_b200_last_error"):
        assert re.search(r'bind\(c,\s*name\s*=\s*"%s"\)' % name, f90), name


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(H.LetkfError, match="no CUDA device|no CPU path"):
        H.LetkfB200(8)


@pytest.mark.parametrize("n,dim,seed", [(1, 2, 0), (13, 3, 1), (14, 3, 2), (200, 2, 3), (5000, 3, 4),
                                         (60000, 3, 5)])
def test_host_tree_and_walk_match_oracle(n, dim, seed):
    rng = np.random.default_rng(seed)
    hclr, vclr = 8.0, (2.0 if dim == 3 else -1.0)
    obs = np.empty((n, 3), np.float32)
    obs[:, :2] = rng.uniform(-40e3, 40e3, (n, 2))
    obs[:, 2] = rng.uniform(300, 12000, n)
    obs[::5, 0] = np.round(obs[::5, 0], -3)            # ties in the split coordinate
    nq = 300
    q = np.empty((nq, 3), np.float32)
    q[:, :2] = rng.uniform(-45e3, 45e3, (nq, 2))
    q[:, 2] = rng.uniform(0, 14000, nq)
    for max_lz in (7, 300):
        ind, nnodes, cnt, idx, r2 = H.selftest_host_search(obs, hclr, vclr, q, max_lz)
        # oracle: same normalisation in real32 (loc:149-157,243-246)
        hinv = np.float32(1.0) / (np.float32(hclr) * np.float32(1e3))
        data = obs.copy()
        data[:, :2] *= hinv
        if dim == 3:
            vinv = np.float32(1.0) / (np.float32(vclr) * np.float32(1e3))
            data[:, 2] *= vinv
        else:
            data[:, 2] = -1.0
        t = O.KdTree(data, dim)
        d = t.dump()
        assert np.array_equal(d["ind"], ind)
        assert len(d["cut_dim"]) == nnodes
        for i in range(nq):
            qq = q[i].copy()
            qq[:2] *= hinv
            if dim == 3:
                qq[2] *= vinv
            oi, od, _ = t.r_nearest(qq[:dim], O.search_r2(), max_lz)
            assert cnt[i] == len(oi)
            assert np.array_equal(idx[i, :cnt[i]], oi)
            assert np.array_equal(r2[i, :cnt[i]].view(np.int32), od.view(np.int32))
