"""ctypes front end of the CPU oracle (oracle/letkf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED by reference tests (none exist); see letkf_oracle.h.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libletkf_oracle.so")
MAX_SLOTS = 5


class _CType(ctypes.Structure):
    _fields_ = [("family", ctypes.c_int32), ("type", ctypes.c_int32), ("use_it", ctypes.c_int32),
                ("max_lz_pts", ctypes.c_int32), ("hclr", ctypes.c_float), ("vclr", ctypes.c_float),
                ("nvar", ctypes.c_int32), ("is_assim", ctypes.c_int32 * MAX_SLOTS),
                ("err_muti", ctypes.c_float * MAX_SLOTS), ("err_rej", ctypes.c_float * MAX_SLOTS)]


class _CVar(ctypes.Structure):
    _fields_ = [("ntypes", ctypes.c_int32), ("weight_function", ctypes.c_int32),
                ("norain_value", ctypes.c_float), ("multi_infl", ctypes.c_float),
                ("use_rtpp", ctypes.c_int32), ("rtpp_alpha", ctypes.c_float),
                ("use_rtps", ctypes.c_int32), ("rtps_alpha", ctypes.c_float),
                ("types", _CType * 16)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ + scipy's OpenBLAS)."""
    src = [os.path.join(_HERE, f) for f in ("letkf_oracle.cpp", "letkf_oracle.h")] + \
          [os.path.join(_HERE, "..", "include", "letkf_b200_math.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.or_create.restype = ctypes.c_void_p
        L.or_create.argtypes = [ctypes.c_int, ctypes.c_int]
        L.or_destroy.argtypes = [ctypes.c_void_p]
        L.or_last_error.restype = ctypes.c_char_p
        L.or_kd_create.restype = ctypes.c_void_p
        L.or_kd_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.or_kd_destroy.argtypes = [ctypes.c_void_p]
        for f in ("or_gc1999", "or_search_r2"):
            getattr(L, f).restype = ctypes.c_float
        for f in ("or_gaspari_cohn", "or_expf"):
            getattr(L, f).restype = ctypes.c_float
            getattr(L, f).argtypes = [ctypes.c_float]
        L.or_error_inv.restype = ctypes.c_float
        L.or_error_inv.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def to_c(cfg) -> _CVar:
    """Any object shaped like cwbnwp_letkf_b200.config.VarConfig."""
    c = _CVar()
    c.ntypes = len(cfg.types)
    c.weight_function = int(cfg.weight_function)
    c.norain_value = cfg.norain_value
    c.multi_infl = cfg.multi_infl
    c.use_rtpp, c.rtpp_alpha = int(cfg.use_rtpp), cfg.rtpp_alpha
    c.use_rtps, c.rtps_alpha = int(cfg.use_rtps), cfg.rtps_alpha
    for i, t in enumerate(cfg.types):
        ct = c.types[i]
        ct.family, ct.type, ct.use_it = t.family, t.type, int(t.use_it)
        ct.max_lz_pts, ct.hclr, ct.vclr, ct.nvar = t.max_lz_pts, t.hclr, t.vclr, t.nvar
        for s in range(MAX_SLOTS):
            ct.is_assim[s] = int(t.is_assim[s])
            ct.err_muti[s] = t.err_muti[s]
            ct.err_rej[s] = t.err_rej[s]
    return c


class KdTree:
    """kdtree2 (module_kdtree2.f90) on data[(n,3)] using the first ``dim`` coordinates."""

    def __init__(self, data: np.ndarray, dim: int):
        self.data = np.ascontiguousarray(data, np.float32)
        assert self.data.ndim == 2 and self.data.shape[1] == 3
        self.n, self.dim = self.data.shape[0], dim
        self.h = lib().or_kd_create(_p(self.data), self.n, dim)

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_kd_destroy(ctypes.c_void_p(self.h))
            self.h = None

    def r_nearest(self, qv, r2: float, nalloc: int):
        q = np.ascontiguousarray(qv, np.float32)
        idx = np.zeros(nalloc, np.int32)
        dis = np.zeros(nalloc, np.float32)
        tot = ctypes.c_int(0)
        nf = lib().or_kd_r_nearest(ctypes.c_void_p(self.h), _p(q), ctypes.c_float(r2), nalloc,
                                   _p(idx), _p(dis), ctypes.byref(tot))
        return idx[:nf].copy(), dis[:nf].copy(), tot.value

    def brute(self, qv, r2: float):
        q = np.ascontiguousarray(qv, np.float32)
        idx = np.zeros(self.n, np.int32)
        dis = np.zeros(self.n, np.float32)
        nf = lib().or_kd_brute(ctypes.c_void_p(self.h), _p(q), ctypes.c_float(r2), self.n, _p(idx), _p(dis))
        return idx[:nf].copy(), dis[:nf].copy()

    def dump(self):
        nn = lib().or_kd_num_nodes(ctypes.c_void_p(self.h))
        d = dict(cut_dim=np.zeros(nn, np.int32), cut_val=np.zeros(nn, np.float32),
                 cut_l=np.zeros(nn, np.float32), cut_r=np.zeros(nn, np.float32),
                 l=np.zeros(nn, np.int32), u=np.zeros(nn, np.int32), left=np.zeros(nn, np.int32),
                 right=np.zeros(nn, np.int32), box=np.zeros((nn, 3, 2), np.float32),
                 ind=np.zeros(self.n, np.int32))
        lib().or_kd_dump(ctypes.c_void_p(self.h), *[_p(d[k]) for k in
                         ("cut_dim", "cut_val", "cut_l", "cut_r", "l", "u", "left", "right", "box", "ind")])
        return d


class Oracle:
    """The reference hot path on the CPU: build_tree / get_lz / letkf_yoyb / letkf_solve and the
    grid-point loop of letkf_driver."""

    def __init__(self, nmember: int, real64: bool = True):
        self.k, self.real64 = nmember, real64
        self.h = lib().or_create(nmember, int(real64))
        self._keep = []
        self.ntrees = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_destroy(ctypes.c_void_p(self.h))
            self.h = None

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError(lib().or_last_error().decode())
        return rc

    def set_obs(self, o):
        """o: cwbnwp_letkf_b200.synthetic.ObsSet-shaped object."""
        arrs = [np.ascontiguousarray(o.xyz, np.float32), np.ascontiguousarray(o.obs, np.float32),
                None if o.error is None else np.ascontiguousarray(o.error, np.float32),
                np.ascontiguousarray(o.hdxb, np.float32),
                None if o.qc is None else np.ascontiguousarray(o.qc, np.int32)]
        self._keep.append(arrs)
        self._chk(lib().or_set_obs(ctypes.c_void_p(self.h), o.family, o.type, arrs[0].shape[0], o.nvar,
                                   *[_p(a) for a in arrs]))

    def build_tree(self, cfg) -> int:
        self.ccfg = to_c(cfg)
        self.ntrees = self._chk(lib().or_build_trees(ctypes.c_void_p(self.h), ctypes.byref(self.ccfg)))
        self.stride = max([t.max_lz_pts for t in cfg.types] + [1])
        return self.ntrees

    def destroy_tree(self):
        lib().or_destroy_trees(ctypes.c_void_p(self.h))
        self.ntrees = 0

    def get_lz(self, xyz) -> List[Tuple[int, int, np.ndarray, np.ndarray]]:
        nt, st = self.ntrees, self.stride
        fam, typ, n = (np.zeros(max(nt, 1), np.int32) for _ in range(3))
        idx = np.zeros((max(nt, 1), st), np.int32)
        r2 = np.zeros((max(nt, 1), st), np.float32)
        q = np.ascontiguousarray(xyz, np.float32)
        self._chk(lib().or_get_lz(ctypes.c_void_p(self.h), ctypes.byref(self.ccfg), _p(q), st, _p(fam),
                                  _p(typ), _p(n), _p(idx), _p(r2)))
        self._last = (fam, typ, n, idx, r2)
        return [(int(fam[t]), int(typ[t]), idx[t, :n[t]].copy(), r2[t, :n[t]].copy()) for t in range(nt)]

    def letkf_yoyb(self, xyz, pmax: int = 8192):
        self.get_lz(xyz)
        fam, typ, n, idx, r2 = self._last
        yo = np.zeros(pmax, np.float32)
        yb = np.zeros((pmax, self.k), np.float32)
        p = self._chk(lib().or_yoyb(ctypes.c_void_p(self.h), ctypes.byref(self.ccfg), self.ntrees,
                                    self.stride, _p(fam), _p(typ), _p(n), _p(idx), _p(r2), pmax,
                                    _p(yo), _p(yb)))
        return yo[:p].copy(), yb[:p].copy()      # yb[p,k] == Fortran yb(k,p)

    def letkf_solve(self, xb, yo, yb, inflat, use_rtpp=False, rtpp_alpha=0.0, use_rtps=False,
                    rtps_alpha=0.0):
        k = self.k
        xb = np.ascontiguousarray(xb, np.float32)
        yo = np.ascontiguousarray(yo, np.float32)
        yb = np.ascontiguousarray(yb, np.float32)
        xa = np.zeros(k, np.float32)
        wbar, Wa, raw = np.zeros(k), np.zeros((k, k)), np.zeros(k)
        self._chk(lib().or_solve(ctypes.c_void_p(self.h), _p(xb), yo.shape[0], _p(yo), _p(yb),
                                 ctypes.c_float(inflat), int(use_rtpp), ctypes.c_float(rtpp_alpha),
                                 int(use_rtps), ctypes.c_float(rtps_alpha), _p(xa), _p(wbar), _p(Wa),
                                 _p(raw)))
        return xa, wbar, Wa, raw

    def analyze(self, cfg, xyz_grid, var, nthreads: int = 1):
        """var: (nfields,k,npts) or (k,npts) float32, updated in place."""
        xyz_grid = np.ascontiguousarray(xyz_grid, np.float32)
        assert var.dtype == np.float32 and var.flags.c_contiguous
        npts = xyz_grid.shape[0]
        nfields = 1 if var.ndim == 2 else var.shape[0]
        self.build_tree(cfg)
        npo, rows = ctypes.c_int64(0), ctypes.c_int64(0)
        self._chk(lib().or_analyze(ctypes.c_void_p(self.h), ctypes.byref(self.ccfg),
                                   ctypes.c_int64(npts), _p(xyz_grid), nfields, _p(var), nthreads,
                                   ctypes.byref(npo), ctypes.byref(rows)))
        self.destroy_tree()
        return npo.value, rows.value


def tune_q(var: np.ndarray):
    assert var.dtype == np.float32 and var.flags.c_contiguous and var.ndim == 2
    lib().or_tune_q(var.shape[0], ctypes.c_int64(var.shape[1]), _p(var))


def syevd_batch(A: np.ndarray, nthreads: int = 1):
    """LAPACK ?syevd('V','L') over A[b,k,k] (column-major per matrix == numpy A[b].T)."""
    A = np.ascontiguousarray(A)
    b, k, _ = A.shape
    real64 = A.dtype == np.float64
    W = np.zeros((b, k), A.dtype)
    V = np.zeros_like(A)
    rc = lib().or_syevd_batch(k, ctypes.c_int64(b), int(real64), _p(A), _p(W), _p(V), nthreads)
    if rc < 0:
        raise RuntimeError(lib().or_last_error().decode())
    return W, V


def gc1999():
    return lib().or_gc1999()


def search_r2():
    return lib().or_search_r2()
