"""Host-side mirror of the reference's hot-path interface over the C ABI (include/letkf_b200.h).

The reference is Fortran; its seam for this path is a set of module procedures, restated here
with the same names and argument meaning so the parity tests read like calls into the
reference:

=====================================  =========================================================
reference (file:line)                   here
=====================================  =========================================================
set_optimal_workspace_for_eigen         ``LetkfB200(nmember, real64)``
  (module_eigen.f90:16) +
  set_ensemble_constants (param:126)
gts%distribute / rad%distribute ->      ``set_obs(ObsSet)``
  platform(:), radarobs(:) complete
  (module_letkf_core.f90:50)
build_tree + get_lz                     ``get_lz(cfg, xyz_grid)``
  (module_localization.f90:35-331)
letkf_yoyb (core:300-595)               ``letkf_yoyb(cfg, xyz_grid)``
letkf_solve internals (core:649-679)    ``letkf_weights(cfg, xyz_grid, xb)``
loop body of letkf_driver               ``analyze(cfg, xyz_grid, var)`` (host arrays) /
  (core:209-240) [+ letkf_tune_q]       ``analyze_dev`` (torch CUDA tensors)
letkf_tune_q (core:702-733)             ``tune_q(var)``
?syevd in inverse_matrix (eig:49/66)    ``syevd_batched(A)``
destroy_eigen_array (eig:110)           ``finalize()``
=====================================  =========================================================

This module is plumbing only: every computation happens in libletkf_b200.so on the GPU.
There is no CPU path; without the library or a CUDA device the constructor raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Tuple

import numpy as np

from . import config as C

_LIB = None


class LetkfError(RuntimeError):
    pass


class Stats(ctypes.Structure):
    _fields_ = [("npts", ctypes.c_int64), ("npts_analysed", ctypes.c_int64), ("rows", ctypes.c_int64),
                ("units", ctypes.c_int64), ("ntrees", ctypes.c_int32), ("max_sweeps", ctypes.c_int32),
                ("ms_tree", ctypes.c_float), ("ms_search", ctypes.c_float), ("ms_gram", ctypes.c_float),
                ("ms_eigen", ctypes.c_float), ("ms_transform", ctypes.c_float), ("ms_total", ctypes.c_float),
                ("sweeps_sum", ctypes.c_int64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


EXPORTS = ["letkf_b200_init", "letkf_b200_finalize", "letkf_b200_last_error", "letkf_b200_version",
           "letkf_b200_set_obs", "letkf_b200_set_obs_dev", "letkf_b200_clear_obs", "letkf_b200_analyze",
           "letkf_b200_analyze_dev", "letkf_b200_tune_q", "letkf_b200_tune_q_dev", "letkf_b200_search",
           "letkf_b200_yoyb", "letkf_b200_weights", "letkf_b200_syevd_batched",
           "letkf_b200_syevd_batched_dev", "letkf_b200_fma_peak", "letkf_b200_launch_count",
           "letkf_b200_stream", "letkf_b200_set_chunk", "letkf_b200_set_levels",
           "letkf_b200_selftest_host_search", "letkf_b200_selftest_pole_table"]


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libletkf_b200.so")


def load_library():
    """dlopen libletkf_b200.so.  Raises if it has not been built: there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LetkfError(f"{path} is missing: run __graft_entry__.build() (nvcc, sm_100a); "
                         "this package has no CPU implementation")
    L = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.letkf_b200_last_error.restype = ctypes.c_char_p
    L.letkf_b200_init.argtypes = [ctypes.POINTER(vp), i32, i32, i32]
    L.letkf_b200_finalize.argtypes = [vp]
    L.letkf_b200_set_obs.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.letkf_b200_set_obs_dev.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.letkf_b200_clear_obs.argtypes = [vp]
    L.letkf_b200_analyze.argtypes = [vp, vp, i64, vp, i32, vp, vp]
    L.letkf_b200_analyze_dev.argtypes = [vp, vp, i64, vp, i32, vp, vp]
    L.letkf_b200_tune_q.argtypes = [vp, i64, vp]
    L.letkf_b200_tune_q_dev.argtypes = [vp, i64, vp]
    L.letkf_b200_search.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.letkf_b200_yoyb.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.letkf_b200_weights.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.letkf_b200_syevd_batched.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp]
    L.letkf_b200_syevd_batched_dev.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp]
    L.letkf_b200_fma_peak.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]
    L.letkf_b200_launch_count.argtypes = [vp]
    L.letkf_b200_launch_count.restype = i64
    L.letkf_b200_stream.argtypes = [vp]
    L.letkf_b200_stream.restype = vp
    L.letkf_b200_set_chunk.argtypes = [vp, i64]
    L.letkf_b200_set_levels.argtypes = [vp, i32]
    L.letkf_b200_selftest_host_search.argtypes = [i32, vp, ctypes.c_float, ctypes.c_float, i64, vp, i32, vp,
                                                  vp, vp, vp, vp]
    L.letkf_b200_selftest_pole_table.argtypes = [vp, i32]
    _LIB = L
    return L


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype)


def _sync_producer(*tensors):
    """The *_dev entry points launch on the library's own stream (letkf_b200_stream): whatever produced the
    caller's device buffers -- torch kernels, an NCCL collective -- must have completed first (contract stated in
    include/letkf_b200.h).  The library's calls return only when their own work is complete."""
    for t in tensors:
        if t is not None and hasattr(t, "is_cuda") and t.is_cuda:
            import torch
            torch.cuda.current_stream(t.device).synchronize()
            return


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(a.data_ptr())  # torch tensor


class LetkfB200:
    """One GPU's local-analysis engine (one per process / rank, like one MPI rank of the reference)."""

    fuses_tune_q = True   # analyze() applies letkf_tune_q itself when cfg.tune_q is set (driver.py)

    def __init__(self, nmember: int, real64: bool = True, device: int = 0):
        self.L = load_library()
        self.k, self.real64, self.device = int(nmember), bool(real64), int(device)
        h = ctypes.c_void_p()
        if self.L.letkf_b200_init(ctypes.byref(h), self.k, int(self.real64), self.device):
            raise LetkfError(self.L.letkf_b200_last_error().decode())
        self.h = h
        self.last_stats: Optional[Stats] = None

    # -- life cycle ---------------------------------------------------------------------------
    def finalize(self):
        if getattr(self, "h", None):
            self.L.letkf_b200_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise LetkfError(self.L.letkf_b200_last_error().decode())

    @property
    def launch_count(self) -> int:
        return int(self.L.letkf_b200_launch_count(self.h))

    @property
    def stream_ptr(self) -> int:
        return int(self.L.letkf_b200_stream(self.h) or 0)

    def set_levels(self, nz: int):
        """Declare npts = ncol*nz (levels slowest) so that 2-D localised variables share weights per column."""
        self._chk(self.L.letkf_b200_set_levels(self.h, int(nz)))

    def set_chunk(self, n: int):
        self._chk(self.L.letkf_b200_set_chunk(self.h, int(n)))

    # -- observations ---------------------------------------------------------------------------
    def set_obs(self, o):
        """``o``: ObsSet-like with the reference layouts (see synthetic.py)."""
        xyz, obs, hdxb = _np(o.xyz, np.float32), _np(o.obs, np.float32), _np(o.hdxb, np.float32)
        err = None if o.error is None else _np(o.error, np.float32)
        qc = None if o.qc is None else _np(o.qc, np.int32)
        n = xyz.shape[0]
        assert hdxb.shape == (self.k, n, o.nvar), (hdxb.shape, (self.k, n, o.nvar))
        self._chk(self.L.letkf_b200_set_obs(self.h, o.family, o.type, n, o.nvar, _ptr(xyz), _ptr(obs),
                                            _ptr(err), _ptr(hdxb), _ptr(qc)))

    def set_obs_dev(self, family, type_, n, nvar, xyz, obs, error, hdxb, qc):
        """Device-resident variant (torch CUDA tensors, same layouts)."""
        _sync_producer(xyz, obs, error, hdxb, qc)
        self._chk(self.L.letkf_b200_set_obs_dev(self.h, family, type_, n, nvar, _ptr(xyz), _ptr(obs),
                                                _ptr(error), _ptr(hdxb), _ptr(qc)))

    def clear_obs(self):
        self._chk(self.L.letkf_b200_clear_obs(self.h))

    # -- the hot path -----------------------------------------------------------------------------
    def analyze(self, cfg: C.VarConfig, xyz_grid: np.ndarray, var: np.ndarray) -> Stats:
        """Loop body of letkf_driver for all points.  ``var``: (k,npts) or (nfields,k,npts) float32,
        C-contiguous host array (pinned or pageable), updated in place."""
        xyz = _np(xyz_grid, np.float32)
        assert var.dtype == np.float32 and var.flags.c_contiguous
        npts = xyz.shape[0]
        nfields = 1 if var.ndim == 2 else var.shape[0]
        assert var.shape[-2:] == (self.k, npts)
        cc, st = C.to_c(cfg), Stats()
        self._chk(self.L.letkf_b200_analyze(self.h, ctypes.byref(cc), npts, _ptr(xyz), nfields, _ptr(var),
                                            ctypes.byref(st)))
        self.last_stats = st
        return st

    def analyze_dev(self, cfg: C.VarConfig, xyz_grid, var) -> Stats:
        """Same with torch CUDA tensors already resident in HBM."""
        assert xyz_grid.is_cuda and var.is_cuda and xyz_grid.is_contiguous() and var.is_contiguous()
        npts = xyz_grid.shape[0]
        nfields = 1 if var.dim() == 2 else var.shape[0]
        assert tuple(var.shape[-2:]) == (self.k, npts)
        cc, st = C.to_c(cfg), Stats()
        _sync_producer(xyz_grid, var)
        self._chk(self.L.letkf_b200_analyze_dev(self.h, ctypes.byref(cc), npts, _ptr(xyz_grid), nfields,
                                                _ptr(var), ctypes.byref(st)))
        self.last_stats = st
        return st

    def analyze_ptr(self, ccfg, npts: int, xyz_ptr: int, nfields: int, var_ptr: int, dev: bool) -> Stats:
        """Raw-pointer call (bench.py: pinned host buffers or device buffers, no wrapper overhead)."""
        st = Stats()
        fn = self.L.letkf_b200_analyze_dev if dev else self.L.letkf_b200_analyze
        self._chk(fn(self.h, ctypes.byref(ccfg), npts, ctypes.c_void_p(xyz_ptr), nfields,
                     ctypes.c_void_p(var_ptr), ctypes.byref(st)))
        self.last_stats = st
        return st

    def tune_q(self, var: np.ndarray):
        assert var.dtype == np.float32 and var.flags.c_contiguous and var.shape[0] == self.k
        self._chk(self.L.letkf_b200_tune_q(self.h, var.shape[1], _ptr(var)))

    # -- stage-level (parity) ---------------------------------------------------------------------
    def get_lz(self, cfg: C.VarConfig, xyz_grid: np.ndarray):
        """build_tree + get_lz for every point.  Returns a list, one entry per tree in the
        reference's visiting order: (family, type, count[npts], idx[npts,max_lz], r2[npts,max_lz])."""
        xyz = _np(xyz_grid, np.float32)
        npts = xyz.shape[0]
        cc = C.to_c(cfg)
        nt = ctypes.c_int32(0)
        fam, typ, stride = (np.zeros(C.MAX_TYPES, np.int32) for _ in range(3))
        self._chk(self.L.letkf_b200_search(self.h, ctypes.byref(cc), npts, _ptr(xyz), ctypes.byref(nt),
                                           _ptr(fam), _ptr(typ), _ptr(stride), None, None, None))
        n = nt.value
        tot = int(sum(int(stride[t]) * npts for t in range(n)))
        count = np.zeros((max(n, 1), npts), np.int32)
        idx = np.zeros(max(tot, 1), np.int32)
        r2 = np.zeros(max(tot, 1), np.float32)
        if n and npts:
            self._chk(self.L.letkf_b200_search(self.h, ctypes.byref(cc), npts, _ptr(xyz), ctypes.byref(nt),
                                               _ptr(fam), _ptr(typ), _ptr(stride), _ptr(count), _ptr(idx),
                                               _ptr(r2)))
        out, off = [], 0
        for t in range(n):
            s = int(stride[t])
            out.append((int(fam[t]), int(typ[t]), count[t].copy(),
                        idx[off:off + npts * s].reshape(npts, s).copy(),
                        r2[off:off + npts * s].reshape(npts, s).copy()))
            off += npts * s
        return out

    def letkf_yoyb(self, cfg: C.VarConfig, xyz_grid: np.ndarray):
        """Returns row_offset[npts+1], yo[rows], yb[rows,k] (== Fortran yb(k,rows))."""
        xyz = _np(xyz_grid, np.float32)
        npts = xyz.shape[0]
        cc = C.to_c(cfg)
        off = np.zeros(npts + 1, np.int64)
        self._chk(self.L.letkf_b200_yoyb(self.h, ctypes.byref(cc), npts, _ptr(xyz), _ptr(off), None, None))
        rows = int(off[-1])
        yo = np.zeros(max(rows, 1), np.float32)
        yb = np.zeros((max(rows, 1), self.k), np.float32)
        if rows:
            self._chk(self.L.letkf_b200_yoyb(self.h, ctypes.byref(cc), npts, _ptr(xyz), _ptr(off), _ptr(yo),
                                             _ptr(yb)))
        return off, yo[:rows], yb[:rows]

    def letkf_weights(self, cfg: C.VarConfig, xyz_grid: np.ndarray, xb: Optional[np.ndarray] = None,
                      want_Wa: bool = True):
        """p[npts], wbar[npts,k], Wa[npts,k,k] (Wa[n].T is the Fortran matrix; it is symmetric),
        xa_raw[npts,k] (if xb[(k,npts)] is given)."""
        xyz = _np(xyz_grid, np.float32)
        npts, k = xyz.shape[0], self.k
        cc = C.to_c(cfg)
        p = np.zeros(npts, np.int32)
        wbar = np.zeros((npts, k), np.float64)
        Wa = np.zeros((npts, k, k), np.float64) if want_Wa else None
        raw = None
        xbc = None
        if xb is not None:
            xbc = _np(xb, np.float32)
            assert xbc.shape == (k, npts)
            raw = np.zeros((npts, k), np.float64)
        self._chk(self.L.letkf_b200_weights(self.h, ctypes.byref(cc), npts, _ptr(xyz), _ptr(xbc), _ptr(p),
                                            _ptr(wbar), _ptr(Wa), _ptr(raw)))
        return p, wbar, Wa, raw

    def syevd_batched(self, A: np.ndarray):
        """A[b,k,k] (each matrix column-major, lower triangle referenced; symmetric input makes the
        layout moot).  Returns W[b,k] ascending, V[b,k,k] with V[b][j] = j-th eigenvector, sweeps."""
        A = np.ascontiguousarray(A)
        assert A.dtype in (np.float32, np.float64) and A.ndim == 3 and A.shape[1] == A.shape[2]
        b, k, _ = A.shape
        W = np.zeros((b, k), A.dtype)
        V = np.zeros_like(A)
        sw = ctypes.c_int32(0)
        self._chk(self.L.letkf_b200_syevd_batched(self.h, k, b, int(A.dtype == np.float64), _ptr(A), _ptr(W),
                                                  _ptr(V), ctypes.byref(sw)))
        return W, V, sw.value

    def syevd_batched_dev(self, A, W, V) -> int:
        """torch CUDA tensors; A[b,k,k], W[b,k], V[b,k,k] of one dtype."""
        b, k, _ = A.shape
        import torch
        sw = ctypes.c_int32(0)
        _sync_producer(A)
        self._chk(self.L.letkf_b200_syevd_batched_dev(self.h, k, b, int(A.dtype == torch.float64), _ptr(A),
                                                      _ptr(W), _ptr(V), ctypes.byref(sw)))
        return sw.value

    def fma_peak(self, kind: int = 0) -> float:
        t = ctypes.c_double(0)
        self._chk(self.L.letkf_b200_fma_peak(self.h, kind, ctypes.byref(t)))
        return t.value


def selftest_host_search(obs_xyz: np.ndarray, hclr: float, vclr: float, xyz_grid: np.ndarray, max_lz_pts: int):
    """Host-only self-test (no GPU): the pipeline's tree builder + the search routine compiled for
    the host.  Returns ind[n], nnodes, count[nq], idx[nq,max_lz], r2[nq,max_lz]."""
    L = load_library()
    ox, q = _np(obs_xyz, np.float32), _np(xyz_grid, np.float32)
    n, nq = ox.shape[0], q.shape[0]
    ind = np.zeros(n, np.int32)
    nn = ctypes.c_int32(0)
    cnt = np.zeros(max(nq, 1), np.int32)
    idx = np.zeros((max(nq, 1), max_lz_pts), np.int32)
    r2 = np.zeros((max(nq, 1), max_lz_pts), np.float32)
    if L.letkf_b200_selftest_host_search(n, _ptr(ox), hclr, vclr, nq, _ptr(q), max_lz_pts, _ptr(ind),
                                         ctypes.byref(nn), _ptr(cnt), _ptr(idx), _ptr(r2)):
        raise LetkfError(L.letkf_b200_last_error().decode())
    return ind, nn.value, cnt[:nq], idx[:nq], r2[:nq]
