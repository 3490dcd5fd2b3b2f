"""Synthetic WRF-shaped ensembles and observation sets.

The reference ships no input data (its ``../input/*`` files are absent), so every test and
benchmark runs on these seeded generators.  Array layouts are exactly those of the reference
containers, in Fortran (column-major) order:

* GTS platform ``t`` (module_gts_omboma.f90:13-22): ``xyz[3,n]`` metres, ``obs[nvar,n]``,
  ``error[nvar,n]``, ``hdxb[nvar,n,0:k-1]`` (H(x) itself), ``qc[nvar,n,0:k-1]``;
* radar type ``t`` (module_radar.f90:13-16): ``xyz[3,n]``, ``obs[n]``, ``hdxb[n,0:k-1]``;
* a field ``var(loc_nx,loc_ny,nz,0:k-1)`` (module_letkf_core.f90:85): member slowest, so as a
  numpy C array it is ``var[k, npts]`` with point index ``i + nx*(j + ny*l)``.

numpy arrays returned here are C-ordered with the axes REVERSED relative to the Fortran
declaration, which is the same memory.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np

from . import config as C


@dataclass
class ObsSet:
    family: int
    type: int
    nvar: int
    xyz: np.ndarray            # (n,3) float32  == Fortran xyz(3,n)
    obs: np.ndarray            # (n,nvar) float32 == Fortran obs(nvar,n)   [radar: (n,1)]
    hdxb: np.ndarray           # (k,n,nvar) float32 == Fortran hdxb(nvar,n,0:k-1)
    error: np.ndarray | None = None   # (n,nvar) float32, GTS only
    qc: np.ndarray | None = None      # (k,n,nvar) int32, GTS only

    @property
    def n(self) -> int:
        return int(self.xyz.shape[0])


@dataclass
class Scenario:
    name: str
    nx: int
    ny: int
    nz: int
    k: int
    dx: float
    xyz_grid: np.ndarray                  # (npts,3) float32 metres
    obs: Dict[Tuple[int, int], ObsSet] = field(default_factory=dict)
    seed: int = 20261018

    @property
    def npts(self) -> int:
        return self.nx * self.ny * self.nz

    def total_obs_values(self) -> int:
        return int(sum(o.n * o.nvar for o in self.obs.values()))


def _terrain(x, y, span):
    return (750.0 + 750.0 * np.sin(2 * np.pi * x / span) * np.cos(2 * np.pi * y / span)).astype(np.float32)


def make_grid(nx: int, ny: int, nz: int, dx: float) -> np.ndarray:
    """Grid-point coordinates in metres: what ``proj%lonlat_to_xy`` (module_projection.f90:37-50)
    plus ``alt(i,j,k)`` (module_letkf_core.f90:211-214) hand to ``get_lz``.  Point order is
    x fastest, level slowest."""
    xs = ((np.arange(nx, dtype=np.float64) - (nx - 1) / 2) * dx)
    ys = ((np.arange(ny, dtype=np.float64) - (ny - 1) / 2) * dx)
    X, Y = np.meshgrid(xs, ys, indexing="xy")          # (ny,nx)
    span = max(nx, ny) * dx
    ter = _terrain(X, Y, span)                        # 0..1500 m
    lev = (20000.0 * ((np.arange(nz) + 1) / nz) ** 1.5).astype(np.float32)
    out = np.empty((nz, ny, nx, 3), np.float32)
    out[..., 0] = X.astype(np.float32)[None]
    out[..., 1] = Y.astype(np.float32)[None]
    out[..., 2] = ter[None] + lev[:, None, None]
    return out.reshape(-1, 3)


def make_field(rng: np.random.Generator, k: int, xyz_grid: np.ndarray, mean: float, amp: float,
               sigma: float, positive: bool = False) -> np.ndarray:
    """One ensemble field ``var[k,npts]``: smooth mean + member noise."""
    npts = xyz_grid.shape[0]
    smooth = mean + amp * np.sin(xyz_grid[:, 0] / 9.0e4) * np.cos(xyz_grid[:, 1] / 7.0e4) \
        * np.exp(-xyz_grid[:, 2] / 1.2e4)
    var = rng.standard_normal((k, npts), dtype=np.float32) * np.float32(sigma)
    var += smooth.astype(np.float32)[None]
    if positive:
        np.maximum(var, 0.0, out=var)
    return np.ascontiguousarray(var)


def _gts_set(rng, typ, n_sta, nvar, k, xy_span, z_fn, truth_fn):
    xy = rng.uniform(-xy_span / 2, xy_span / 2, size=(n_sta, 2))
    z = z_fn(xy)
    xyz = np.concatenate([xy, z[:, None]], 1).astype(np.float32)
    n = xyz.shape[0]
    err = rng.uniform(0.5, 2.0, size=(n, nvar)).astype(np.float32)
    truth = truth_fn(xyz, nvar).astype(np.float32)
    obs = truth + err * rng.standard_normal((n, nvar), dtype=np.float32)
    hdxb = truth[None] + err[None] * rng.standard_normal((k, n, nvar), dtype=np.float32)
    shifted = rng.random((n, nvar)) < 0.01                    # gross-error candidates
    obs = np.where(shifted, obs + 20.0 * err, obs).astype(np.float32)
    qc = np.zeros((k, n, nvar), np.int32)
    bad = rng.random((n, nvar)) < 0.02
    qc[:, bad] = -88
    return ObsSet(C.GTS, typ, nvar, xyz, obs, np.ascontiguousarray(hdxb.astype(np.float32)), err, qc)


def _truth(xyz, nvar):
    base = np.stack([10 * np.sin(xyz[:, 0] / 1.1e5 + s) + 5 * np.cos(xyz[:, 1] / 0.9e5 - s)
                     + xyz[:, 2] * 1e-3 * (s + 1) for s in range(nvar)], 1)
    return base


def add_gts(sc: Scenario, rng, n_synop=1200, n_metar=400, n_ships=100, n_sound=25, n_lev=40,
            n_gpspw=0):
    span = max(sc.nx, sc.ny) * sc.dx
    sfc = lambda xy: _terrain(xy[:, 0], xy[:, 1], span).astype(np.float64)
    k = sc.k
    if n_synop:
        sc.obs[(C.GTS, C.SYNOP)] = _gts_set(rng, C.SYNOP, n_synop, 5, k, span, sfc, _truth)
    if n_metar:
        sc.obs[(C.GTS, C.METAR)] = _gts_set(rng, C.METAR, n_metar, 5, k, span, sfc, _truth)
    if n_ships:
        sc.obs[(C.GTS, C.SHIPS)] = _gts_set(rng, C.SHIPS, n_ships, 5, k, span, lambda xy: np.zeros(len(xy)), _truth)
    if n_sound:
        base = rng.uniform(-span / 2, span / 2, size=(n_sound, 2))
        levs = np.linspace(200.0, 18000.0, n_lev)
        xy = np.repeat(base, n_lev, 0) + rng.normal(0, 500.0, size=(n_sound * n_lev, 2))
        z = np.tile(levs, n_sound) + _terrain(xy[:, 0], xy[:, 1], span)
        o = _gts_set(rng, C.SOUND, n_sound * n_lev, 4, k, span, lambda q: z, _truth)
        o.xyz[:, :2] = xy.astype(np.float32)
        sc.obs[(C.GTS, C.SOUND)] = o
    if n_gpspw:
        sc.obs[(C.GTS, C.GPSPW)] = _gts_set(rng, C.GPSPW, n_gpspw, 1, k, span, sfc, _truth)


def add_radar(sc: Scenario, rng, n_dbz=600_000, n_vr=400_000, n_sites=8, radius=150e3):
    """Radar-like clusters: uniform in discs around fixed sites, 0.3-12 km high."""
    span_x, span_y = sc.nx * sc.dx, sc.ny * sc.dx
    srng = np.random.default_rng(7)
    sites = np.stack([srng.uniform(-0.33, 0.33, n_sites) * span_x,
                      srng.uniform(-0.33, 0.33, n_sites) * span_y], 1)
    radius = min(radius, 0.45 * min(span_x, span_y))
    k = sc.k

    def positions(n):
        s = rng.integers(0, n_sites, n)
        r = radius * np.sqrt(rng.random(n))
        th = rng.uniform(0, 2 * np.pi, n)
        xyz = np.empty((n, 3), np.float32)
        xyz[:, 0] = sites[s, 0] + r * np.cos(th)
        xyz[:, 1] = sites[s, 1] + r * np.sin(th)
        xyz[:, 2] = rng.uniform(300.0, 12000.0, n)
        return xyz

    if n_dbz:
        xyz = positions(n_dbz)
        norain = rng.random(n_dbz) < 0.5
        obs = np.where(norain, np.float32(-5.0), rng.uniform(5, 55, n_dbz)).astype(np.float32)
        hdxb = obs[None] + 5.0 * rng.standard_normal((k, n_dbz), dtype=np.float32)
        np.maximum(hdxb, np.float32(-5.0), out=hdxb)
        alldry = norain & (rng.random(n_dbz) < 0.3)            # exercises core:507
        hdxb[:, alldry] = -5.0
        sc.obs[(C.RADAR, C.DBZ)] = ObsSet(C.RADAR, C.DBZ, 1, xyz, obs[:, None].copy(),
                                          np.ascontiguousarray(hdxb[:, :, None]))
    if n_vr:
        xyz = positions(n_vr)
        obs = (10.0 * rng.standard_normal(n_vr, dtype=np.float32)).astype(np.float32)
        hdxb = obs[None] + 3.0 * rng.standard_normal((k, n_vr), dtype=np.float32)
        sc.obs[(C.RADAR, C.VR)] = ObsSet(C.RADAR, C.VR, 1, xyz, obs[:, None].copy(),
                                         np.ascontiguousarray(hdxb[:, :, None]))


def scenario_S(k: int = 32, nx: int = 100, ny: int = 100, nz: int = 30, seed: int = 20261018,
               gpspw: int = 0) -> Tuple[Scenario, np.random.Generator]:
    """BASELINE config 1: 100x100x30, dx = 3 km, ~10^4 GTS obs values, radar off."""
    rng = np.random.default_rng(seed)
    sc = Scenario("S", nx, ny, nz, k, 3000.0, make_grid(nx, ny, nz, 3000.0), seed=seed)
    add_gts(sc, rng, n_gpspw=gpspw)
    return sc, rng


def scenario_M(k: int = 32, nx: int = 450, ny: int = 450, nz: int = 50, seed: int = 20261019,
               n_dbz: int = 600_000, n_vr: int = 400_000, gts: bool = True
               ) -> Tuple[Scenario, np.random.Generator]:
    """BASELINE config 2 (k=32) / 3 (k=256): 450x450x50, dx = 2 km, ~10^6 radar + ~10^4 GTS."""
    rng = np.random.default_rng(seed)
    sc = Scenario("M", nx, ny, nz, k, 2000.0, make_grid(nx, ny, nz, 2000.0), seed=seed)
    if gts:
        add_gts(sc, rng)
    add_radar(sc, rng, n_dbz, n_vr)
    return sc, rng


def scenario_tiny(k: int = 8, nx: int = 12, ny: int = 10, nz: int = 6, seed: int = 3,
                  n_dbz: int = 900, n_vr: int = 700, dx: float = 3000.0
                  ) -> Tuple[Scenario, np.random.Generator]:
    """Small mixed case for quick parity checks (all families, truncation exercised)."""
    rng = np.random.default_rng(seed)
    sc = Scenario("tiny", nx, ny, nz, k, dx, make_grid(nx, ny, nz, dx), seed=seed)
    add_gts(sc, rng, n_synop=60, n_metar=30, n_ships=10, n_sound=4, n_lev=12)
    add_radar(sc, rng, n_dbz, n_vr, n_sites=2, radius=15e3)
    return sc, rng
