"""Host-side mirror of ``letkf_driver`` (module_letkf_core.f90:21-298): the per-variable dispatch that
sits either side of the hot path.  One instance = one rank = one GPU.

What it restates, with the reference line it follows:

* the variable table -- which model field a name selects, its horizontal stagger (0 mass, 1 U, 2 V),
  its vertical stagger (0 mass levels, 1 W/PH levels, -1 surface) and whether ``letkf_tune_q`` runs
  afterwards (core:96-162, 243-291);
* array extents: U is one column wider (``loc_nx_u``), V one row taller (``loc_ny_v``), W/PH have
  nz+1 levels, MU has one (core:71-82) -- but the loop always runs over the MASS extents
  ``cpu%loc_nx x cpu%loc_ny`` (core:209-210; SURVEY Q6: the last staggered column / row is never analysed);
* coordinate caching: lat/lon are re-sliced only when the horizontal stagger changes, the height of the
  grid points only when the vertical stagger changes (``check_coordinate``, core:735-747; core:165-206).
  ``alt`` is allocated with the mass extents and the nz of the variable that triggered the refresh;
* the height itself: ensemble mean of the full geopotential divided by g (``sgemv`` with
  alpha = 1/(g*nmember), module_mpi_util.f90:528-539), averaged to mass levels when unstaggered; terrain
  height for MU;
* ``proj%lonlat_to_xy`` (module_projection.f90:21-50) for the query point (core:211);
* ``letkf_scatter_grid`` / ``letkf_gather_grid`` reduced to what they mean for one rank: take / put the
  owned columns (index tables of ``letkf_local_info``, see partition.local_index_tables).

The analysis itself is ``backend.analyze`` -- the C-ABI call (host.LetkfB200).  The backend is duck
typed so that the parity tests can drive the CPU oracle through the same dispatch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional

import numpy as np

from . import config as C
from . import partition

G = np.float32(9.81)  # module_param.f90:109

# name -> (field key, Hstag, Vstag, tune_q)                                  core:96-162, 243-291
VARIABLES = {
    "U": ("u", 1, 0, False), "V": ("v", 2, 0, False), "W": ("w", 0, 1, False), "T": ("t", 0, 0, False),
    "QVAPOR": ("qv", 0, 0, True), "QRAIN": ("qr", 0, 0, True), "QSNOW": ("qs", 0, 0, True),
    "QGRAUP": ("qg", 0, 0, True), "QHAIL": ("qh", 0, 0, True), "QNRAIN": ("nqr", 0, 0, True),
    "QNSNOW": ("nqs", 0, 0, True), "QNGRAUPEL": ("nqg", 0, 0, True), "QNHAIL": ("nqh", 0, 0, True),
    "P": ("p", 0, 0, False), "MU": ("mu", 0, -1, False), "PH": ("ph", 0, 1, False),
}


@dataclass
class Projection:
    """Lambert conformal projection of module_projection.f90 in real32.  (The Fortran evaluates the
    transcendentals with the compiler's real32 libm, so the last bit is not defined by the source; at the
    C ABI the Fortran host passes x, y it computed itself.)"""
    cen_lat: float
    truelat1: float
    truelat2: float
    sta_lon: float
    earthradius: float = 6.37122e6   # module_param.f90:108 (used by module_projection.f90:34,45)

    def __post_init__(self):
        f32 = np.float32
        d2r = f32(np.pi) / f32(180.0)
        pi = f32(np.pi)
        lat0, lat1, lat2 = f32(self.cen_lat) * d2r, f32(self.truelat1) * d2r, f32(self.truelat2) * d2r
        self._d2r, self._pi = d2r, pi
        self.lon0 = f32(self.sta_lon) * d2r
        half = f32(0.5)
        cot = lambda x: f32(1.0) / np.tan(x, dtype=f32)
        self.n = (np.log(np.cos(lat1) / np.cos(lat2), dtype=f32) /
                  np.log(np.tan(half * (half * pi + lat2)) * cot(half * (half * pi + lat1)), dtype=f32))
        self.f = np.cos(lat1) * np.exp(self.n * np.log(np.tan(half * (half * pi + lat1)))) / self.n
        self.rh0 = f32(self.earthradius) * self.f * np.exp(self.n * np.log(cot(half * (half * pi + lat0))))

    def lonlat_to_xy(self, lon: np.ndarray, lat: np.ndarray):
        f32 = np.float32
        lon, lat = np.asarray(lon, f32), np.asarray(lat, f32)
        half = f32(0.5)
        cot = f32(1.0) / np.tan(half * (half * self._pi + lat * self._d2r), dtype=f32)
        rh = f32(self.earthradius) * f32(self.f) * np.exp(f32(self.n) * np.log(cot, dtype=f32), dtype=f32)
        dlon = f32(self.n) * (lon * self._d2r - f32(self.lon0))
        return (rh * np.sin(dlon, dtype=f32)).astype(f32), (f32(self.rh0) - rh * np.cos(dlon, dtype=f32)).astype(f32)


def ensemble_mean_height(ph: np.ndarray, vstag: int) -> np.ndarray:
    """module_mpi_util.f90:528-539.  ph: [nx, ny, nz+1, k] full geopotential (m2 s-2), real32.  The
    reference calls sgemv('n', ..., alpha = 1/(g*nmember), x = 1): column by column, y += (alpha*x_j) * A(:,j)
    in real32.  Returns [nx, ny, nz+1] (vstag 1) or the mass-level average [nx, ny, nz] (vstag 0)."""
    k = ph.shape[-1]
    alpha = np.float32(1.0) / (G * np.float32(k))
    tmp = np.zeros(ph.shape[:-1], np.float32)
    for m in range(k):
        tmp = tmp + alpha * ph[..., m]
    if vstag == 1:
        return tmp
    return ((tmp[:, :, 1:] + tmp[:, :, :-1]) * np.float32(0.5)).astype(np.float32)


def group_variables(var_update: Iterable[str], namelist: Callable[[str], object], batch: bool = True):
    """var_update (core:59-61: the list ends at the first blank entry) split into runs of CONSECUTIVE variables
    that letkf_driver would analyse with identical settings -- same stagger, same tune_q rule and a bit-identical
    namelist slice (types, hclr/vclr, errors, inflation, RTPP/RTPS).  Such variables have the same local
    observations and the same weights at every grid point, so one pass with nfields = len(group) replaces
    len(group) passes.  With input.nml this groups the eight hydrometeor variables QRAIN .. QNHAIL."""
    groups, prev = [], None
    for name in var_update:
        name = name.strip()
        if not name:
            break
        if name not in VARIABLES:
            raise ValueError("Need to code for unknown variable %s" % name)  # core:159-161
        cfg = namelist(name)
        saved, cfg.tune_q = cfg.tune_q, False
        sig = (VARIABLES[name][1:], bytes(C.to_c(cfg)))
        cfg.tune_q = saved
        if batch and prev == sig and name != "MU":
            groups[-1].append(name)
        else:
            groups.append([name])
        prev = sig
    return groups


class LetkfDriver:
    """``run(wrf, var_update)`` == the ``update`` loop of letkf_driver for this rank."""

    def __init__(self, backend, namelist: Callable[[str], object], proj: Projection, rank: int = 0, world: int = 1,
                 nxb: int = 1, nyb: int = 1, batch: bool = True):
        self.backend = backend          # .analyze(cfg, xyz[npts,3], var[k,npts]) and .tune_q(var[k,npts])
        self.namelist = namelist        # variable name -> VarConfig (module_config.f90:7-75)
        self.proj = proj
        self.rank, self.world, self.nxb, self.nyb = rank, world, nxb, nyb
        self.batch = batch              # analyse variables with identical settings in one call (nfields > 1)
        self.log = []

    @staticmethod
    def _uses_any_tree(cfg) -> bool:
        # build_tree succeeds if some used type has hclr(ivar) > 0 (module_localization.f90:74,113)
        return any(t.use_it and t.hclr > 0 for t in cfg.types)

    def run(self, wrf: Dict[str, np.ndarray], var_update: Iterable[str]):
        nx, ny = wrf["xlat"].shape
        nz = wrf["t"].shape[2]
        tab = partition.local_index_tables(self.rank, self.world, nx, ny, self.nxb, self.nyb)
        xloc, yloc = tab["xloc"], tab["yloc"]
        loc_nx, loc_ny = len(xloc), len(yloc)
        hstag, vstag = 0, 0                      # core:57-58
        lat = lon = alt = None
        for group in group_variables(var_update, self.namelist, batch=self.batch):
            name = group[0]
            key, hs, vs, is_q = VARIABLES[name]
            cfg = self.namelist(name)
            if not self._uses_any_tree(cfg):
                for nm in group:
                    self.log.append((nm, "skipped: no observation type localises this variable"))
                continue                         # core:66
            vnz = nz + 1 if vs == 1 else (1 if vs == -1 else nz)          # core:79-82
            xi = tab["xloc_u"] if hs == 1 else xloc                       # core:71-78
            yj = tab["yloc_v"] if hs == 2 else yloc
            hreset, hstag = hstag != hs, hs      # check_coordinate
            vreset, vstag = vstag != vs, vs
            if lat is None or hreset:            # core:165-186
                sfx = {0: "", 1: "_u", 2: "_v"}[hs]
                lat = wrf["xlat" + sfx][np.ix_(xi, yj)]
                lon = wrf["xlon" + sfx][np.ix_(xi, yj)]
            if alt is None or vreset:            # core:189-206: mass extents, this variable's nz
                if vs == -1:
                    alt = np.asarray(wrf["hgt"], np.float32)[np.ix_(xloc, yloc)][:, :, None]
                else:
                    alt = ensemble_mean_height(wrf["ph"][np.ix_(xloc, yloc)], vs)
            # the loop (core:209-240) runs over the mass extents whatever the array extents are
            x, y = self.proj.lonlat_to_xy(lon[:loc_nx, :loc_ny], lat[:loc_nx, :loc_ny])
            xyz = np.empty((vnz, loc_ny, loc_nx, 3), np.float32)           # point index i + lx*(j + ly*l)
            xyz[..., 0] = x.T[None]
            xyz[..., 1] = y.T[None]
            xyz[..., 2] = np.transpose(alt[:, :, :vnz], (2, 1, 0))
            # letkf_scatter_grid for every variable of the group: [lx, ly, vnz, k] each
            vars_, works = [], []
            for nm in group:
                field = wrf[VARIABLES[nm][0]]
                if nm == "MU":
                    field = field[:, :, None, :]     # core:142-146
                var = np.ascontiguousarray(field[np.ix_(xi, yj)])
                assert var.shape[2] == vnz
                vars_.append(var)
                works.append(np.ascontiguousarray(np.transpose(var[:loc_nx, :loc_ny], (3, 2, 1, 0))).reshape(var.shape[3], -1))
            # variables with identical localisation / inflation share one set of weights per grid point:
            # one call with nfields = len(group) (the eight hydrometeor variables of input.nml:7,37-38,162)
            work = works[0] if len(group) == 1 else np.ascontiguousarray(np.stack(works, 0))
            if hasattr(self.backend, "set_levels"):
                # declare the level count: when every active type is 2-D localised the library solves once
                # per column and all levels share the weights (letkf_b200_set_levels)
                self.backend.set_levels(vnz)
            # letkf_tune_q (core:252-278) runs as an epilogue of the device pass when the backend offers that
            fused_q = bool(getattr(self.backend, "fuses_tune_q", False))
            cfg.tune_q = bool(is_q and fused_q)
            stats = self.backend.analyze(cfg, xyz.reshape(-1, 3), work)
            if hasattr(self.backend, "set_levels"):
                self.backend.set_levels(1)
            for gi, nm in enumerate(group):
                var = vars_[gi]
                w = work if len(group) == 1 else work[gi]
                if is_q and not fused_q:
                    assert var.shape[:2] == (loc_nx, loc_ny)   # q variables are unstaggered
                    w = np.ascontiguousarray(w)
                    self.backend.tune_q(w)       # core:252-278: the whole local array
                var[:loc_nx, :loc_ny] = np.transpose(w.reshape(var.shape[3], vnz, loc_ny, loc_nx), (3, 2, 1, 0))
                k_ = VARIABLES[nm][0]
                out = wrf[k_][:, :, None, :] if nm == "MU" else wrf[k_]
                out[np.ix_(xi, yj)] = var        # letkf_gather_grid (this rank's columns)
                self.log.append((nm, stats))
        return self.log
