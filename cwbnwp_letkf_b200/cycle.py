"""One analysis cycle of ``letkf_driver`` (module_letkf_core.f90:21-297) with the ensemble resident in HBM.

The reference keeps every member's full 3-D field on the rank that read it (member-major) and, for each entry of
``var_update``, transposes it to column-major with ``letkf_scatter_grid`` (module_mpi_util.f90:190-290), runs the
grid-point loop (core:209-240), applies ``letkf_tune_q`` (core:252-278) and transposes back with
``letkf_gather_grid`` (mpi:292-358); the vertical coordinate is the ensemble-mean geopotential height that
``letkf_scatter_vcoord`` builds with one ``sgemv`` over the members (mpi:445-580).

Here the same sequence runs on device tensors:

* ``state[key]``: torch CUDA tensor ``[m_r, nz(+1), ny(+1), nx(+1)]`` -- this rank's members of the full grid
  (x fastest), exactly the memory of the reference's ``wrf(member)%var(nx, ny, nz)``;
* scatter / gather are the NCCL exchanges of ``partition`` (a plain slice on one rank);
* variables that ``letkf_driver`` would analyse with identical settings go through ONE library call with
  ``nfields = len(group)`` (``driver.group_variables``): with input.nml the eight hydrometeor variables share
  their local observation lists and weights;
* the ensemble-mean height is formed on the device in the reference's summation order (member 0..k-1, real32);
* ``letkf_b200_set_levels`` lets 2-D localised variables (MU, P, PH) solve once per column.

Coordinates x, y come from ``driver.Projection`` on the host (module_projection.f90:37-50; a few hundred
kilobytes per stagger), as at the C ABI where the Fortran host passes the x, y it computed itself.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable

import numpy as np

from . import partition
from .driver import G, VARIABLES, Projection, group_variables


def lonlat_to_xy_device(proj: Projection, lon, lat):
    """module_projection.f90:37-50 on device tensors (real32, the constants of proj_init evaluated once on the host
    exactly as driver.Projection does).  The transcendental functions are the device library's, so x, y can differ
    from the host evaluation in the last bits (~1 m at rh ~ 1e7 m) -- the Fortran source does not define those bits
    either (compiler libm); DeviceCycle uses it only with device_projection=True and defaults to the host
    projection so that parity with the letkf_driver mirror stays bit-exact."""
    import torch
    f32 = torch.float32
    lon, lat = lon.to(f32), lat.to(f32)
    half, pi, d2r = 0.5, float(proj._pi), float(proj._d2r)
    cot = 1.0 / torch.tan(half * (half * pi + lat * d2r))
    rh = float(np.float32(proj.earthradius)) * float(proj.f) * torch.exp(float(proj.n) * torch.log(cot))
    dlon = float(proj.n) * (lon * d2r - float(proj.lon0))
    return rh * torch.sin(dlon), float(proj.rh0) - rh * torch.cos(dlon)


class DeviceCycle:
    """``run(state, geo, var_update)`` == the ``update`` loop of letkf_driver for this rank, fields in HBM.

    eng      : host.LetkfB200 with the (replicated) observations set
    namelist : variable name -> config.VarConfig
    geo      : dict of host numpy arrays ``xlat/xlon[nx,ny]``, ``xlat_u/xlon_u[nx+1,ny]``, ``xlat_v/xlon_v[nx,ny+1]``,
               ``hgt[nx,ny]`` (the layout driver.LetkfDriver takes)
    """

    def __init__(self, eng, namelist: Callable[[str], object], proj: Projection, rank: int = 0, world: int = 1,
                 nxb: int = 1, nyb: int = 1, batch: bool = True, device_projection: bool = False):
        self.eng, self.namelist, self.proj = eng, namelist, proj
        self.device_projection = device_projection   # proj%lonlat_to_xy (core:211) on the device, see lonlat_to_xy_device
        self.rank, self.world, self.nxb, self.nyb, self.batch = rank, world, nxb, nyb, batch
        self.log = []
        self.ms_exchange = 0.0
        self.ms_analysis = 0.0

    def _mean_height(self, ph_cols, vs: int):
        """ph_cols: [k, nz+1, ly, lx] full geopotential of this rank's columns.  mpi:528-539: sgemv('n') with
        alpha = 1/(g k), x = 1 accumulates member by member in real32."""
        import torch
        k = ph_cols.shape[0]
        alpha = np.float32(1.0) / (G * np.float32(k))
        tmp = torch.zeros_like(ph_cols[0])
        for m in range(k):
            tmp = tmp + ph_cols[m] * float(alpha)   # separate multiply and add: no contraction
        if vs == 1:
            return tmp
        return (tmp[1:] + tmp[:-1]) * 0.5

    def run(self, state: Dict[str, "object"], geo: Dict[str, np.ndarray], var_update: Iterable[str],
            before_group=None, after_group=None):
        """before_group(group) / after_group(group): optional hooks around each group of variables (bench_cycle.py
        moves the group's fields between pinned host memory and the device there)."""
        import torch
        eng, k = self.eng, self.eng.k
        nx, ny = geo["xlat"].shape
        tab = partition.local_index_tables(self.rank, self.world, nx, ny, self.nxb, self.nyb)
        xloc, yloc = tab["xloc"], tab["yloc"]
        loc_nx, loc_ny = len(xloc), len(yloc)
        dev = state["ph"].device
        on_gpu = dev.type == "cuda"              # (CPU tensors + gloo: the world-size-2 test of this loop)
        hstag, vstag = 0, 0                      # core:57-58
        xy = alt = None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if on_gpu else None

        def mark(i):
            if on_gpu:
                ev[i].record()
        for group in group_variables(var_update, self.namelist, batch=self.batch):
            name = group[0]
            key, hs, vs, is_q = VARIABLES[name]
            cfg = self.namelist(name)
            if before_group is not None:
                before_group(group)
            if not any(t.use_it and t.hclr > 0 for t in cfg.types):
                for nm in group:
                    self.log.append((nm, "skipped: no observation type localises this variable"))
                continue                         # core:66
            hreset, hstag = hstag != hs, hs      # check_coordinate
            vreset, vstag = vstag != vs, vs
            mark(0)
            # letkf_scatter_grid of every variable of the group: [k, vnz, ly, lx] each
            cols = []
            for nm in group:
                f = state[VARIABLES[nm][0]]
                if nm == "MU":
                    f = f[:, None]               # core:142-146: one level
                cols.append(partition.scatter_grid(f, k, self.rank, self.world, self.nxb, self.nyb, stagger=hs))
            vnz = cols[0].shape[1]
            if xy is None or hreset:             # core:165-186
                sfx = {0: "", 1: "_u", 2: "_v"}[hs]
                xi = tab["xloc_u"] if hs == 1 else xloc
                yj = tab["yloc_v"] if hs == 2 else yloc
                lat = geo["xlat" + sfx][np.ix_(xi, yj)]
                lon = geo["xlon" + sfx][np.ix_(xi, yj)]
                if self.device_projection and isinstance(self.proj, Projection):
                    dlon = torch.from_numpy(np.ascontiguousarray(lon[:loc_nx, :loc_ny].T)).to(dev)
                    dlat = torch.from_numpy(np.ascontiguousarray(lat[:loc_nx, :loc_ny].T)).to(dev)
                    xy = lonlat_to_xy_device(self.proj, dlon, dlat)                  # [loc_ny, loc_nx]
                else:
                    x, y = self.proj.lonlat_to_xy(lon[:loc_nx, :loc_ny], lat[:loc_nx, :loc_ny])
                    xy = (torch.from_numpy(np.ascontiguousarray(x.T)).to(dev),
                          torch.from_numpy(np.ascontiguousarray(y.T)).to(dev))        # [loc_ny, loc_nx]
            if alt is None or vreset:            # core:189-206
                if vs == -1:
                    h = np.asarray(geo["hgt"], np.float32)[np.ix_(xloc, yloc)]
                    alt = torch.from_numpy(np.ascontiguousarray(h.T)).to(dev)[None]          # [1, ly, lx]
                else:
                    ph_cols = partition.scatter_grid(state["ph"], k, self.rank, self.world, self.nxb, self.nyb, 0)
                    alt = self._mean_height(ph_cols, vs)                                       # [vnz, ly, lx]
                    del ph_cols
            xyz = torch.empty((vnz, loc_ny, loc_nx, 3), dtype=torch.float32, device=dev)
            xyz[..., 0] = xy[0][None]
            xyz[..., 1] = xy[1][None]
            xyz[..., 2] = alt[:vnz]
            # the loop (core:209-240) runs over the mass extents whatever the array extents are
            npts = vnz * loc_ny * loc_nx
            work = torch.empty((len(group), k, npts), dtype=torch.float32, device=dev)
            for gi, c in enumerate(cols):
                work[gi] = c[:, :, :loc_ny, :loc_nx].reshape(k, npts)
            mark(1)
            if on_gpu:
                torch.cuda.current_stream().synchronize()    # the library runs on its own stream
            cfg.tune_q = bool(is_q)                       # letkf_tune_q as the epilogue of the pass (core:252-278)
            eng.set_levels(vnz)
            stats = eng.analyze_dev(cfg, xyz.reshape(-1, 3), work if len(group) > 1 else work[0])
            eng.set_levels(1)
            mark(2)
            # letkf_gather_grid
            for gi, (nm, c) in enumerate(zip(group, cols)):
                c[:, :, :loc_ny, :loc_nx] = work[gi].reshape(k, vnz, loc_ny, loc_nx)
                f = state[VARIABLES[nm][0]]
                partition.gather_grid(c, f[:, None] if nm == "MU" else f, k, self.rank, self.world, self.nxb,
                                      self.nyb, stagger=hs)
                self.log.append((nm, stats))
            mark(3)
            if on_gpu:
                ev[3].synchronize()
                self.ms_exchange += ev[0].elapsed_time(ev[1]) + ev[2].elapsed_time(ev[3])
            self.ms_analysis += float(getattr(stats, "ms_total", 0.0))
            del cols, work, xyz
            if after_group is not None:
                after_group(group)
        return self.log


def to_member_major(a: np.ndarray) -> np.ndarray:
    """driver.LetkfDriver layout [nx, ny, (nz,) k] -> member-major [k, (nz,) ny, nx] (x fastest)."""
    return np.ascontiguousarray(np.transpose(a, tuple(range(a.ndim - 1, -1, -1))))


def from_member_major(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.transpose(a, tuple(range(a.ndim - 1, -1, -1))))
