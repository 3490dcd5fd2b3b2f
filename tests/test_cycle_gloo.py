"""CPU tier, world_size = 2 over gloo: the whole update loop of letkf_driver as cycle.DeviceCycle runs it --
member-major fields, letkf_scatter_grid / letkf_gather_grid as batched send/recv (module_mpi_util.f90:190-358),
ensemble-mean height from the scattered PH in the reference's summation order (mpi:528-539), grouped hydrometeor
pass, tune_q -- with the CPU oracle behind the engine interface, against the single-process letkf_driver mirror.
The local analysis is the same arithmetic in both, so every field must be bit-identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cwbnwp_letkf_b200 import cycle as CY
from cwbnwp_letkf_b200 import driver as D
from cwbnwp_letkf_b200 import partition as P
from oracle import oracle as O

from _driver_case import KEYS_ALL, OracleBackend, VARS_ALL, copy_state, make_state, namelist

GEO = ("xlat", "xlon", "xlat_u", "xlon_u", "xlat_v", "xlon_v", "hgt")


class OracleEngine:
    """The oracle behind the interface DeviceCycle expects of host.LetkfB200 (CPU tensors instead of device ones)."""

    class _Stats:
        ms_total = 0.0

        def __init__(self, npo, rows):
            self.npts_analysed, self.rows, self.units = npo, rows, npo

    def __init__(self, sc):
        self.k = sc.k
        self.orc = O.Oracle(sc.k, True)
        for o in sc.obs.values():
            self.orc.set_obs(o)

    def set_levels(self, nz):
        pass

    def analyze_dev(self, cfg, xyz, var):
        v = var.numpy()                      # shares memory with the tensor: updated in place
        npo, rows = self.orc.analyze(cfg, xyz.numpy(), v, nthreads=2)
        if cfg.tune_q:                       # the device engine applies letkf_tune_q as its epilogue
            for f in (v if v.ndim == 3 else [v]):
                O.tune_q(f)
        return self._Stats(npo, rows)


def test_mean_height_on_tensors_equals_the_numpy_restatement():
    rng = np.random.default_rng(2)
    ph = (9.81 * (500.0 + 1000.0 * rng.random((7, 6, 5, 9)))).astype(np.float32)       # [nx, ny, nz+1, k]
    cyc = CY.DeviceCycle(None, namelist, None)
    for vs in (0, 1):
        ref = D.ensemble_mean_height(ph, vs)                                               # [nx, ny, nz(+1)]
        got = cyc._mean_height(torch.from_numpy(CY.to_member_major(ph)), vs).numpy()       # [nz(+1), ny, nx]
        assert np.array_equal(np.transpose(got, (2, 1, 0)), ref)


def test_projection_on_tensors_agrees_with_the_host_projection_to_real32_rounding():
    from _driver_case import PROJ
    proj = D.Projection(**PROJ)
    rng = np.random.default_rng(4)
    lon = (120.5 + rng.uniform(-4, 4, (40, 30))).astype(np.float32)
    lat = (23.5 + rng.uniform(-4, 4, (40, 30))).astype(np.float32)
    x, y = proj.lonlat_to_xy(lon, lat)
    xt, yt = CY.lonlat_to_xy_device(proj, torch.from_numpy(lon), torch.from_numpy(lat))
    # rh ~ 1e7 m in real32: one ulp is ~1 m; the two libm implementations may differ by a few
    assert np.abs(xt.numpy() - x).max() < 8.0 and np.abs(yt.numpy() - y).max() < 8.0
    assert xt.dtype == torch.float32 and np.abs(x).max() > 1e5


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, wrf, proj = make_state(k=8)
    lo, hi = P.member_slice(rank, world, sc.k)
    state = {key: torch.from_numpy(CY.to_member_major(wrf[key])[lo:hi].copy()) for key in KEYS_ALL}
    geo = {g: wrf[g] for g in GEO}
    log = CY.DeviceCycle(OracleEngine(sc), namelist, proj, rank, world, nxb=2).run(state, geo, VARS_ALL)
    ok = [n for n, _ in log] == VARS_ALL
    full = {key: CY.from_member_major(P.allgather_members(state[key].contiguous(), sc.k, rank, world).numpy())
            for key in KEYS_ALL}
    if rank == 0:
        ref = copy_state(wrf)
        D.LetkfDriver(OracleBackend(sc), namelist, proj).run(ref, VARS_ALL)
        for key in KEYS_ALL:
            a, b = full[key], ref[key]
            same = np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
            ok = ok and same
        ok = ok and any((full[k_] != wrf[k_])[~np.isnan(full[k_])].any() for k_ in ("t", "qs", "ph"))
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_device_cycle_equals_single_process_driver():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == 1 and out[1] == 1
