"""Synthetic WRF-shaped state for the letkf_driver mirror tests (shared by the CPU and GPU tiers)."""
import numpy as np

from cwbnwp_letkf_b200 import config as C
from cwbnwp_letkf_b200 import driver as D
from cwbnwp_letkf_b200 import synthetic as S
from oracle import oracle as O

PROJ = dict(cen_lat=23.5, truelat1=10.0, truelat2=40.0, sta_lon=120.5)
VARS = ["U", "V", "W", "T", "QVAPOR", "QRAIN", "P", "MU", "PH"]
VARS_ALL = list(C.VAR_UPDATE)            # input.nml:7: all 16, the 8 hydrometeor variables share one configuration
KEYS_ALL = ("u", "v", "w", "t", "qv", "qr", "qs", "qg", "qh", "nqr", "nqs", "nqg", "nqh", "p", "mu", "ph")


def inverse_projection(p: D.Projection, x, y):
    """float64 inverse of module_projection.f90:37-50 (test-side only: builds lat/lon whose forward
    projection lands on the synthetic grid)."""
    n, f, rh0, R = float(p.n), float(p.f), float(p.rh0), float(p.earthradius)
    rh = np.hypot(x, rh0 - y)
    dlon = np.arctan2(x, rh0 - y)
    lon = (float(p.lon0) + dlon / n) * 180.0 / np.pi
    cot = (rh / (R * f)) ** (1.0 / n)
    lat = (2.0 * np.arctan(1.0 / cot) - np.pi / 2) * 180.0 / np.pi
    return lon.astype(np.float32), lat.astype(np.float32)


def make_state(k=8, nx=9, ny=8, nz=5, dx=3000.0, seed=11):
    rng = np.random.default_rng(seed)
    sc = S.Scenario("drv", nx, ny, nz, k, dx, S.make_grid(nx, ny, nz, dx), seed=seed)
    S.add_gts(sc, rng, n_synop=50, n_metar=25, n_ships=8, n_sound=4, n_lev=10)
    S.add_radar(sc, rng, 700, 500, n_sites=2, radius=12e3)
    proj = D.Projection(**PROJ)

    def coords(nxx, nyy, offx, offy):
        xs = (np.arange(nxx) - (nx - 1) / 2 + offx) * dx
        ys = (np.arange(nyy) - (ny - 1) / 2 + offy) * dx
        X, Y = np.meshgrid(xs, ys, indexing="ij")
        return inverse_projection(proj, X, Y)

    wrf = {}
    wrf["xlon"], wrf["xlat"] = coords(nx, ny, 0.0, 0.0)
    wrf["xlon_u"], wrf["xlat_u"] = coords(nx + 1, ny, -0.5, 0.0)
    wrf["xlon_v"], wrf["xlat_v"] = coords(nx, ny + 1, 0.0, -0.5)
    ter = (300.0 + 200.0 * rng.random((nx, ny))).astype(np.float32)
    wrf["hgt"] = ter
    lev = (15000.0 * (np.arange(nz + 1) / nz) ** 1.4).astype(np.float32)
    ph = (ter[:, :, None] + lev[None, None, :]) * np.float32(9.81)
    wrf["ph"] = (ph[..., None] + rng.normal(0, 30.0, (nx, ny, nz + 1, k))).astype(np.float32)

    def fld(shape, mean, amp):
        return (mean + amp * rng.standard_normal(shape + (k,))).astype(np.float32)

    wrf["u"] = fld((nx + 1, ny, nz), 5.0, 2.0)
    wrf["v"] = fld((nx, ny + 1, nz), -3.0, 2.0)
    wrf["w"] = fld((nx, ny, nz + 1), 0.0, 0.5)
    wrf["t"] = fld((nx, ny, nz), 290.0, 1.5)
    wrf["p"] = fld((nx, ny, nz), 8.0e4, 300.0)
    wrf["mu"] = fld((nx, ny), 9.0e4, 200.0)
    q = np.abs(fld((nx, ny, nz), 2e-3, 2e-3))
    q[rng.random((nx, ny, nz)) < 0.2] = 0.0          # clear-air points (tune_q 0/0)
    wrf["qv"] = q
    qr = np.abs(fld((nx, ny, nz), 1e-4, 3e-4))
    qr[rng.random(qr.shape) < 0.3] = 0.0
    wrf["qr"] = qr
    for i, key in enumerate(("qs", "qg", "qh", "nqr", "nqs", "nqg", "nqh")):   # the other hydrometeor variables
        q2 = np.abs(fld((nx, ny, nz), 1e-4 * (i + 1), 3e-4))
        q2[rng.random(q2.shape) < 0.3] = 0.0
        wrf[key] = q2
    return sc, wrf, proj


def namelist(name):
    # radar and GTS on for every variable of this case; PH/W/MU/... take whatever input.nml gives them
    return C.sample_namelist(name)


class OracleBackend:
    """The CPU oracle behind the backend interface of driver.LetkfDriver (test infrastructure)."""

    def __init__(self, sc, real64=True):
        self.orc = O.Oracle(sc.k, real64)
        for o in sc.obs.values():
            self.orc.set_obs(o)

    def analyze(self, cfg, xyz, var):
        return self.orc.analyze(cfg, xyz, var, nthreads=4)

    def tune_q(self, var):
        O.tune_q(var)


def copy_state(wrf):
    return {k: v.copy() for k, v in wrf.items()}
