"""CPU tier: the letkf_driver mirror (cwbnwp_letkf_b200/driver.py) driven through the CPU oracle.

Pins the dispatch logic of module_letkf_core.f90:59-297 that sits either side of the hot path: stagger
rules, coordinate caching (Q14), tune_q dispatch, the decomposition tables of
letkf_local_info.  The GPU tier (test_gpu_parity.py::test_driver_*) runs the same dispatch through the
C ABI and compares with what this file produces."""
import numpy as np
import pytest

from cwbnwp_letkf_b200 import config as C
from cwbnwp_letkf_b200 import driver as D
from cwbnwp_letkf_b200 import partition as P

from _driver_case import OracleBackend, VARS, copy_state, make_state, namelist


def test_projection_round_trip():
    from _driver_case import PROJ, inverse_projection
    proj = D.Projection(**PROJ)
    x = np.array([[-30e3, 0.0, 45e3]]).repeat(2, 0)
    y = np.array([[-20e3], [25e3]]).repeat(3, 1)
    lon, lat = inverse_projection(proj, x, y)
    xx, yy = proj.lonlat_to_xy(lon, lat)
    # real32 transcendentals at earth-radius scale: ~1e-7 * 1.3e7 m
    assert np.abs(xx - x).max() < 8.0 and np.abs(yy - y).max() < 8.0
    # true-scale latitude: one degree of longitude at 23.5N is ~102 km
    x1, _ = proj.lonlat_to_xy(np.float32(121.5), np.float32(23.5))
    assert 95e3 < float(x1) < 108e3


def test_projection_known_answer_with_the_reference_constants():
    """lonlat_to_xy against a float64 evaluation of module_projection.f90:27-50 with the reference's own
    earthradius = 6.37122e6 (module_param.f90:108) -- an independent restatement, not the Projection class."""
    cen_lat, t1, t2, sta_lon = 23.5, 10.0, 40.0, 120.5          # input.nml:15-19 style
    R = 6.37122e6
    d2r = np.pi / 180.0
    lat0, lat1, lat2, lon0 = cen_lat * d2r, t1 * d2r, t2 * d2r, sta_lon * d2r
    cot = lambda x: 1.0 / np.tan(x)
    n = np.log(np.cos(lat1) / np.cos(lat2)) / np.log(np.tan(0.5 * (0.5 * np.pi + lat2)) * cot(0.5 * (0.5 * np.pi + lat1)))
    f = np.cos(lat1) * np.exp(n * np.log(np.tan(0.5 * (0.5 * np.pi + lat1)))) / n
    rh0 = R * f * np.exp(n * np.log(cot(0.5 * (0.5 * np.pi + lat0))))
    lon = np.array([118.0, 120.5, 123.25, 121.0])
    lat = np.array([21.5, 23.5, 25.75, 27.0])
    rh = R * f * np.exp(n * np.log(cot(0.5 * (0.5 * np.pi + lat * d2r))))
    dl = n * (lon * d2r - lon0)
    x_ref, y_ref = rh * np.sin(dl), rh0 - rh * np.cos(dl)
    proj = D.Projection(cen_lat, t1, t2, sta_lon)
    assert proj.earthradius == 6.37122e6
    x, y = proj.lonlat_to_xy(lon, lat)
    # real32 evaluation of rh ~ 1e7 m: a few metres of rounding; the old 6.37e6 radius was ~2 km off
    assert np.abs(x - x_ref).max() < 30.0 and np.abs(y - y_ref).max() < 30.0
    assert abs(x[1]) < 1.0 and abs(y[1]) < 30.0                 # the projection centre maps to the origin


def test_index_tables_cover_the_grid_once():
    nx, ny = 11, 7
    for world, nxb, nyb in [(1, 1, 1), (2, 1, 1), (4, 1, 1), (6, 2, 3), (8, 1, 2)]:
        seen = np.zeros((nx, ny), int)
        seen_u = np.zeros(nx + 1, int)
        seen_v = np.zeros(ny + 1, int)
        npx, npy = P.process_grid(world)
        for r in range(world):
            t = P.local_index_tables(r, world, nx, ny, nxb, nyb)
            seen[np.ix_(t["xloc"], t["yloc"])] += 1
            if r // npx == 0:
                seen_u[t["xloc_u"]] += 1
            if r % npx == 0:
                seen_v[t["yloc_v"]] += 1
            # the staggered table is the mass table plus at most one trailing entry (core:71-78 relies on it)
            assert np.array_equal(t["xloc_u"][:len(t["xloc"])], t["xloc"]) and len(t["xloc_u"]) - len(t["xloc"]) in (0, 1)
            assert np.array_equal(t["yloc_v"][:len(t["yloc"])], t["yloc"]) and len(t["yloc_v"]) - len(t["yloc"]) in (0, 1)
        assert (seen == 1).all() and (seen_u == 1).all() and (seen_v == 1).all()
    # block size 1 == the cyclic partition the multi-GPU bench uses
    t = P.local_index_tables(3, 4, nx, ny)
    cols = (t["xloc"][None, :] + nx * t["yloc"][:, None]).reshape(-1)
    assert np.array_equal(cols, P.local_columns(3, 4, nx, ny))


def test_ensemble_mean_height():
    rng = np.random.default_rng(0)
    ph = (9.81 * (1000.0 + 500.0 * np.arange(4)[None, None, :, None]) + rng.normal(0, 20, (3, 2, 4, 6))).astype(np.float32)
    full = D.ensemble_mean_height(ph, 1)
    mass = D.ensemble_mean_height(ph, 0)
    ref = ph.astype(np.float64).mean(-1) / 9.81
    assert full.shape == (3, 2, 4) and mass.shape == (3, 2, 3)
    assert np.abs(full - ref).max() < 2e-3
    assert np.abs(mass - 0.5 * (ref[:, :, 1:] + ref[:, :, :-1])).max() < 2e-3


@pytest.fixture(scope="module")
def case():
    sc, wrf, proj = make_state()
    out = copy_state(wrf)
    drv = D.LetkfDriver(OracleBackend(sc), namelist, proj)
    log = drv.run(out, VARS)
    return sc, wrf, proj, out, log


def test_driver_dispatch_and_stagger_rules(case):
    sc, wrf, proj, out, log = case
    nx, ny, nz = sc.nx, sc.ny, sc.nz
    names = [n for n, _ in log]
    assert names == VARS
    # every variable of this case has observations in range somewhere
    for key in ("u", "v", "w", "t", "qv", "qr", "p", "mu", "ph"):
        assert not np.array_equal(out[key], wrf[key]), key
    # the last staggered column of U and row of V are never analysed
    assert np.array_equal(out["u"][nx], wrf["u"][nx])
    assert np.array_equal(out["v"][:, ny], wrf["v"][:, ny])
    # fields the driver has no business touching
    for key in ("xlat", "xlon", "hgt"):
        assert np.array_equal(out[key], wrf[key])


def test_driver_matches_direct_calls(case):
    """T and QRAIN through the driver == the hot path called directly with independently built
    coordinates (mass-level ensemble-mean height, projected lat/lon), + tune_q for the q variable."""
    from oracle import oracle as O
    sc, wrf, proj, out, log = case
    nx, ny, nz, k = sc.nx, sc.ny, sc.nz, sc.k
    x, y = proj.lonlat_to_xy(wrf["xlon"], wrf["xlat"])
    alt = D.ensemble_mean_height(wrf["ph"], 0)
    xyz = np.stack([np.broadcast_to(x.T[None], (nz, ny, nx)), np.broadcast_to(y.T[None], (nz, ny, nx)),
                    np.transpose(alt, (2, 1, 0))], -1).astype(np.float32).reshape(-1, 3)
    be = OracleBackend(sc)
    for name, key in (("T", "t"), ("QRAIN", "qr")):
        # PH is updated last in VARS, so the heights seen by T / QRAIN are those of the input state
        work = np.ascontiguousarray(np.transpose(wrf[key], (3, 2, 1, 0))).reshape(k, -1)
        be.analyze(namelist(name), xyz, work)
        if name == "QRAIN":
            O.tune_q(work)
        ref = np.transpose(work.reshape(k, nz, ny, nx), (3, 2, 1, 0))
        a, b = out[key], ref
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), name


def test_driver_two_ranks_equal_one(case):
    """Column ownership does not change a single value (every grid point is an independent unit)."""
    sc, wrf, proj, out, log = case
    merged = copy_state(wrf)
    for r in range(2):
        loc = copy_state(wrf)
        D.LetkfDriver(OracleBackend(sc), namelist, proj, rank=r, world=2).run(loc, ["U", "T", "MU"])
        t = P.local_index_tables(r, 2, sc.nx, sc.ny)
        for key, xi, yj in (("u", t["xloc_u"], t["yloc"]), ("t", t["xloc"], t["yloc"]), ("mu", t["xloc"], t["yloc"])):
            merged[key][np.ix_(xi, yj)] = loc[key][np.ix_(xi, yj)]
    for key in ("u", "t", "mu"):
        a, b = merged[key], out[key]
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), key


def test_driver_skips_variable_without_trees(case):
    sc, wrf, proj, out, log = case

    def nml_off(name):
        cfg = C.sample_namelist(name)
        for t in cfg.types:
            t.hclr = -1.0
        return cfg

    st = copy_state(wrf)
    lg = D.LetkfDriver(OracleBackend(sc), nml_off, proj).run(st, ["T", "", "U"])
    assert len(lg) == 1 and lg[0][0] == "T" and isinstance(lg[0][1], str)   # '' ends the list (core:61)
    assert np.array_equal(st["t"], wrf["t"])
    with pytest.raises(ValueError):
        D.LetkfDriver(OracleBackend(sc), namelist, proj).run(copy_state(wrf), ["QCLOUD"])


def test_batched_hydrometeor_pass_equals_single_variable_passes(case):
    """The eight hydrometeor variables of input.nml share their configuration: one pass with nfields = 8
    (driver.group_variables) gives bit-identical fields to eight passes (core:59-297 analyses them one by one)."""
    from _driver_case import KEYS_ALL, VARS_ALL
    sc, wrf, proj = case[:3]
    backend = OracleBackend(sc)
    groups = D.group_variables(VARS_ALL, namelist)
    assert [len(g) for g in groups] == [1, 1, 1, 1, 1, 8, 1, 1, 1]
    a, b = copy_state(wrf), copy_state(wrf)
    la = D.LetkfDriver(backend, namelist, proj, batch=True).run(a, VARS_ALL)
    lb = D.LetkfDriver(backend, namelist, proj, batch=False).run(b, VARS_ALL)
    assert [n for n, _ in la] == [n for n, _ in lb] == VARS_ALL
    for key in KEYS_ALL:
        assert np.array_equal(np.isnan(a[key]), np.isnan(b[key])), key
        ok = ~np.isnan(a[key])
        assert np.array_equal(a[key][ok], b[key][ok]), key
    assert any((a[k_] != wrf[k_])[~np.isnan(a[k_])].any() for k_ in ("qs", "nqh"))
