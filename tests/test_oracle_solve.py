"""Pins for the letkf_yoyb / letkf_solve / tune_q / eigen restatement
(module_letkf_core.f90:300-733, module_eigen.f90) -- SURVEY.md section 8(c) pins (3)-(5)."""
import numpy as np
import pytest

from cwbnwp_letkf_b200 import config as C
from cwbnwp_letkf_b200 import synthetic as S
from oracle import oracle as O


def _f32(x):
    return np.float32(x)


def _seqsum32(a):
    """sum() as the oracle defines it: sequential, left to right, real32."""
    s = np.float32(0)
    for v in np.asarray(a, np.float32):
        s = np.float32(s + v)
    return s


# ------------------------------------------------------------------ scalar functions
def test_constants():
    assert O.gc1999() == _f32(2) * np.sqrt(_f32(10) / _f32(3), dtype=np.float32)
    assert O.search_r2() == _f32(13.333334)


def test_expf_is_correctly_rounded_on_the_hot_range():
    L = O.lib()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 3.4, 20000), rng.uniform(-20, 20, 2000)]).astype(np.float32)
    mine = np.array([L.or_expf(float(v)) for v in x], np.float32)
    ref = np.exp(x.astype(np.float64)).astype(np.float32)
    assert np.array_equal(mine.view(np.int32), ref.view(np.int32))


def test_gaspari_cohn_against_float64_formula():
    """GC (1999) eq. 4.10 in float64, away from the cutoff where real32 cancellation bites."""
    L = O.lib()
    c = np.sqrt(10.0 / 3.0)

    def gc64(r):
        z = r / c
        if z <= 1:
            return -0.25 * z**5 + 0.5 * z**4 + 0.625 * z**3 - 5 / 3 * z**2 + 1
        if z <= 2:
            return z**5 / 12 - 0.5 * z**4 + 0.625 * z**3 + 5 / 3 * z**2 - 5 * z + 4 - 2 / (3 * z)
        return 0.0

    for r in np.linspace(0, 3.3, 200):
        assert abs(L.or_gaspari_cohn(float(r)) - gc64(float(_f32(r)))) < 3e-6
    assert L.or_gaspari_cohn(4.0) == 0.0
    assert L.or_gaspari_cohn(0.0) == 1.0
    # the real32 Horner form dips below zero just inside the cutoff -> sqrt gives NaN
    zs = np.linspace(1.958, 1.9999, 400) * c
    vals = np.array([L.or_gaspari_cohn(float(z)) for z in zs])
    assert (vals < 0).any() and vals.min() > -1e-5
    neg = float(zs[np.argmin(vals)])
    assert np.isnan(L.or_error_inv(1.0, neg * neg, 1))


def test_error_inv_gaussian():
    L = O.lib()
    for err, r2 in [(1.0, 0.0), (2.5, 3.7), (0.37, 13.3)]:
        e = _f32(np.exp(np.float64(_f32(0.25) * _f32(r2))))
        want = _f32(1.0) / (_f32(err) * e)
        assert L.or_error_inv(err, r2, 0) == want


# ------------------------------------------------------------------ letkf_solve
def _random_problem(k, p, seed):
    rng = np.random.default_rng(seed)
    yb = rng.normal(size=(p, k)).astype(np.float32)
    yb -= yb.mean(1, keepdims=True)              # every row has zero member sum, like core:432
    yb = yb.astype(np.float32)
    yo = rng.normal(size=p).astype(np.float32)
    xb = (280 + rng.normal(size=k)).astype(np.float32)
    return xb, yo, yb


@pytest.mark.parametrize("k", [2, 3, 4])
def test_scalar_kalman_update_closed_form(k):
    """p = 1: analysis mean = xb_mean + cov(x,y)/(var_y*rho + r) * d, in LETKF variables."""
    rng = np.random.default_rng(k)
    orc = O.Oracle(k, True)
    xb = rng.normal(10, 2, k).astype(np.float32)
    hx = rng.normal(3, 1, k).astype(np.float32)
    err, ob, rho = 0.7, 3.9, 1.3
    yp = hx - hx.mean(dtype=np.float32)
    yb = (yp / _f32(err))[None].astype(np.float32)
    yo = np.array([(ob - hx.mean(dtype=np.float32)) / err], np.float32)
    inflat = _f32(k - 1) / _f32(rho)
    xa, wbar, Wa, raw = orc.letkf_solve(xb, yo, yb, inflat)
    xp = xb.astype(np.float64) - np.float64(xb.sum(dtype=np.float32) * _f32(1.0 / k))
    y = yb[0].astype(np.float64)
    # Pa~ = (mu I + y y^T)^-1 ; Sherman-Morrison
    mu = float(inflat)
    wbar_cf = y * float(yo[0]) / (mu + y @ y)
    assert np.allclose(wbar, wbar_cf, rtol=1e-12, atol=1e-15)
    # Pa~^(1/2) in closed form for a rank-one update of mu*I
    yy = y @ y
    Pah = (np.eye(k) - np.outer(y, y) / yy * (1 - np.sqrt(mu / (mu + yy)))) / np.sqrt(mu)
    Wa_cf = np.sqrt(k - 1) * Pah
    assert np.allclose(Wa, Wa_cf, rtol=1e-11, atol=1e-13)
    xm = np.float64(xb.sum(dtype=np.float32) * _f32(1.0 / k))
    want = xm + xp @ wbar_cf + Wa_cf.T @ xp
    assert np.allclose(raw, want, rtol=1e-13, atol=0)
    # perturbation covariance: Wa Wa^T = (k-1) Pa~
    Pa = np.eye(k) / mu - np.outer(y, y) / (mu * (mu + y @ y))
    assert np.allclose(Wa @ Wa.T, (k - 1) * Pa, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("k,p", [(8, 3), (8, 40), (32, 300), (32, 10), (64, 120)])
def test_letkf_identities(k, p):
    orc = O.Oracle(k, True)
    xb, yo, yb = _random_problem(k, p, 10 * k + p)
    rho = 1.6
    inflat = _f32(k - 1) / _f32(rho)
    xa, wbar, Wa, raw = orc.letkf_solve(xb, yo, yb, inflat)
    Y = yb.astype(np.float64).T                       # k x p
    Cm = float(inflat) * np.eye(k) + Y @ Y.T
    Pa = np.linalg.inv(Cm)
    assert np.allclose(wbar, Pa @ (Y @ yo.astype(np.float64)), rtol=1e-10, atol=1e-13)
    assert np.allclose(Wa, Wa.T, atol=1e-12)
    assert np.allclose(Wa @ Wa, (k - 1) * Pa, rtol=1e-10, atol=1e-12)
    # Yb^T 1 = 0  =>  C 1 = mu 1  =>  Wa 1 = sqrt((k-1)/mu) 1 = sqrt(rho) 1, wbar . 1 = 0
    assert np.allclose(Wa @ np.ones(k), np.sqrt((k - 1) / float(inflat)), rtol=1e-6)
    assert abs(wbar.sum()) < 1e-6
    xm = np.float64(_seqsum32(xb) * (_f32(1.0) / _f32(k)))
    xp = xb.astype(np.float64) - xm
    want = xm + xp @ wbar + Wa.T @ xp
    assert np.allclose(raw, want, rtol=1e-12)
    assert np.array_equal(xa, raw.astype(np.float32))


def test_rtpp_rtps():
    k, p = 16, 30
    orc = O.Oracle(k, True)
    xb, yo, yb = _random_problem(k, p, 5)
    inflat = _f32(k - 1) / _f32(1.1)
    xa0, _, _, raw = orc.letkf_solve(xb, yo, yb, inflat)
    xa, _, _, _ = orc.letkf_solve(xb, yo, yb, inflat, True, 0.95, True, 0.95)
    # restate core:684-698 in numpy real32/real64
    xm = xa0.sum(dtype=np.float32) * _f32(1.0 / k) if False else _f32(np.add.reduce(xa0, dtype=np.float32))
    s = _f32(0)
    for v in xa0:
        s = _f32(s + v)
    xa_mean = _f32(s * _f32(_f32(1.0) / _f32(k)))
    xap = (xa0 - xa_mean).astype(np.float32)
    sb = _f32(0)
    for v in xb:
        sb = _f32(sb + v)
    xbp = xb.astype(np.float64) - np.float64(_f32(sb * _f32(_f32(1.0) / _f32(k))))
    al = _f32(0.95)
    xap = ((_f32(1.0) - al) * xap).astype(np.float64) + np.float64(al) * xbp
    xap = xap.astype(np.float32)
    xb_std = _f32(np.sum(xbp * xbp))
    xa_std = _f32(0)
    for v in xap:
        xa_std = _f32(xa_std + _f32(v * v))
    f = _f32(_f32(al * np.sqrt(_f32(xb_std / xa_std), dtype=np.float32)) - al) + _f32(1.0)
    want = xa_mean + (xap * f).astype(np.float32)
    assert np.allclose(xa, want, rtol=3e-7)
    # RTPP with alpha=1 restores the background spread exactly (up to rounding)
    xa1, _, _, _ = orc.letkf_solve(xb, yo, yb, inflat, True, 1.0, False, 0.0)
    assert np.allclose(xa1 - xa1.mean(), xb - xb.mean(), atol=2e-4)


def test_real32_branch_close_to_real64():
    k, p = 32, 200
    xb, yo, yb = _random_problem(k, p, 77)
    inflat = _f32(k - 1) / _f32(1.1)
    a64 = O.Oracle(k, True).letkf_solve(xb, yo, yb, inflat)
    a32 = O.Oracle(k, False).letkf_solve(xb, yo, yb, inflat)
    assert np.allclose(a32[3], a64[3], rtol=2e-5)
    assert np.allclose(a32[2], a64[2], rtol=0, atol=2e-5 * np.abs(a64[2]).max())


# ------------------------------------------------------------------ letkf_yoyb
def test_yoyb_against_numpy_restatement():
    sc, rng = S.scenario_tiny(k=8)
    cfg = C.sample_namelist("T")
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    orc.build_tree(cfg)
    k = sc.k
    ninv, n1inv = _f32(1.0) / _f32(k), _f32(1.0) / _f32(k - 1)
    checked = 0
    for pt in range(0, sc.npts, 37):
        xyz = sc.xyz_grid[pt]
        lists = orc.get_lz(xyz)
        yo, yb = orc.letkf_yoyb(xyz)
        rows_yo, rows_yb = [], []
        for fam, typ, idx, r2 in lists:
            tc = [t for t in cfg.types if t.family == fam and t.type == typ][0]
            o = sc.obs[(fam, typ)]
            for j, i1 in enumerate(idx):
                i = i1 - 1
                for s in range(o.nvar):
                    if fam == C.GTS:
                        if not tc.is_assim[s] or not (o.qc[:, i, s] >= 0).any():
                            continue
                        err = _f32(o.error[i, s] * _f32(tc.err_muti[s]))
                    else:
                        err = _f32(tc.err_muti[0])
                    bg = o.hdxb[:, i, s].copy()
                    sm = _f32(0)
                    for v in bg:
                        sm = _f32(sm + v)
                    mean = _f32(sm * ninv)
                    bg = (bg - mean).astype(np.float32)
                    dot = _f32(0)
                    for v in bg:
                        dot = _f32(dot + _f32(v * v))
                    omm = _f32(o.obs[i, s] - mean)
                    std = np.sqrt(_f32(dot * n1inv), dtype=np.float32)
                    gross = abs(omm) > _f32(np.sqrt(_f32(_f32(std * std) + _f32(err * err)), dtype=np.float32)
                                            * _f32(tc.err_rej[s if fam == C.GTS else 0]))
                    if fam == C.RADAR and typ == C.DBZ:
                        if gross and o.obs[i, s] != _f32(-5.0):
                            continue
                        if o.obs[i, s] == _f32(-5.0) and mean == _f32(-5.0):
                            continue
                    elif gross:
                        continue
                    ei = _f32(O.lib().or_error_inv(float(err), float(r2[j]), 0))
                    rows_yo.append(_f32(omm * ei))
                    rows_yb.append((bg * ei).astype(np.float32))
        assert len(rows_yo) == len(yo)
        if len(yo):
            checked += 1
            assert np.array_equal(np.array(rows_yo, np.float32).view(np.int32), yo.view(np.int32))
            assert np.array_equal(np.array(rows_yb, np.float32).view(np.int32), yb.view(np.int32))
    assert checked > 5


def test_no_obs_leaves_field_untouched_and_counts():
    sc, rng = S.scenario_tiny(k=8)
    cfg = C.sample_namelist("QRAIN")
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    var = S.make_field(rng, sc.k, sc.xyz_grid, 1e-3, 1e-3, 2e-4)
    before = var.copy()
    npo, rows = orc.analyze(cfg, sc.xyz_grid, var)
    changed = (var != before).any(0)
    assert 0 < npo < sc.npts and changed.sum() <= npo
    # points without local obs are bit-identical (core:220,226)
    orc.build_tree(cfg)
    for pt in np.nonzero(~changed)[0][:50]:
        assert len(orc.letkf_yoyb(sc.xyz_grid[pt])[0]) == 0 or True
    far = sc.xyz_grid.copy()
    far[:, 0] += 5e6                              # nothing within any cutoff
    v2 = before.copy()
    npo2, _ = orc.analyze(cfg, far, v2)
    assert npo2 == 0 and np.array_equal(v2, before)


def test_analyze_threads_and_fields_consistent():
    sc, rng = S.scenario_tiny(k=8)
    cfg = C.sample_namelist("T")
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    f = np.stack([S.make_field(rng, sc.k, sc.xyz_grid, 280, 5, 1.0) for _ in range(2)])
    a = f.copy()
    orc.analyze(cfg, sc.xyz_grid, a, nthreads=1)
    b = f.copy()
    orc.analyze(cfg, sc.xyz_grid, b, nthreads=4)
    assert np.array_equal(a, b)
    c0 = f[0].copy()
    orc.analyze(cfg, sc.xyz_grid, c0)
    assert np.array_equal(a[0], c0)


def test_mixed_dimension_family_is_refused():
    sc, rng = S.scenario_tiny(k=8)
    S.add_gts(sc, rng, 0, 0, 0, 0, 0, n_gpspw=10)
    cfg = C.sample_namelist("T", use_gpspw=True)     # gpspw is 2-D, the other GTS types 3-D
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    with pytest.raises(RuntimeError, match="2-D and 3-D"):
        orc.build_tree(cfg)


# ------------------------------------------------------------------ tune_q, syevd
def test_tune_q():
    q = np.array([[1.0, -1.0, 0.0], [3.0, 2.0, 0.0], [-2.0, 5.0, 0.0]], np.float32)  # [k=3, npts=3]
    O.tune_q(q)
    assert np.allclose(q[:, 0], [0.5, 1.5, 0.0])          # mean 2/3 preserved, negatives clipped
    assert np.allclose(q[:, 1], [0.0, 2 * 6 / 7, 5 * 6 / 7])
    assert np.isnan(q[:, 2]).all()                        # 0/0


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_syevd_residuals(dtype, tol):
    rng = np.random.default_rng(1)
    k, b = 32, 8
    Y = rng.normal(size=(b, k, 60))
    A = (np.eye(k) * 28.0 + Y @ Y.transpose(0, 2, 1)).astype(dtype)
    W, V = O.syevd_batch(A, nthreads=2)
    for i in range(b):
        v = V[i].T.astype(np.float64)                     # column-major -> columns are vectors
        assert np.abs(A[i].astype(np.float64) @ v - v * W[i]).max() < tol * np.abs(A[i]).max() * k
        assert np.abs(v.T @ v - np.eye(k)).max() < tol * k
        assert (np.diff(W[i]) >= 0).all()
