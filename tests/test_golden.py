"""Golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle: the Fortran
reference cannot be executed here).  CPU tier: the oracle still reproduces them bit for bit.
GPU tier: the CUDA path reproduces the lists / rows bit for bit and the analysis to real32 rounding, NaNs
(real32 Gaspari-Cohn, tune_q 0/0) at the same places."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402

NAMES = list(MG.CASES)


def _load(name):
    G = np.load(os.path.join(HERE, "golden", name + ".npz"))
    sc, cfg, field = MG.case_inputs(name)
    assert np.array_equal(field, G["field_in"]), "synthetic generator changed: regenerate the golden files"
    return G, sc, cfg, field


def _bits_equal(a, b):
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.int32), b[~nb].view(np.int32))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden_vectors(name):
    from oracle import oracle as O
    G, sc, cfg, field = _load(name)
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    orc.build_tree(cfg)
    for pt in G["pts"]:
        for t, (fam, typ, idx, r2) in enumerate(orc.get_lz(sc.xyz_grid[pt])):
            assert np.array_equal(idx, G[f"idx_{pt}_{t}"])
            assert np.array_equal(r2.view(np.int32), G[f"r2_{pt}_{t}"].view(np.int32))
        yo, yb = orc.letkf_yoyb(sc.xyz_grid[pt])
        assert _bits_equal(yo, G[f"yo_{pt}"]) and _bits_equal(yb, G[f"yb_{pt}"])
    ana = field.copy()
    npo, rows = orc.analyze(cfg, sc.xyz_grid, ana, nthreads=2)
    if cfg.tune_q:
        O.tune_q(ana)
    assert [npo, rows] == list(G["counts"])
    assert _bits_equal(ana, G["analysis"])


def test_golden_cases_exercise_the_hazards():
    """The fixtures are only useful if they contain what they claim to pin."""
    G = np.load(os.path.join(HERE, "golden", "tiny_QRAIN_gc_k8.npz"))
    assert np.isnan(G["analysis"]).any()                         # real32 Gaspari-Cohn / tune_q NaNs
    assert any(np.isnan(G[f"yb_{pt}"]).any() for pt in G["pts"]) or np.isnan(G["analysis"]).any()
    G = np.load(os.path.join(HERE, "golden", "tiny_T_k8.npz"))
    assert max(len(G[k]) for k in G.files if k.startswith("idx_")) == 300   # truncation at max_lz_pts (vr)
    assert G["counts"][0] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_golden_vectors(name):
    from cwbnwp_letkf_b200 import host as H
    G, sc, cfg, field = _load(name)
    eng = H.LetkfB200(sc.k, True)
    for o in sc.obs.values():
        eng.set_obs(o)
    lists = eng.get_lz(cfg, sc.xyz_grid)
    off, yo, yb = eng.letkf_yoyb(cfg, sc.xyz_grid)
    for pt in G["pts"]:
        for t, (fam, typ, cnt, idx, r2) in enumerate(lists):
            assert np.array_equal(idx[pt, :cnt[pt]], G[f"idx_{pt}_{t}"])
            assert np.array_equal(r2[pt, :cnt[pt]].view(np.int32), G[f"r2_{pt}_{t}"].view(np.int32))
        a, b = off[pt], off[pt + 1]
        assert _bits_equal(yo[a:b], G[f"yo_{pt}"]) and _bits_equal(yb[a:b], G[f"yb_{pt}"])
    ana = field.copy()
    st = eng.analyze(cfg, sc.xyz_grid, ana)                       # applies tune_q when cfg.tune_q is set
    assert [st.npts_analysed, st.rows] == list(G["counts"])
    ref = G["analysis"]
    assert np.array_equal(np.isnan(ana), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert (ana[ok] == ref[ok]).mean() > 0.98 and np.abs(ana[ok] - ref[ok]).max() <= 5e-7 * np.abs(ref[ok]).max()
