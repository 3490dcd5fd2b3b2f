"""Golden vectors (tests/golden/tiny_T_k8.npz, produced by tests/golden/make_golden.py from the oracle:
the Fortran reference cannot be executed here).  CPU tier: the oracle still reproduces them bit for bit.
GPU tier: the CUDA path reproduces the lists / rows bit for bit and the analysis to real32 rounding."""
import os

import numpy as np
import pytest

from cwbnwp_letkf_b200 import config as C
from cwbnwp_letkf_b200 import synthetic as S

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "tiny_T_k8.npz"))


def _case():
    sc, rng = S.scenario_tiny(k=8)
    field = S.make_field(rng, sc.k, sc.xyz_grid, 280.0, 5.0, 1.0)
    assert np.array_equal(field, G["field_in"]), "synthetic generator changed: regenerate the golden file"
    return sc, C.sample_namelist("T"), field


def test_oracle_reproduces_golden_vectors():
    from oracle import oracle as O
    sc, cfg, field = _case()
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    orc.build_tree(cfg)
    for pt in G["pts"]:
        for t, (fam, typ, idx, r2) in enumerate(orc.get_lz(sc.xyz_grid[pt])):
            assert np.array_equal(idx, G[f"idx_{pt}_{t}"])
            assert np.array_equal(r2.view(np.int32), G[f"r2_{pt}_{t}"].view(np.int32))
        yo, yb = orc.letkf_yoyb(sc.xyz_grid[pt])
        assert np.array_equal(yo.view(np.int32), G[f"yo_{pt}"].view(np.int32))
        assert np.array_equal(yb.view(np.int32), G[f"yb_{pt}"].view(np.int32))
    ana = field.copy()
    npo, rows = orc.analyze(cfg, sc.xyz_grid, ana, nthreads=2)
    assert [npo, rows] == list(G["counts"])
    assert np.array_equal(ana.view(np.int32), G["analysis"].view(np.int32))


@pytest.mark.gpu
def test_cuda_path_reproduces_golden_vectors():
    from cwbnwp_letkf_b200 import host as H
    sc, cfg, field = _case()
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
        assert np.array_equal(yo[a:b].view(np.int32), G[f"yo_{pt}"].view(np.int32))
        assert np.array_equal(yb[a:b].view(np.int32), G[f"yb_{pt}"].view(np.int32))
    ana = field.copy()
    st = eng.analyze(cfg, sc.xyz_grid, ana)
    assert [st.npts_analysed, st.rows] == list(G["counts"])
    ref = G["analysis"]
    assert (ana == ref).mean() > 0.98 and np.abs(ana - ref).max() <= 5e-7 * np.abs(ref).max()
