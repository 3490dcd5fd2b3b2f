"""Regenerates tests/golden/*.npz.

The reference (Fortran) cannot be run in this environment, so these vectors come from the ORACLE
(oracle/letkf_oracle.cpp) on the seeded `scenario_tiny(k=8)` case; they freeze the oracle's behaviour
(a change that alters local observation lists, yo/Yb rows or analyses shows up as a diff) and give the
GPU tier fixtures that do not depend on building the oracle.  Cases: variable T (GTS + Vr, 3-D
localisation, Gaussian weights), QRAIN with Gaspari-Cohn weights (dBZ no-rain rules, real32 GC NaNs,
letkf_tune_q), P (2-D localisation).  Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cwbnwp_letkf_b200 import config as C  # noqa: E402
from cwbnwp_letkf_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

PTS = [0, 7, 100, 333, 500, 719]
# name -> (variable, weight_function, field mean / amplitude / member spread)
CASES = {
    "tiny_T_k8": ("T", 0, (280.0, 5.0, 1.0)),
    "tiny_QRAIN_gc_k8": ("QRAIN", 1, (1e-3, 1e-3, 5e-4)),
    "tiny_P_k8": ("P", 0, (8.0e4, 500.0, 100.0)),
}


def case_inputs(name):
    var, wf, (mean, amp, spread) = CASES[name]
    sc, rng = S.scenario_tiny(k=8)
    cfg = C.sample_namelist(var, weight_function=wf)
    field = S.make_field(rng, sc.k, sc.xyz_grid, mean, amp, spread)
    return sc, cfg, field


def build(name):
    sc, cfg, field = case_inputs(name)
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    out = {"pts": np.array(PTS), "field_in": field}
    orc.build_tree(cfg)
    for pt in PTS:
        for t, (fam, typ, idx, r2) in enumerate(orc.get_lz(sc.xyz_grid[pt])):
            out[f"idx_{pt}_{t}"] = idx
            out[f"r2_{pt}_{t}"] = r2
        yo, yb = orc.letkf_yoyb(sc.xyz_grid[pt])
        out[f"yo_{pt}"], out[f"yb_{pt}"] = yo, yb
    ana = field.copy()
    npo, rows = orc.analyze(cfg, sc.xyz_grid, ana, nthreads=1)
    if cfg.tune_q:
        O.tune_q(ana)                                   # core:252-278
    out["analysis"] = ana
    out["counts"] = np.array([npo, rows])
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        path = os.path.join(here, name + ".npz")
        np.savez_compressed(path, **build(name))
        print("wrote", path, os.path.getsize(path), "bytes")
