"""Regenerates tests/golden/tiny_T_k8.npz.

The reference (Fortran) cannot be run in this environment, so these vectors come from the ORACLE
(oracle/letkf_oracle.cpp) on the seeded `scenario_tiny(k=8)` case; they freeze the oracle's behaviour
(a change that alters local observation lists, yo/Yb rows or analyses shows up as a diff) and give the
GPU tier a fixture that does not depend on building the oracle.  Usage: python tests/golden/make_golden.py
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


def build():
    sc, rng = S.scenario_tiny(k=8)
    cfg = C.sample_namelist("T")
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    field = S.make_field(rng, sc.k, sc.xyz_grid, 280.0, 5.0, 1.0)
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
    out["analysis"] = ana
    out["counts"] = np.array([npo, rows])
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_T_k8.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path), "bytes")
