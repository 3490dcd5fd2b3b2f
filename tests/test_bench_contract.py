"""CPU tier: the pieces of the bench contract that need no GPU -- the stage / roofline report, the stamp that ties
capture-derived numbers to the tree they were captured from, the variable grouping and shapes of the cycle16 workload."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import bench  # noqa: E402
import bench_cycle  # noqa: E402
from tree_stamp import tree_stamp  # noqa: E402


class _Stats:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def test_stage_report_uses_the_survey_counts_and_names_the_dominant_stage():
    st = _Stats(units=1000, rows=900000, npts=1200, ms_search=1.0, ms_gram=4.0, ms_eigen=2.0, ms_transform=0.0,
                ms_total=7.5, ms_tree=0.01)
    k = 32
    stage, roof = bench.stage_report(st, k, 36.0, k)
    assert roof["kernel"] == "gram" and roof["bound"] == "fp64" and roof["unit"] == "TFLOP/s"
    # k(k+1)p + 2kp flop for the Gram, 4k^3 per solve, 12 B + 8 B per kept entry for the search
    assert np.isclose(stage["gram"]["achieved"], (k * (k + 1) + 2 * k) * 900000 / 4.0e-3 / 1e12)
    assert np.isclose(stage["solve"]["achieved"], 4.0 * k ** 3 * 1000 / 2.0e-3 / 1e12)
    assert np.isclose(stage["search"]["achieved"], (12 * 1200 + 8 * 900000) / 1.0e-3 / 1e9)
    assert np.isclose(roof["frac"], roof["achieved"] / roof["peak"])
    assert stage["solve"]["executed_TFLOPs"] < stage["solve"]["achieved"]        # 4/3 k^3 executed vs 4 k^3 modelled
    assert abs(sum(stage[s]["share_of_step"] for s in ("search", "gram", "solve")) - 7.0 / 7.5) < 1e-12


def test_capture_derived_numbers_are_tied_to_the_source_tree():
    tt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    st = _Stats(units=1 << 20, rows=10 ** 9, npts=1 << 20, ms_search=1.0, ms_gram=4.0, ms_eigen=2.0, ms_transform=0.0,
                ms_total=7.5, ms_tree=0.01)
    stage, roof = bench.stage_report(st, 32, 36.0, 32)
    if tt.get("stamp") == tree_stamp():
        assert roof["traffic"] == tt["k32"]["gram"]["bytes_per_unit"] * (1 << 18)
        assert "not measured in this run" in roof["traffic_source"] and tt["stamp"] in roof["traffic_source"]
    else:   # a capture from another tree must not be reported
        assert roof["traffic"] is None and roof["traffic_source"] is None
    # large k: one launch processes the library's chunk, far fewer than 2^18 units
    stage, roof = bench.stage_report(st, 256, 36.0, 256)
    if roof["traffic"] is not None:
        assert roof["traffic"] < tt["k256"][roof["kernel"]]["bytes_per_unit"] * 100000


def test_tree_stamp_is_deterministic():
    assert tree_stamp() == tree_stamp() and len(tree_stamp()) == 16


def test_cycle16_shapes_and_groups():
    from cwbnwp_letkf_b200 import config as C
    from cwbnwp_letkf_b200 import driver as D
    sh = bench_cycle.shapes(9, 7, 5)
    assert sh["u"] == (5, 7, 10) and sh["v"] == (5, 8, 9) and sh["w"] == (6, 7, 9) and sh["mu"] == (7, 9) and sh["ph"] == (6, 7, 9)
    assert set(sh) == {D.VARIABLES[n][0] for n in C.VAR_UPDATE}
    geo = bench_cycle.make_geo(9, 7, 2000.0)
    assert geo["xlon"].shape == (9, 7) and geo["xlon_u"].shape == (10, 7) and geo["xlat_v"].shape == (9, 8)
    # staggered points sit half a cell from the mass points
    assert np.isclose(geo["xlon_u"][0, 0], geo["xlon"][0, 0] - 1000.0) and np.isclose(geo["xlat_v"][0, 0], geo["xlat"][0, 0] - 1000.0)
    groups = D.group_variables(C.VAR_UPDATE, C.sample_namelist)
    assert sum(len(g) for g in groups) == 16 and max(len(g) for g in groups) == 8
