"""Parity checks of a C-ABI engine (cwbnwp_letkf_b200.host.LetkfB200) against the CPU oracle on the same points.

TEST INFRASTRUCTURE ONLY, like everything under oracle/: used by tests/ and by bench.py's cpu_baseline leg (the
`parity` block of the bench line).  The bars are the north star's: local observation lists bit-exact (here in
kdtree2's visiting order, with bit-exact r2), yo / Yb rows bit-exact real32, wbar / Wa / pre-cast analysis
within 1e-10 (FP64), final real32 field within 5e-7 with NaNs at the same sites and untouched points
bit-identical.
"""
from __future__ import annotations

import numpy as np

TOL64 = 1e-10
TOL_FIELD = 5e-7


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def bits_equal(a, b):
    """Bit-exact real32 equality; NaNs must sit at the same places (their sign / payload is implementation
    defined: x86 sqrt(negative) gives -qNaN, CUDA the canonical one)."""
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a[~na].view(np.int32), b[~nb].view(np.int32)))


def point_parity(eng, orc, cfg, xyz, xb, check_lists=True, strict=True):
    """module_localization.f90:188-331 (lists), module_letkf_core.f90:300-595 (rows), :649-679 (weights) at
    the points xyz[(n,3)] with background xb[(k,n)].  Returns a dict of counts and worst errors; with
    strict=True a violated bar raises AssertionError."""
    n = xyz.shape[0]
    lists = eng.get_lz(cfg, xyz) if check_lists else None
    off, yo, yb = eng.letkf_yoyb(cfg, xyz)
    p, wbar, Wa, raw = eng.letkf_weights(cfg, xyz, xb)
    ntrees = orc.build_tree(cfg)
    inflat = np.float32(eng.k - 1) / np.float32(cfg.multi_infl)
    res = dict(points=n, analysed=0, rows=0, lists_bit_equal=True, yoyb_bit_equal=True, max_rel_wbar=0.0,
               max_rel_Wa=0.0, max_rel_raw=0.0, nan_points=0)
    for i in range(n):
        if check_lists:
            ref = orc.get_lz(xyz[i])
            ok = len(ref) == ntrees == len(lists)
            for t, (fam, typ, idx, r2) in enumerate(ref if ok else []):
                gf, gt, cnt, gidx, gr2 = lists[t]
                ok = ok and (gf, gt) == (fam, typ) and cnt[i] == len(idx) and \
                    np.array_equal(gidx[i, :cnt[i]], idx) and \
                    np.array_equal(gr2[i, :cnt[i]].view(np.int32), r2.view(np.int32))
            res["lists_bit_equal"] = res["lists_bit_equal"] and bool(ok)
            assert ok or not strict, ("local observation lists differ", i)
        ryo, ryb = orc.letkf_yoyb(xyz[i])
        a, b = off[i], off[i + 1]
        same_p = (b - a == len(ryo) == p[i])
        assert same_p or not strict, (i, b - a, len(ryo), p[i])
        if not same_p:
            res["yoyb_bit_equal"] = False
            continue
        if len(ryo) == 0:
            assert (not wbar[i].any() and not Wa[i].any()) or not strict
            continue
        eq = bits_equal(yo[a:b], ryo) and bits_equal(yb[a:b], ryb)
        res["yoyb_bit_equal"] = res["yoyb_bit_equal"] and eq
        assert eq or not strict, ("yo / Yb rows differ", i)
        _, rw, rWa, rraw = orc.letkf_solve(xb[:, i], ryo, ryb, inflat)
        if not np.isfinite(rw).all():  # real32 Gaspari-Cohn NaN rows: NaN sites are compared on the field
            res["nan_points"] += 1
            continue
        res["max_rel_wbar"] = max(res["max_rel_wbar"], relerr(wbar[i], rw))
        res["max_rel_Wa"] = max(res["max_rel_Wa"], relerr(Wa[i], rWa))
        res["max_rel_raw"] = max(res["max_rel_raw"], relerr(raw[i], rraw))
        res["rows"] += len(ryo)
        res["analysed"] += 1
    orc.destroy_tree()
    worst = max(res["max_rel_wbar"], res["max_rel_Wa"], res["max_rel_raw"])
    assert worst < TOL64 or not strict, res
    return res


def compare_fields(got, ref, f0):
    """got / ref: analysed fields [(k,n)], f0 the background.  Returns the parity numbers of the field."""
    nan_equal = bool(np.array_equal(np.isnan(got), np.isnan(ref)))
    ok = ~np.isnan(ref) & ~np.isnan(got)
    changed = (ref != f0).any(0)
    untouched_equal = bool(np.array_equal(got[:, ~changed], f0[:, ~changed]))
    scale = float(np.abs(ref[ok]).max()) if ok.any() else 1.0
    err = float(np.abs(got[ok] - ref[ok]).max() / scale) if ok.any() else 0.0
    same = float((got[ok] == ref[ok]).mean()) if ok.any() else 1.0
    return dict(nan_sites_equal=nan_equal, untouched_bit_identical=untouched_equal, field_max_rel=err,
                field_bit_identical=same)


def field_parity(eng, orc, cfg, xyz, f, nthreads=1, strict=True):
    """The loop body of letkf_driver (core:209-240) over xyz through the C ABI against the oracle."""
    ref = f.copy()
    npo, rows = orc.analyze(cfg, xyz, ref, nthreads=nthreads)
    got = f.copy()
    st = eng.analyze(cfg, xyz, got)
    counts_equal = st.npts == xyz.shape[0] and st.npts_analysed == npo and st.rows == rows
    res = compare_fields(got, ref, f)
    res.update(counts_equal=bool(counts_equal), analysed=int(npo), rows=int(rows))
    if strict:
        assert counts_equal, (st.as_dict(), npo, rows)
        assert res["nan_sites_equal"] and res["untouched_bit_identical"] and res["field_max_rel"] <= TOL_FIELD, res
    return res
