#!/usr/bin/env python
"""Benchmark of the LETKF local-analysis hot path (BASELINE.json metric: analysed grid points/s).

A *step* is one pass of the hot path -- the loop body of letkf_driver for every grid point of one
3-D variable -- over BASELINE config M: 450x450x50 grid, 32 members, ~10^6 radar + ~10^4 GTS
observations, variable `T` (GTS + radial velocity, the heaviest localisation mix of input.nml).

  value        : swept grid points/s with grid + ensemble resident in HBM (letkf_b200_analyze_dev),
                 timed with CUDA events on the library's stream, max over ranks.
  e2e          : same metric through the host-pointer C-ABI call (letkf_b200_analyze) with pinned
                 HOST buffers; H2D of xyz+ensemble and D2H of the analysis inside the timed region.
  roofline     : dominant kernel against the FP64 FMA peak measured by the committed micro-benchmark
                 (letkf_b200_fma_peak) or the HBM copy bandwidth of MEASURED_PEAKS.json; per-stage
                 numbers in `stages`.
  cpu_baseline : the oracle (a port of the reference algorithm; the Fortran reference cannot be
                 compiled here) on the host cores over a bounded sample of the same workload.
  parity       : the GPU path through the C ABI against the oracle on the SAME sampled points and field
                 the cpu_baseline leg just analysed (lists, yo/Yb rows, weights, field, NaN sites); a
                 violated bar makes the run exit with status 3 after printing the line.
  secondary    : the other two thirds of BASELINE.json's metric, time bounded: config L (256 members)
                 on a 96x96x50 sub-grid, and config E (batched eigensolves/s, distinct matrices).

`--impl reference` times that CPU restatement alone.  N > 1 (torchrun): grid columns are
partitioned cyclically across ranks as in module_mpi_util.f90:73-188, observations replicated with
the member-sliced NCCL all-gather that mirrors module_gts_omboma.f90:601-605; strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VAR = os.environ.get("LETKF_BENCH_VAR", "T")
EXIT_CODE = 0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=32)
    ap.add_argument("--nx", type=int, default=450)
    ap.add_argument("--ny", type=int, default=450)
    ap.add_argument("--nz", type=int, default=50)
    ap.add_argument("--var", default=VAR)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-secondary", action="store_true", help="skip the config L / config E block")
    ap.add_argument("--localization", default="nml", choices=["nml", "3d"],
                    help="3d: BASELINE config 4 -- every active observation type localised in 3-D (vclr > 0; types "
                         "that input.nml localises in 2-D get vclr = 3 km) with max_lz_pts = 300 (README TODO, SURVEY Q17)")
    ap.add_argument("--workload", default="T", choices=["T", "cycle16"],
                    help="T: the hot path over one 3-D variable (BASELINE config M, the headline); cycle16: one full "
                         "analysis cycle of all 16 variables from host buffers incl. the field exchange (bench_cycle.py)")
    ap.add_argument("--parity-points", type=int, default=256,
                    help="points of the cpu_baseline sample on which lists / rows / weights are compared")
    ap.add_argument("--nxb", type=int, default=0,
                    help="block size along x of the block-cyclic column decomposition (the reference's nxb, "
                         "module_mpi_util.f90:10; results do not depend on it); 0 = the largest of 16, 8, .. 1 "
                         "that keeps the ranks balanced")
    ap.add_argument("--nyb", type=int, default=1)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.period_ms = int(os.environ.get("LETKF_BENCH_SMI_MS", "200"))   # the recipe's clocks line: -lms 200

    def start(self):
        if os.environ.get("LETKF_BENCH_NO_SMI"):   # experiment knob: shows what the in-run polling itself costs
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(sc, cfg, seconds: float, threads: int, keep=None):
    """Oracle (port of the reference hot path) on the host cores over a bounded random sample of the
    workload's grid points, full observation set.  Returns points/s and what the sample was."""
    from oracle import oracle as O
    orc = O.Oracle(sc.k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    rng = np.random.default_rng(11)
    from cwbnwp_letkf_b200 import synthetic as S

    def run_on(n, xyz, f0, f):
        orc.build_tree(cfg)                       # build time excluded, like the GPU tree cache
        t0 = time.perf_counter()
        from oracle.oracle import lib, to_c, _p
        import ctypes
        npo, rows = ctypes.c_int64(0), ctypes.c_int64(0)
        rc = lib().or_analyze(ctypes.c_void_p(orc.h), ctypes.byref(orc.ccfg), ctypes.c_int64(n), _p(xyz), 1,
                              _p(f), threads, ctypes.byref(npo), ctypes.byref(rows))
        dt = time.perf_counter() - t0
        assert rc == 0
        return dt, npo.value, rows.value, (xyz, f0, f)

    def run(n):
        sel = np.sort(rng.choice(sc.npts, n, replace=False))
        xyz = np.ascontiguousarray(sc.xyz_grid[sel])
        f0 = S.make_field(rng, sc.k, xyz, 280.0, 5.0, 1.0)
        return run_on(n, xyz, f0, f0.copy())

    pilot = min(2000, sc.npts)
    dt, _, _, _ = run(pilot)
    n = int(min(sc.npts, max(pilot, pilot * seconds / max(dt, 1e-3))))
    dt, npo, rows, sample = run(n)
    if keep is not None:
        keep["orc"], keep["sample"], keep["npo"], keep["rows"] = orc, sample, npo, rows
    return {"value": n / dt, "unit": "grid points/s", "cores": threads, "kind": "port",
            "sample": f"{n} random grid points of the {sc.nx}x{sc.ny}x{sc.nz} grid, full obs set, variable {cfg_name(cfg)}, "
                      f"{npo} analysed, {rows / max(npo, 1):.0f} rows/point, {dt:.1f} s; oracle = C++ port of the "
                      "reference loop with OpenBLAS dsyrk/dsyevd/dgemm, one OpenMP thread per core",
            "seconds": dt}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def stage_report(stats, k: int, fma64: float, members: int):
    """Per-stage achieved rates against their rooflines (SURVEY.md 8(d) algorithmic counts) and the roofline
    object of the dominant stage.  `solve` is the kernel that replaces module_eigen + the weight application:
    its contract figure uses the 4k^3 LAPACK model of SURVEY 8(d); the flops it really executes
    (4/3 k^3 tridiagonalisation + O(k^2) per vector) are reported beside it."""
    peaks = load_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    units, rows0 = stats.units, stats.rows
    ntree_bytes = 12 * stats.npts + 8 * rows0   # xyz in + (idx, r2) written for kept entries
    ms_solve = stats.ms_eigen + stats.ms_transform
    executed = (4.0 / 3.0) * k**3 + 2 * 8.0 * k * k + 3 * 32 * 8.0 * k   # tridiagonalisation + Q applications + poles
    stage = {
        "search": {"ms": stats.ms_search, "bound": "hbm", "achieved": ntree_bytes / max(stats.ms_search, 1e-9) / 1e6,
                   "peak": hbm, "unit": "GB/s"},
        "gram": {"ms": stats.ms_gram, "bound": "fp64",
                 "achieved": (k * (k + 1) + 2 * k) * rows0 / max(stats.ms_gram, 1e-9) / 1e9, "peak": fma64,
                 "unit": "TFLOP/s", "gathered_GBs": (4 * k + 8) * rows0 / max(stats.ms_gram, 1e-9) / 1e6},
        "solve": {"ms": ms_solve, "bound": "fp64", "achieved": 4.0 * k**3 * units / max(ms_solve, 1e-9) / 1e9,
                  "peak": fma64, "unit": "TFLOP/s", "solves_per_s": units / max(ms_solve, 1e-9) * 1e3,
                  "model": "4k^3 flop per unit (SURVEY 8(d): ?syevd + 2 ?gemm, the reference's count)",
                  "executed_TFLOPs": executed * units / max(ms_solve, 1e-9) / 1e9,
                  "executed_model": "4/3 k^3 (Householder tridiagonalisation) + 16 k^2 (Q^T, Q on two vectors) + "
                                    "768 k (32 shifted tridiagonal solves) per unit: no eigendecomposition",
                  "includes": "weight application + RTPP/RTPS epilogue (xb in, xa out: 8k bytes per unit and field)"},
        "tree_build_ms": stats.ms_tree,
    }
    for s_ in ("search", "gram", "solve"):
        stage[s_]["frac"] = stage[s_]["achieved"] / stage[s_]["peak"] if stage[s_]["ms"] > 0 else None
        stage[s_]["share_of_step"] = stage[s_]["ms"] / max(stats.ms_total, 1e-9)
    # DRAM traffic per launch: NOT measured in this run -- taken from the committed ncu --set full capture
    try:
        tt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        tt = {}
    tab = tt.get("k%d" % members, {})
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from tree_stamp import tree_stamp
        if tt.get("stamp") != tree_stamp():
            tab = {}            # the committed capture is from another tree: its numbers are not this run's
    except Exception:
        tab = {}
    # units one launch processes = one pipeline chunk: at most 2^18 grid points, fewer for large k (the library sizes
    # chunks against a 40 GB budget for the per-unit matrices, api.cu pick_chunk)
    chunk_est = max(1024, min(1 << 18, int((40 << 30) / (16 + (k * k + 3 * k) * 8 + 8192))))
    per_launch_units = min(units, chunk_est)
    for s_ in ("search", "gram", "solve"):
        t_ = tab.get(s_)
        stage[s_]["traffic_bytes_per_launch"] = t_["bytes_per_unit"] * per_launch_units if t_ else None
        stage[s_]["algorithmic_bytes_per_unit"] = t_.get("algorithmic_bytes_per_unit") if t_ else None
    dom = max(("search", "gram", "solve"), key=lambda s_: stage[s_]["ms"])
    roofline = {"kernel": dom, "bound": stage[dom]["bound"], "achieved": stage[dom]["achieved"],
                "peak": stage[dom]["peak"], "unit": stage[dom]["unit"], "frac": stage[dom]["frac"],
                "traffic": stage[dom]["traffic_bytes_per_launch"],
                "traffic_source": ("committed ncu --set full capture %s (tree %s), scaled by the units of one launch; "
                                   "not measured in this run" % (tt.get("source", "profiles/ncu_traffic.json"),
                                                                 tt.get("git", "?"))) if tab else None,
                "peak_source": ("letkf_b200_fma_peak micro-benchmark (FP64 FMA, measured in this run; "
                                "MEASURED_PEAKS.json has no FP64 figure)" if stage[dom]["bound"] == "fp64" else
                                ("MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)")),
                "share_of_step": stage[dom]["share_of_step"]}
    return stage, roofline


def parity_block(eng, cfg, keep, npoints: int):
    """GPU (through the C ABI) against the oracle on the sample the cpu_baseline leg analysed."""
    from oracle import parity as PAR
    orc = keep["orc"]
    xyz, f0, fref = keep["sample"]
    got = f0.copy()
    st = eng.analyze(cfg, xyz, got)
    par = PAR.compare_fields(got, fref, f0)
    par.update(points=int(xyz.shape[0]), analysed=int(st.npts_analysed),
               counts_equal=bool(st.npts_analysed == keep["npo"] and st.rows == keep["rows"]))
    n = min(npoints, xyz.shape[0])
    sel = np.sort(np.random.default_rng(5).choice(xyz.shape[0], n, replace=False))
    pp = PAR.point_parity(eng, orc, cfg, np.ascontiguousarray(xyz[sel]), np.ascontiguousarray(f0[:, sel]),
                          strict=False)
    par.update(weight_points=pp["analysed"], lists_bit_equal=pp["lists_bit_equal"], yoyb_bit_equal=pp["yoyb_bit_equal"],
               wbar_Wa_max_rel=max(pp["max_rel_wbar"], pp["max_rel_Wa"]), xa_raw_max_rel=pp["max_rel_raw"],
               rows_per_point=pp["rows"] / max(pp["analysed"], 1))
    par["ok"] = bool(par["counts_equal"] and par["nan_sites_equal"] and par["untouched_bit_identical"] and
                     par["lists_bit_equal"] and par["yoyb_bit_equal"] and par["wbar_Wa_max_rel"] < 1e-10 and
                     par["xa_raw_max_rel"] < 1e-10 and par["field_max_rel"] <= 5e-7)
    par["bars"] = "lists / rows bit-exact, wbar Wa xa_raw < 1e-10, field < 5e-7, NaN sites equal"
    return par


def secondary_L(device: int, budget_s: float = 60.0):
    """Config L (256 members) on the 96x96x50 sub-grid of config M, variable T: one timed step."""
    import torch
    from cwbnwp_letkf_b200 import config as C
    from cwbnwp_letkf_b200 import host as H
    from cwbnwp_letkf_b200 import synthetic as S
    t0 = time.perf_counter()
    k = 256
    sc, rng = S.scenario_M(k=k, nx=96, ny=96, nz=50)
    eng = H.LetkfB200(k, True, device)
    for o in sc.obs.values():
        eng.set_obs(o)
    cfg = C.sample_namelist("T")
    cfg.tune_q = False
    dev = torch.device("cuda", device)
    f0 = S.make_field(np.random.default_rng(77), k, sc.xyz_grid, 280.0, 5.0, 1.0)
    d_xyz, d_f0 = torch.from_numpy(sc.xyz_grid).to(dev), torch.from_numpy(f0).to(dev)
    d_var = torch.empty_like(d_f0)
    ccfg = C.to_c(cfg)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    fma64 = eng.fma_peak(0)
    t_setup = time.perf_counter() - t0
    ms, st, steps = [], None, 0
    for it in range(3):                       # 1 warm-up + up to 2 timed steps inside the budget
        d_var.copy_(d_f0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = eng.analyze_ptr(ccfg, sc.npts, d_xyz.data_ptr(), 1, d_var.data_ptr(), dev=True)
        e1.record(stream)
        e1.synchronize()
        if it > 0:
            ms.append(e0.elapsed_time(e1))
        if it > 0 and time.perf_counter() - t0 > budget_s:
            break
    stage, roofline = stage_report(st, k, fma64, k)
    out = {"workload": "L: 96x96x50 sub-grid of config M (same generator: %d obs values), k=256, variable T" %
                       sc.total_obs_values(),
           "metric": "analysed grid points/s", "value": sc.npts / (np.mean(ms) * 1e-3), "unit": "grid points/s",
           "ms_per_step": float(np.mean(ms)), "steps": len(ms), "warmup": 1, "dtype": "f64",
           "rows_per_analysed_point": st.rows / max(st.npts_analysed, 1), "roofline": roofline, "stages": stage,
           "setup_s": t_setup}
    eng.finalize()
    del d_xyz, d_f0, d_var
    torch.cuda.empty_cache()
    return out


def secondary_E(device: int, budget_s: float = 40.0):
    """Config E: batched eigensolves/s (values + vectors) over DISTINCT matrices, FP64 and FP32."""
    import bench_eig
    out = []
    cases = [(k, dt) for dt in ("f64", "f32") for k in (32, 64, 128, 256)]
    per = budget_s / len(cases)
    for k, dt in cases:
        r = bench_eig.run_case(k, dt, total=1_000_000, budget_s=0.6 * per, sample=64, cpu=False, device=device)
        out.append({q: r[q] for q in ("k", "dtype", "value", "unit", "matrices_solved", "distinct", "max_sweeps",
                                      "frac_of_fma_peak_4k3_model", "fma_peak_tflops", "check")})
    return out


def cfg_name(cfg):
    return getattr(cfg, "_name", "?")


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from cwbnwp_letkf_b200 import config as C
    from cwbnwp_letkf_b200 import partition as P
    from cwbnwp_letkf_b200 import synthetic as S

    if a.nxb <= 0:
        a.nxb = P.auto_block(a.nx, P.process_grid(world)[0])
    cfg = C.sample_namelist(a.var)
    cfg.tune_q = False if a.var == "T" else cfg.tune_q
    cfg._name = a.var
    if a.localization == "3d":
        for t in cfg.types:
            if t.use_it and t.hclr > 0:
                t.vclr = t.vclr if t.vclr > 0 else 3.0
                t.max_lz_pts = 300
    workload = f"M: {a.nx}x{a.ny}x{a.nz} grid, dx=2km, k={a.members}, variable {a.var}" + \
        (", 3-D localisation on every type, max_lz_pts 300 (config 3D)" if a.localization == "3d" else "")
    config = {"workload": workload, "members": a.members, "variable": a.var,
              "namelist": "input.nml (hclr/vclr/max_lz_pts/inflation/RTPP/RTPS as shipped)",
              "l2": "inputs larger than L2 (ensemble field %.2f GB)" % (a.nx * a.ny * a.nz * a.members * 4 / 1e9),
              "partition": "1 GPU" if world == 1 else
              "columns block-cyclic (nxb=%d, nyb=%d) over a %dx%d process grid, obs replicated" %
              ((a.nxb, a.nyb) + P.process_grid(world))}

    if a.workload == "cycle16":
        import bench_cycle
        if a.impl == "reference":
            if rank == 0:
                print(json.dumps(bench_cycle.run_reference(a)))
            return
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        out = bench_cycle.run(a, rank, world, local_rank, ClockSampler)
        if rank == 0:
            print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return

    if a.impl == "reference":
        if rank != 0:
            return
        sc, rng = S.scenario_M(k=a.members, nx=a.nx, ny=a.ny, nz=a.nz)
        threads = os.cpu_count() or 1
        vals = []
        per_step = max(2.0, min(a.cpu_seconds, 60.0 / max(a.steps + a.warmup, 1)))
        last = None
        for i in range(a.warmup + a.steps):
            last = cpu_baseline(sc, cfg, per_step, threads)
            if i >= a.warmup:
                vals.append(last["value"])
        v = float(np.mean(vals))
        out = {"impl": "reference", "metric": "analysed grid points/s", "value": v, "unit": "grid points/s",
               "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": 1e3 * (a.nx * a.ny * a.nz) / v, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": v, "unit": "grid points/s", "cores": threads, "kind": "port",
                                "sample": last["sample"]},
               "e2e": {"value": v, "unit": "grid points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "note": "the Fortran/MPI reference cannot be compiled in this image (no Fortran compiler, MPI, "
                       "NetCDF, SSL2); this arm times the C++ port (oracle/) on a bounded sample, ms_per_step "
                       "is the extrapolation to the full grid"}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from cwbnwp_letkf_b200 import host as H
    from cwbnwp_letkf_b200 import partition as P

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # ---- synthetic workload (same seed on every rank) ----
    sc, rng = S.scenario_M(k=a.members, nx=a.nx, ny=a.ny, nz=a.nz)
    k = sc.k
    eng = H.LetkfB200(k, True, local_rank)

    # ---- observations: member-sliced all-gather of H(x), like mpi_iallgatherv (gts:601-605, rad:179) ----
    t_obs0 = time.perf_counter()
    for key, o in sc.obs.items():
        if world == 1:
            eng.set_obs(o)
            continue
        n, nv = o.n, o.nvar
        lo, hi = P.member_slice(rank, world, k)
        mine = torch.from_numpy(np.ascontiguousarray(o.hdxb[lo:hi])).to(dev)      # this rank "read" members lo:hi
        full = P.allgather_members(mine, k, rank, world)
        qc = None
        if o.qc is not None:
            mq = torch.from_numpy(np.ascontiguousarray(o.qc[lo:hi])).to(dev)
            qc = P.allgather_members(mq, k, rank, world)
        xyz = torch.from_numpy(o.xyz).to(dev)
        obs = torch.from_numpy(o.obs).to(dev)
        err = None if o.error is None else torch.from_numpy(o.error).to(dev)
        torch.cuda.synchronize()
        eng.set_obs_dev(o.family, o.type, n, nv, xyz, obs, err, full, qc)
        del full, qc, mine
    torch.cuda.synchronize()
    t_obs = time.perf_counter() - t_obs0

    # ---- this rank's columns (2-D cyclic process grid, module_mpi_util.f90:80-127) ----
    pts = P.local_points(rank, world, sc.nx, sc.ny, sc.nz, a.nxb, a.nyb)
    xyz_local = np.ascontiguousarray(sc.xyz_grid[pts]) if world > 1 else sc.xyz_grid
    npts_local = xyz_local.shape[0]
    total_pts = sc.npts
    frng = np.random.default_rng(1234 + rank)
    field0 = S.make_field(frng, k, xyz_local, 280.0, 5.0, 1.0)                      # var[k, npts_local]

    d_xyz = torch.from_numpy(xyz_local).to(dev)
    d_field0 = torch.from_numpy(field0).to(dev)
    d_var = torch.empty_like(d_field0)
    ccfg = C.to_c(cfg)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_step():
        d_var.copy_(d_field0)                # restore the background (not timed: outside the events)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = eng.analyze_ptr(ccfg, npts_local, d_xyz.data_ptr(), 1, d_var.data_ptr(), dev=True)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1), st

    # FP64 / FP32 FMA peaks (roofline denominators the driver does not measure)
    fma64 = eng.fma_peak(0)
    fma32 = eng.fma_peak(1)
    dmma64 = eng.fma_peak(2)       # FP64 tensor pipe alone
    mixed64 = eng.fma_peak(3)      # DMMA + DFMA interleaved: not additive on B200 (one FP64 resource)

    for _ in range(a.warmup):
        dev_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    t_steps, stats = [], None
    barrier()
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        ms, stats = dev_step()
        t_steps.append(ms)
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_local = float(np.sum(t_steps))
    if world > 1:
        t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        cnt = torch.tensor([stats.npts_analysed, stats.rows, launches], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        analysed, rows, launches = int(cnt[0].item()), int(cnt[1].item()), int(cnt[2].item())
    else:
        ms_total, analysed, rows = ms_local, stats.npts_analysed, stats.rows
    ms_per_step = ms_total / a.steps
    value = total_pts / (ms_per_step * 1e-3)

    # ---- end to end through the host-pointer ABI with pinned host buffers ----
    e2e = None
    if not a.no_e2e:
        h_xyz = torch.from_numpy(xyz_local).pin_memory()
        h_field0 = torch.from_numpy(field0).pin_memory()
        h_var = torch.empty_like(h_field0).pin_memory()

        def e2e_step():
            h_var.copy_(h_field0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.analyze_ptr(ccfg, npts_local, h_xyz.data_ptr(), 1, h_var.data_ptr(), dev=False)
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1)

        e2e_step()
        barrier()
        ms_e = float(np.sum([e2e_step() for _ in range(a.steps)]))
        if world > 1:
            t = torch.tensor([ms_e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        # the device result equals the resident-path result
        same = bool(torch.equal(h_var.to(dev), d_var))
        e2e = {"value": total_pts / (ms_e / a.steps * 1e-3), "unit": "grid points/s",
               "h2d_bytes_per_step": int(total_pts * (3 + k) * 4),
               "d2h_bytes_per_step": int(total_pts * k * 4), "ms_per_step": ms_e / a.steps,
               "pipeline_ms_inside": eng.last_stats.ms_total,
               "matches_resident_path": same}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (SURVEY.md 8(d) algorithmic counts) from the last timed step of rank 0 ----
    fma = {"fp64": fma64, "fp32": fma32, "fp64_dmma": dmma64, "fp64_dmma_plus_fma": mixed64}
    stage, roofline = stage_report(stats, k, fma64, a.members)

    cpu, par = None, None
    if not a.no_cpu_baseline:
        keep = {}
        cpu = cpu_baseline(sc, cfg, a.cpu_seconds, os.cpu_count() or 1, keep)
        if world == 1:
            par = parity_block(eng, cfg, keep, a.parity_points)
        del keep

    secondary = None
    if world == 1 and not a.no_secondary and a.members == 32 and (a.nx, a.ny, a.nz) == (450, 450, 50):
        eng.finalize()
        del d_xyz, d_field0, d_var
        torch.cuda.empty_cache()
        secondary = {"config_L_subgrid": secondary_L(local_rank), "config_E_eigensolves": secondary_E(local_rank)}

    out = {"metric": "analysed grid points/s", "value": value, "unit": "grid points/s", "n_gpus": world,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
           "parity": par, "stages": stage, "points_analysed": int(analysed), "analysed_fraction": analysed / total_pts,
           "points_per_s_analysed_only": analysed / (ms_per_step * 1e-3), "rows_per_analysed_point": rows / max(analysed, 1),
           "ms_steps": [round(float(x), 2) for x in t_steps], "fma_peak_tflops": fma, "obs_setup_s": t_obs,
           "wall_s_timed": wall, "obs_values": sc.total_obs_values(), "secondary": secondary}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if par is not None and not par["ok"]:
        global EXIT_CODE
        EXIT_CODE = 3


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner) write
    # to fd 1 behind Python's back, so fd 1 is pointed at stderr while the benchmark runs and restored
    # for the final line.
    sys.stdout.flush()
    _saved = os.dup(1)
    os.dup2(2, 1)
    import io
    _buf = io.StringIO()
    _py_stdout = sys.stdout
    sys.stdout = _buf
    try:
        main()
    finally:
        sys.stdout = _py_stdout
        sys.stdout.flush()
        os.dup2(_saved, 1)
        os.close(_saved)
    lines = [l for l in _buf.getvalue().splitlines() if l.strip()]
    for l in lines[:-1]:
        print(l, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)
    sys.exit(EXIT_CODE)
