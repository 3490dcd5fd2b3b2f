#!/usr/bin/env python
"""`bench.py --workload cycle16`: one full analysis cycle of letkf_driver (module_letkf_core.f90:21-297) -- all 16
variables of input.nml:7 on the config-M grid (450x450x50, 32 members, ~10^6 observations) -- end to end from
pinned HOST buffers, with the field transposition of letkf_scatter_grid / letkf_gather_grid
(module_mpi_util.f90:190-358) INSIDE the timed region.

Every rank holds its members of the full grid (member-major, like the reference after reading the ensemble).
One timed step = for every group of variables (driver.group_variables: the eight hydrometeor variables form one
group): H2D of the rank's member slabs -> scatter (NCCL all-to-all; a slice on one GPU) -> local analysis
(letkf_b200_analyze_dev, nfields = len(group), per-column weight sharing for MU / P / PH) incl. letkf_tune_q ->
gather -> D2H.  The metric stays BASELINE.json's `analysed grid points/s`, here summed over the 16 variables;
`seconds_per_cycle` is the number an operator reads.  bench.py imports run() / run_reference()."""
from __future__ import annotations

import os
import time

import numpy as np

G = 9.81


class XYProjection:
    """The synthetic grid is defined in metres: 'lat/lon' hold y/x and the projection is the identity.  (At the
    C ABI the host passes x, y it computed with module_projection.f90:37-50 itself.)"""

    def lonlat_to_xy(self, lon, lat):
        return np.asarray(lon, np.float32), np.asarray(lat, np.float32)


def shapes(nx, ny, nz):
    sh = {"u": (nz, ny, nx + 1), "v": (nz, ny + 1, nx), "w": (nz + 1, ny, nx), "mu": (ny, nx), "ph": (nz + 1, ny, nx)}
    for key in ("t", "qv", "qr", "qs", "qg", "qh", "nqr", "nqs", "nqg", "nqh", "p"):
        sh[key] = (nz, ny, nx)
    return sh


MEAN = {"u": (5.0, 2.0), "v": (-3.0, 2.0), "w": (0.0, 0.5), "t": (290.0, 1.5), "p": (8.0e4, 300.0), "mu": (9.0e4, 200.0)}


def make_geo(nx, ny, dx):
    def axis(n, off):
        return ((np.arange(n, dtype=np.float64) - (n - 1) / 2 + off) * dx).astype(np.float32)
    geo = {}
    for sfx, (nxx, nyy, ox, oy) in {"": (nx, ny, 0.0, 0.0), "_u": (nx + 1, ny, 0.0, 0.0), "_v": (nx, ny + 1, 0.0, 0.0)}.items():
        # nx + 1 staggered points centred on the same centre sit exactly half a cell off the nx mass points
        xs = axis(nx, 0.0) if nxx == nx else axis(nx + 1, 0.0)
        ys = axis(ny, 0.0) if nyy == ny else axis(ny + 1, 0.0)
        X, Y = np.meshgrid(xs, ys, indexing="ij")
        geo["xlon" + sfx], geo["xlat" + sfx] = X.astype(np.float32), Y.astype(np.float32)
    span = max(nx, ny) * dx
    geo["hgt"] = (750.0 + 750.0 * np.sin(2 * np.pi * geo["xlon"] / span) * np.cos(2 * np.pi * geo["xlat"] / span)).astype(np.float32)
    return geo


def make_state(torch, dev, k, nx, ny, nz, geo, lo, hi, pinned=True):
    """Pinned host copy of this rank's members [lo, hi) of all 16 fields, generated on the device."""
    host = {}
    ter = torch.from_numpy(np.ascontiguousarray(geo["hgt"].T)).to(dev)                       # [ny, nx]
    zw = torch.tensor(20000.0 * (np.arange(nz + 1) / nz) ** 1.5, dtype=torch.float32, device=dev)
    for key, sh in shapes(nx, ny, nz).items():
        g = torch.Generator(device=dev)
        g.manual_seed(977 + 31 * sorted(shapes(nx, ny, nz)).index(key))
        # members are generated one by one from a per-variable seed: every world size sees the same ensemble
        out = torch.empty((hi - lo,) + sh, dtype=torch.float32, pin_memory=pinned)
        for m in range(hi):
            noise = torch.randn(sh, generator=g, device=dev, dtype=torch.float32)
            if m < lo:
                continue
            if key == "ph":
                f = (ter[None] + zw[:, None, None]) * G + 30.0 * noise
            elif key in MEAN:
                f = MEAN[key][0] + MEAN[key][1] * noise
            else:  # moisture / hydrometeor / number variables: non-negative with clear-air zeros
                f = torch.clamp(1e-4 + 3e-4 * noise, min=0.0)
            out[m - lo].copy_(f)
        host[key] = out
    torch.cuda.synchronize()
    return host


def run(a, rank, world, local_rank, sampler_cls):
    import torch
    import torch.distributed as dist
    from cwbnwp_letkf_b200 import config as C
    from cwbnwp_letkf_b200 import cycle as CY
    from cwbnwp_letkf_b200 import driver as D
    from cwbnwp_letkf_b200 import host as H
    from cwbnwp_letkf_b200 import partition as P
    from cwbnwp_letkf_b200 import synthetic as S

    dev = torch.device("cuda", local_rank)
    k, nx, ny, nz = a.members, a.nx, a.ny, a.nz
    sc, rng = S.scenario_M(k=k, nx=nx, ny=ny, nz=nz)
    eng = H.LetkfB200(k, True, local_rank)
    for o in sc.obs.values():          # replicated observations (bench.py times the member-sliced all-gather variant)
        eng.set_obs(o)
    geo = make_geo(nx, ny, sc.dx)
    lo, hi = P.member_slice(rank, world, k)
    h_state = make_state(torch, dev, k, nx, ny, nz, geo, lo, hi)
    h_out = {key: torch.empty_like(t).pin_memory() for key, t in h_state.items()}
    d_state = {key: torch.empty(t.shape, dtype=torch.float32, device=dev) for key, t in h_state.items()}
    nxb = a.nxb if a.nxb > 0 else P.auto_block(nx, P.process_grid(world)[0])
    names = list(C.VAR_UPDATE)
    groups = D.group_variables(names, C.sample_namelist)
    h2d = sum(t.numel() * 4 for t in h_state.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def one_cycle():
        """Host slabs -> device on a copy stream one group ahead of the analysis; results -> host on a second copy
        stream behind it.  The CUDA events bracket everything, the last download included."""
        cyc = CY.DeviceCycle(eng, C.sample_namelist, XYProjection(), rank, world, nxb, a.nyb)
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        ready = {}

        def upload(gi):
            if gi >= len(groups) or gi in ready:
                return
            with torch.cuda.stream(s_in):
                for nm in groups[gi]:
                    key = D.VARIABLES[nm][0]
                    if key != "ph":
                        d_state[key].copy_(h_state[key], non_blocking=True)
                ready[gi] = torch.cuda.Event()
                ready[gi].record(s_in)

        with torch.cuda.stream(s_in):
            d_state["ph"].copy_(h_state["ph"], non_blocking=True)        # the vertical coordinate needs PH first
        upload(0)

        def h2d_group(group):
            gi = groups.index(group)
            upload(gi)
            cur.wait_event(ready[gi])
            upload(gi + 1)                                               # next group's slabs move under this analysis

        def d2h_group(group):
            done = torch.cuda.Event()
            done.record(cur)
            s_out.wait_event(done)
            with torch.cuda.stream(s_out):
                for nm in group:
                    key = D.VARIABLES[nm][0]
                    h_out[key].copy_(d_state[key], non_blocking=True)

        cyc.run(d_state, geo, names, before_group=h2d_group, after_group=d2h_group)
        cur.wait_stream(s_out)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1), cyc

    for _ in range(a.warmup):
        one_cycle()
    barrier()
    sampler = sampler_cls(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    ms, cyc = [], None
    barrier()
    for _ in range(a.steps):
        t, cyc = one_cycle()
        ms.append(t)
    barrier()
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    tot = torch.tensor([float(np.sum(ms)), cyc.ms_exchange, cyc.ms_analysis, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, ms_ex, ms_an, launches = float(mx[0]), float(mx[1]), float(mx[2]), int(sm[3])
    else:
        ms_total, ms_ex, ms_an = float(tot[0]), float(tot[1]), float(tot[2])
    if rank != 0:
        return None
    ms_cycle = ms_total / a.steps
    swept = 0
    per_var = {}
    for nm, st in cyc.log:
        if isinstance(st, str):
            continue
        key, hs, vs, _ = D.VARIABLES[nm]
        vnz = nz + 1 if vs == 1 else (1 if vs == -1 else nz)
        swept += nx * ny * vnz
        per_var[nm] = {"ms_analysis_rank0": st.ms_total, "units_rank0": int(st.units), "analysed_rank0": int(st.npts_analysed)}
    fma64 = eng.fma_peak(0)
    out = {"metric": "analysed grid points/s", "value": swept / (ms_cycle * 1e-3), "unit": "grid points/s",
           "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_cycle, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "cycle16: all 16 variables of input.nml:7 on the config-M grid %dx%dx%d, k=%d; "
                                  "host buffers -> scatter -> analysis (+ tune_q) -> gather -> host" % (nx, ny, nz, k),
                      "members": k, "groups": groups, "l2": "inputs larger than L2 (%.1f GB of fields per cycle)" % (h2d * world / 1e9),
                      "partition": "1 GPU" if world == 1 else "columns block-cyclic (nxb=%d) over a %dx%d process grid, "
                                   "members sliced over ranks before the scatter" % ((nxb,) + P.process_grid(world))},
           "seconds_per_cycle": ms_cycle * 1e-3, "points_swept_per_cycle": swept,
           "exchange": {"ms_per_cycle_rank_max": ms_ex, "share": ms_ex / ms_cycle,
                        "what": "scatter_grid + gather_grid (+ ensemble-mean height) on the device; NCCL send/recv when n_gpus > 1"},
           "analysis_ms_per_cycle_rank_max": ms_an,
           "e2e": {"value": swept / (ms_cycle * 1e-3), "unit": "grid points/s", "h2d_bytes_per_step": int(h2d * world),
                   "d2h_bytes_per_step": int(h2d * world), "note": "the timed region IS end to end: pinned host slabs in, pinned host slabs out"},
           "roofline": {"kernel": "whole cycle", "bound": "fp64", "achieved": None, "peak": fma64, "unit": "TFLOP/s",
                        "frac": None, "traffic": None, "note": "per-kernel rooflines are reported by the default workload (bench.py without --workload)"},
           "cpu_baseline": None, "gpu_launches": int(launches), "clocks": clocks, "per_variable": per_var,
           "ms_steps": [round(float(x), 1) for x in ms]}
    return out


def run_reference(a):
    """CPU arm of cycle16: the oracle on a bounded random sample of points for every variable group, all host
    threads; seconds per cycle is the extrapolation sum_g npts_g / rate_g (stated in `sample`)."""
    from cwbnwp_letkf_b200 import config as C
    from cwbnwp_letkf_b200 import driver as D
    from cwbnwp_letkf_b200 import synthetic as S
    from oracle import oracle as O
    k, nx, ny, nz = a.members, a.nx, a.ny, a.nz
    sc, rng = S.scenario_M(k=k, nx=nx, ny=ny, nz=nz)
    orc = O.Oracle(k, True)
    for o in sc.obs.values():
        orc.set_obs(o)
    threads = os.cpu_count() or 1
    groups = D.group_variables(list(C.VAR_UPDATE), C.sample_namelist)
    budget = max(1.0, min(6.0, 50.0 / (len(groups) * max(a.steps + a.warmup, 1))))
    vals = []
    detail = None
    for it in range(a.warmup + a.steps):
        sec, swept, detail = 0.0, 0, {}
        for group in groups:
            key, hs, vs, is_q = D.VARIABLES[group[0]]
            cfg = C.sample_namelist(group[0])
            vnz = nz + 1 if vs == 1 else (1 if vs == -1 else nz)
            npts = nx * ny * vnz
            n = 1500
            rate = None
            for _ in range(2):     # pilot, then sized run
                sel = np.sort(rng.choice(sc.npts, min(n, sc.npts), replace=False))
                xyz = np.ascontiguousarray(sc.xyz_grid[sel])
                f = np.stack([S.make_field(rng, k, xyz, 1.0, 1.0, 0.5) for _ in group], 0)
                t0 = time.perf_counter()
                orc.analyze(cfg, xyz, f, nthreads=threads)
                if is_q:
                    for g_ in range(len(group)):
                        O.tune_q(np.ascontiguousarray(f[g_]))
                dt = time.perf_counter() - t0
                rate = len(sel) / dt
                n = int(max(1500, min(sc.npts, rate * budget)))
            sec += npts / rate
            swept += npts * len(group)
            detail["+".join(group)] = {"points_per_s": rate, "sample": n}
        vals.append(swept / sec)
    v = float(np.mean(vals))
    swept = sum(nx * ny * (nz + 1 if D.VARIABLES[n_][2] == 1 else (1 if D.VARIABLES[n_][2] == -1 else nz)) for n_ in C.VAR_UPDATE)
    sample = ("per variable group a random sample of grid points (sized for ~%.0f s) analysed by the oracle (C++ port, "
              "OpenBLAS) on %d threads, full obs set; cycle time extrapolated as sum npts_g / rate_g" % (budget, threads))
    return {"impl": "reference", "metric": "analysed grid points/s", "value": v, "unit": "grid points/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * swept / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cycle16: all 16 variables of input.nml:7 on the config-M grid %dx%dx%d, k=%d" % (nx, ny, nz, k),
                       "members": k, "groups": groups},
            "seconds_per_cycle": swept / v, "cpu_baseline": {"value": v, "unit": "grid points/s", "cores": threads,
                                                             "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "grid points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "per_group": detail}
