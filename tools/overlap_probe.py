#!/usr/bin/env python
"""Experiment: how much does overlapping the Gram (FP64 tensor) and eigen (FP64 FMA) stages of different
chunks buy?  Two library contexts analyse the two halves of the config-M grid from two host threads, so
that their kernels interleave on the device; compared with one context doing the whole grid."""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cwbnwp_letkf_b200 import config as C, host as H, synthetic as S  # noqa: E402


def main():
    nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 450
    sc, rng = S.scenario_M(k=32, nx=nx, ny=ny, nz=50)
    cfg = C.sample_namelist("T")
    cfg.tune_q = False
    dev = torch.device("cuda", 0)
    engs = [H.LetkfB200(32), H.LetkfB200(32)]
    for e in engs:
        for o in sc.obs.values():
            e.set_obs(o)
    xyz = torch.from_numpy(sc.xyz_grid).to(dev)
    f0 = torch.from_numpy(S.make_field(rng, 32, sc.xyz_grid, 280.0, 5.0, 1.0)).to(dev)
    npts = sc.npts
    half = (npts // 2 // nx) * nx

    def run_full():
        v = f0.clone()
        torch.cuda.synchronize()
        t = time.perf_counter()
        engs[0].analyze_dev(cfg, xyz, v)
        torch.cuda.synchronize()
        return time.perf_counter() - t, v

    def run_split():
        parts = [(0, half), (half, npts)]
        vs = [f0[:, a:b].contiguous() for a, b in parts]
        xs = [xyz[a:b].contiguous() for a, b in parts]
        torch.cuda.synchronize()
        t = time.perf_counter()
        th = [threading.Thread(target=engs[i].analyze_dev, args=(cfg, xs[i], vs[i])) for i in range(2)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t, torch.cat(vs, 1)

    for _ in range(2):
        run_full()
        run_split()
    tf, vf = run_full()
    ts, vs = run_split()
    same = (vf == vs).float().mean().item()
    print("full grid, one context: %.1f ms | two contexts on halves, concurrent: %.1f ms | ratio %.3f | identical values %.4f"
          % (tf * 1e3, ts * 1e3, tf / ts, same))


if __name__ == "__main__":
    main()
