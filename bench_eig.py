#!/usr/bin/env python
"""BASELINE config 5 (SURVEY.md 8(d) "E"): batched symmetric k x k eigensolves/s for
k = 32/64/128/256, FP32 and FP64, against LAPACK ?syevd on the host cores.

Matrices are LETKF-shaped, A_b = mu I + Y_b Y_b^T (mu = (k-1)/1.1, Y_b in R^{k x 300}, columns with
the member mean removed and scaled like exp(-r2/4)/err), generated on the device in chunks (10^6
matrices of k = 256 do not fit in HBM at once: SURVEY H7); `--total` matrices are solved per
(k, dtype), a sample is re-solved with LAPACK and compared through eigenvalues and residuals.
One JSON line per (k, dtype)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def gen(torch, b, k, p, dtype, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Y = torch.randn((b, k, p), generator=g, device=dev, dtype=torch.float64)
    Y -= Y.mean(1, keepdim=True)
    s = torch.exp(-0.25 * 13.33 * torch.rand((b, 1, p), generator=g, device=dev, dtype=torch.float64)) / \
        (0.5 + 2.0 * torch.rand((b, 1, p), generator=g, device=dev, dtype=torch.float64))
    Y *= s
    A = torch.bmm(Y, Y.transpose(1, 2))
    A += (k - 1) / 1.1 * torch.eye(k, device=dev, dtype=torch.float64)
    return A.to(dtype).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ks", default="32,64,128,256")
    ap.add_argument("--total", type=int, default=0, help="matrices per case (0: sized for ~2 s)")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--sample", type=int, default=256)
    ap.add_argument("--cpu-sample", type=int, default=4096)
    a = ap.parse_args()
    import torch
    from cwbnwp_letkf_b200 import host as H
    from oracle import oracle as O
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cores = os.cpu_count() or 1
    for dt in a.dtypes.split(","):
        tdt = torch.float64 if dt == "f64" else torch.float32
        ndt = np.float64 if dt == "f64" else np.float32
        for k in [int(x) for x in a.ks.split(",")]:
            eng = H.LetkfB200(k, dt == "f64", 0)
            fma = eng.fma_peak(0 if dt == "f64" else 1)
            stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
            chunk = max(64, min(1 << 17, int(2e9 / (k * k * (8 if dt == "f64" else 4)) / 3)))
            A = gen(torch, chunk, k, 300, tdt, dev, 20261018)
            W = torch.empty((chunk, k), dtype=tdt, device=dev)
            V = torch.empty_like(A)
            torch.cuda.synchronize()
            # warm-up + rate estimate
            t0 = time.perf_counter()
            sweeps = eng.syevd_batched_dev(A, W, V)
            torch.cuda.synchronize()
            est = time.perf_counter() - t0
            nrep = max(1, int(2.0 / max(est, 1e-4))) if a.total == 0 else max(1, a.total // chunk)
            nrep = min(nrep, 200)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(nrep):
                eng.syevd_batched_dev(A, W, V)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            rate = nrep * chunk / (ms * 1e-3)
            # verification on a sample vs LAPACK
            ns = min(a.sample, chunk)
            An = A[:ns].cpu().numpy()
            Wn, Vn = W[:ns].cpu().numpy(), V[:ns].cpu().numpy()
            Wl, Vl = O.syevd_batch(An, nthreads=cores)
            eps = np.finfo(ndt).eps
            nrm = np.abs(An.astype(np.float64)).sum(2).max(1)
            ev_err = float((np.abs(Wn.astype(np.float64) - Wl.astype(np.float64)).max(1) / nrm).max())
            v = Vn.astype(np.float64).transpose(0, 2, 1)
            res = np.abs(An.astype(np.float64) @ v - v * Wn.astype(np.float64)[:, None, :]).max((1, 2)) / nrm
            orth = np.abs(v.transpose(0, 2, 1) @ v - np.eye(k)).max((1, 2))
            # CPU baseline: LAPACK over a sample, all cores
            nc = min(a.cpu_sample, chunk) if k <= 64 else min(a.cpu_sample // (k // 32) ** 2, chunk)
            Ac = A[:nc].cpu().numpy()
            t0 = time.perf_counter()
            O.syevd_batch(Ac, nthreads=cores)
            tc = time.perf_counter() - t0
            out = {"metric": "batched kxk symmetric eigensolves/s", "k": k, "dtype": dt, "value": rate,
                   "unit": "eigensolves/s", "batch": chunk, "reps": nrep, "ms": ms, "max_sweeps": sweeps,
                   "model_tflops_4k3": 4.0 * k ** 3 * rate / 1e12, "fma_peak_tflops": fma,
                   "frac_of_fma_peak_4k3_model": 4.0 * k ** 3 * rate / 1e12 / fma,
                   "check": {"sample": ns, "eigenvalue_err_over_norm": ev_err, "residual_over_norm": float(res.max()),
                             "orthogonality": float(orth.max()), "eps": float(eps)},
                   "cpu_baseline": {"value": nc / tc, "unit": "eigensolves/s", "cores": cores,
                                    "kind": "LAPACK ?syevd (OpenBLAS 0.3.31.dev via scipy)", "sample": nc},
                   "matrices": "mu*I + Y Y^T, Y k x 300 (LETKF-shaped), generated on device"}
            print(json.dumps(out), flush=True)
            eng.finalize()
            del A, W, V
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
