#!/usr/bin/env python
"""Times letkf_scatter_grid / letkf_gather_grid (module_mpi_util.f90:190-358) as the NCCL exchange of
cwbnwp_letkf_b200.partition: member-major <-> column-major, one variable of the config-M grid.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_exchange.py

Prints one JSON line on rank 0: milliseconds per direction (max over ranks, CUDA events) and the bytes each
rank sends over NVLink.  SURVEY 8(f) rank 2 -- the step either side of the hot path, not part of bench.py."""
import argparse
import json
import os

import torch
import torch.distributed as dist

from cwbnwp_letkf_b200 import partition as P


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=32)
    ap.add_argument("--nx", type=int, default=450)
    ap.add_argument("--ny", type=int, default=450)
    ap.add_argument("--nz", type=int, default=50)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    k = a.members
    lo, hi = P.member_slice(rank, world, k)
    nxb = P.auto_block(a.nx, P.process_grid(world)[0])
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    field = torch.randn((hi - lo, a.nz, a.ny, a.nx), device=dev, generator=g)
    ref = field.clone()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ms_s, ms_g = [], []
    for it in range(a.reps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        var = P.scatter_grid(field, k, rank, world, nxb=nxb)
        ev[1].record()
        P.gather_grid(var, field, k, rank, world, nxb=nxb)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            ms_s.append(ev[0].elapsed_time(ev[1]))
            ms_g.append(ev[1].elapsed_time(ev[2]))
    ok = torch.equal(field, ref)                                   # scatter then gather is the identity
    t = torch.tensor([sum(ms_s) / len(ms_s), sum(ms_g) / len(ms_g), 0.0 if ok else 1.0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sent = field.numel() * 4 * (world - 1) / world
        print(json.dumps({"metric": "scatter_grid / gather_grid exchange", "n_gpus": world, "members": k,
                          "grid": [a.nx, a.ny, a.nz], "nxb": nxb, "ms_scatter": t[0].item(), "ms_gather": t[1].item(),
                          "bytes_sent_per_rank": sent, "GBs_per_rank_scatter": sent / t[0].item() / 1e6 if world > 1 else None,
                          "round_trip_identity": bool(t[2].item() == 0.0)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
