"""Multi-GPU host logic: the reference's domain decomposition and observation replication.

* Grid columns are dealt out cyclically (block size 1), as ``letkf_local_info`` does with its
  ``nproc_x x nproc_y`` process grid (module_mpi_util.f90:10-11,80-127): rank (id_x, id_y) owns global
  columns ``x = id_x + 1 + m*nproc_x``, ``y = id_y + 1 + n*nproc_y``.  All nz levels of a column stay
  on its owner; cyclic ownership is the reference's load-balancing device (observation density is
  clustered).
* Observations are fully replicated.  The reference has each of its last k ranks read ONE member's
  H(x) and assembles ``hdxb`` everywhere with an in-place ``mpi_iallgatherv``
  (module_gts_omboma.f90:582-605, module_radar.f90:160-180).  Here every rank contributes a contiguous
  slice of members and one all-gather (NCCL on GPUs, gloo in the CPU tests) rebuilds the array.

No collective is needed inside the per-variable hot path.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def process_grid(world: int) -> Tuple[int, int]:
    """mpi_dims_create(nproc, 2) (module_mpi_util.f90:50): the most square factorisation, larger
    factor first."""
    best = (world, 1)
    for a in range(1, int(world ** 0.5) + 1):
        if world % a == 0:
            best = (world // a, a)
    return best


def local_columns(rank: int, world: int, nx: int, ny: int, nxb: int = 1, nyb: int = 1) -> np.ndarray:
    """Flattened (i + nx*j) indices of the columns rank owns, i fastest (module_mpi_util.f90:80-127).
    nxb / nyb are the reference's block sizes (compile-time parameters, 1 as shipped, mpi:10-11).  Results
    do not depend on them; a block of 16 along x keeps the eigensolver's warm-start runs on truly adjacent
    columns when the grid is split along x."""
    t = local_index_tables(rank, world, nx, ny, nxb, nyb)
    return (t["xloc"][None, :] + nx * t["yloc"][:, None]).reshape(-1)


def local_points(rank: int, world: int, nx: int, ny: int, nz: int, nxb: int = 1, nyb: int = 1) -> np.ndarray:
    """Global point indices (i + nx*(j + ny*l)) of the rank's slab, local order i, j, then level --
    the memory order of var(loc_nx, loc_ny, nz, :) (module_letkf_core.f90:85)."""
    cols = local_columns(rank, world, nx, ny, nxb, nyb)
    return (cols[None, :] + nx * ny * np.arange(nz)[:, None]).reshape(-1)


def member_slice(rank: int, world: int, k: int) -> Tuple[int, int]:
    """Members whose H(x) this rank 'reads' (contiguous block; k need not divide evenly)."""
    base, rem = divmod(k, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_members(local, k: int, rank: int, world: int):
    """All-gather member slices ``local[(hi-lo), n, nvar]`` into the full ``[k, n, nvar]`` tensor on
    every rank (torch.distributed must be initialised; works for NCCL/CUDA and gloo/CPU tensors)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    shape = (k,) + tuple(local.shape[1:])
    full = torch.empty(shape, dtype=local.dtype, device=local.device)
    if k % world == 0:
        dist.all_gather_into_tensor(full, local.contiguous())
        return full
    parts = []
    for r in range(world):
        lo, hi = member_slice(r, world, k)
        parts.append(full[lo:hi])
    # uneven split: all_gather with per-rank views (gloo and nccl both accept unequal sizes here only
    # through a list of equal-shaped tensors, so pad to the largest slice)
    m = max(p.shape[0] for p in parts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    for r in range(world):
        parts[r].copy_(bufs[r][: parts[r].shape[0]])
    return full


def _block_cyclic(start: int, n: int, nproc: int, nb: int) -> np.ndarray:
    """0-based indices 'do i = start, n, stride; do j = i, min(n, i+nb-1)' (module_mpi_util.f90:84-103)
    with start = id*nb + 1, stride = nproc*nb."""
    out = []
    i = start * nb + 1
    while i <= n:
        out.extend(range(i, min(n, i + nb - 1) + 1))
        i += nproc * nb
    return np.asarray(out, np.int64) - 1


def local_index_tables(rank: int, world: int, nx: int, ny: int, nxb: int = 1, nyb: int = 1) -> dict:
    """The four index tables of letkf_local_info (module_mpi_util.f90:73-172), 0-based: mass columns
    ``xloc``/``yloc`` and the staggered ``xloc_u`` (over nx+1) / ``yloc_v`` (over ny+1).  Block-cyclic
    with block sizes nxb, nyb (1 in the shipped namelist)."""
    npx, npy = process_grid(world)
    idx, idy = rank % npx, rank // npx
    return {"xloc": _block_cyclic(idx, nx, npx, nxb), "yloc": _block_cyclic(idy, ny, npy, nyb),
            "xloc_u": _block_cyclic(idx, nx + 1, npx, nxb), "yloc_v": _block_cyclic(idy, ny + 1, npy, nyb)}


def auto_block(n: int, nproc: int, prefer: int = 16) -> int:
    """Largest block size <= prefer (powers of two) whose block-cyclic split of n columns over nproc ranks
    leaves the fullest rank within 2 % of the cyclic (block 1) split -- long blocks keep neighbouring
    columns on one rank (warm starts), but a short grid must not lose its balance to them."""
    def fullest(b):
        return max(len(_block_cyclic(i, n, nproc, b)) for i in range(nproc))
    base = fullest(1)
    b = prefer
    while b > 1:
        if fullest(b) <= base * 1.02:
            return b
        b //= 2
    return 1
