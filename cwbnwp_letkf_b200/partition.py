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


# ---- letkf_scatter_grid / letkf_gather_grid as an all-to-all (module_mpi_util.f90:190-358) ------------
# The reference keeps each member's full 3-D field on the rank that read it and transposes member-major
# <-> column-major with mpi_alltoallv around the analysis of every variable (mpi:262,325).  Here rank r
# holds members [lo_r, hi_r) of the full grid, `field[m, nz, ny, nx]`, and needs all k members of its own
# columns, `var[k, nz, loc_ny, loc_nx]` -- exactly the `[k, npts]` array letkf_b200_analyze takes.  One
# exchange per direction; NCCL over NVLink on GPUs, gloo in the CPU tests.
def _exchange(send, recv):
    import torch.distributed as dist
    ops = []
    for peer, t in recv.items():
        ops.append(dist.P2POp(dist.irecv, t, peer))
    for peer, t in send.items():
        ops.append(dist.P2POp(dist.isend, t, peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


_PLANS = {}


def _plan(device, world: int, nx: int, ny: int, nxb: int, nyb: int, stagger: int):
    """Index tensors of every rank's columns on `device`, built once per (decomposition, stagger): xi / yj for the
    two-step gather of scatter_grid and the flattened (y * nxx + x) indices for the index_copy_ of gather_grid.
    Rebuilding them per call (numpy tables + a host-to-device copy per peer, variable and direction) cost more
    than the exchange itself."""
    key = (str(device), world, nx, ny, nxb, nyb, stagger)
    p = _PLANS.get(key)
    if p is None:
        import torch
        xk, yk = ("xloc_u" if stagger == 1 else "xloc"), ("yloc_v" if stagger == 2 else "yloc")
        nxx = nx + (1 if stagger == 1 else 0)
        p = []
        for r in range(world):
            t = local_index_tables(r, world, nx, ny, nxb, nyb)
            xi = torch.as_tensor(t[xk], device=device)
            yj = torch.as_tensor(t[yk], device=device)
            p.append({"xi": xi, "yj": yj, "lx": len(t[xk]), "ly": len(t[yk]),
                      "flat": (yj[:, None] * nxx + xi[None, :]).reshape(-1)})
        _PLANS[key] = p
    return p


def scatter_grid(field, k: int, rank: int, world: int, nxb: int = 1, nyb: int = 1, stagger: int = 0):
    """field: torch tensor [m_r, nz, ny(+1), nx(+1)] (this rank's members of the full grid, x fastest).
    Returns var [k, nz, loc_ny, loc_nx] with every member of the columns this rank owns.  stagger: 0 mass,
    1 U (x extent nx+1), 2 V (y extent ny+1), as in letkf_scatter_grid."""
    import torch
    m, nz, nyy, nxx = field.shape
    nx, ny = nxx - (1 if stagger == 1 else 0), nyy - (1 if stagger == 2 else 0)
    plan = _plan(field.device, world, nx, ny, nxb, nyb, stagger)
    mine = plan[rank]
    var = torch.empty((k, nz, mine["ly"], mine["lx"]), dtype=field.dtype, device=field.device)
    flat = field.reshape(m, nz, nyy * nxx)
    send, recv = {}, {}
    for r in range(world):
        # one gather per peer: [m_me, nz, ly_r * lx_r] in the peer's local column order
        part = flat.index_select(2, plan[r]["flat"]).view(m, nz, plan[r]["ly"], plan[r]["lx"])
        lo, hi = member_slice(r, world, k)
        if r == rank:
            var[lo:hi] = part
        else:
            send[r] = part
            recv[r] = var[lo:hi]                                                  # contiguous slab of members
    if world > 1:
        _exchange(send, recv)
    return var


def gather_grid(var, field, k: int, rank: int, world: int, nxb: int = 1, nyb: int = 1, stagger: int = 0):
    """Inverse of scatter_grid: writes the analysed columns of every rank back into this rank's members of
    the full grid (in place; columns nobody analysed -- the last staggered column / row -- keep their values)."""
    import torch
    nz, nyy, nxx = field.shape[1:]
    nx, ny = nxx - (1 if stagger == 1 else 0), nyy - (1 if stagger == 2 else 0)
    plan = _plan(field.device, world, nx, ny, nxb, nyb, stagger)
    lo, hi = member_slice(rank, world, k)
    send, recv = {}, {}
    for r in range(world):
        if r == rank:
            recv[r] = var[lo:hi]
        else:
            send[r] = var[slice(*member_slice(r, world, k))].contiguous()
            recv[r] = torch.empty((hi - lo, nz, plan[r]["ly"], plan[r]["lx"]), dtype=var.dtype, device=var.device)
    if world > 1:
        _exchange(send, {r: t for r, t in recv.items() if r != rank})
    out = field.view(hi - lo, nz, nyy * nxx)
    for r in range(world):
        out.index_copy_(2, plan[r]["flat"], recv[r].reshape(hi - lo, nz, -1))
    return field
