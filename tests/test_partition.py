"""CPU tier, world_size = 2 over gloo: the multi-GPU host logic -- cyclic column ownership
(module_mpi_util.f90:80-127) and the member-sliced all-gather of H(x) that replicates the
observation-space ensemble on every rank (module_gts_omboma.f90:601-605, module_radar.f90:179)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cwbnwp_letkf_b200 import partition as P


def test_process_grid_and_ownership_cover_the_domain():
    assert P.process_grid(1) == (1, 1) and P.process_grid(2) == (2, 1)
    assert P.process_grid(4) == (2, 2) and P.process_grid(8) == (4, 2)
    nx, ny, nz = 13, 7, 3
    for world in (1, 2, 4, 8):
        seen = np.concatenate([P.local_points(r, world, nx, ny, nz) for r in range(world)])
        assert len(seen) == nx * ny * nz and len(np.unique(seen)) == nx * ny * nz
        sizes = [len(P.local_columns(r, world, nx, ny)) for r in range(world)]
        assert max(sizes) - min(sizes) <= max(nx, ny)            # cyclic deal is balanced
    # reference rule: rank (id_x, id_y) owns x = id_x + m*nproc_x, y = id_y + n*nproc_y
    cols = P.local_columns(3, 4, nx, ny)                          # (id_x, id_y) = (1, 1)
    assert all((c % nx) % 2 == 1 and (c // nx) % 2 == 1 for c in cols)
    # levels of a column stay together
    pts = P.local_points(1, 2, nx, ny, nz)
    ncol = len(P.local_columns(1, 2, nx, ny))
    assert np.array_equal(pts[:ncol] + nx * ny, pts[ncol:2 * ncol])


def test_member_slices_partition_the_ensemble():
    for k, world in ((32, 2), (32, 8), (96, 8), (10, 4), (7, 2)):
        sl = [P.member_slice(r, world, k) for r in range(world)]
        assert sl[0][0] == 0 and sl[-1][1] == k
        assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))


def _worker(rank, world, port, k, n, nvar, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(k * n * nvar, dtype=torch.float32).reshape(k, n, nvar)
    lo, hi = P.member_slice(rank, world, k)
    got = P.allgather_members(full[lo:hi].clone(), k, rank, world)
    ok = torch.equal(got, full)
    # every rank analyses only its own columns; together they cover the grid once
    mine = torch.from_numpy(P.local_points(rank, world, 6, 5, 2))
    cnt = torch.zeros(60, dtype=torch.int32)
    cnt[mine] = 1
    dist.all_reduce(cnt)
    ok = ok and bool((cnt == 1).all())
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [8, 7])
def test_gloo_world2_allgather_members(k):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, 11, 3, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world


def test_auto_block_keeps_ranks_balanced():
    from cwbnwp_letkf_b200 import partition as P
    for n, npx in [(96, 4), (450, 4), (450, 2), (64, 4), (100, 4), (150, 2), (7, 4)]:
        b = P.auto_block(n, npx)
        sizes = [len(P._block_cyclic(i, n, npx, b)) for i in range(npx)]
        cyc = [len(P._block_cyclic(i, n, npx, 1)) for i in range(npx)]
        assert sum(sizes) == n and max(sizes) <= max(cyc) * 1.02 and b in (1, 2, 4, 8, 16)
    assert P.auto_block(450, 4) == 16 and P.auto_block(96, 4) == 8


def _a2a_worker(rank, world, port, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    nx, ny, nz = 9, 6, 3
    for stagger, nxb in ((0, 1), (1, 2), (2, 1)):
        nxx, nyy = nx + (stagger == 1), ny + (stagger == 2)
        g = torch.Generator().manual_seed(5)
        full = torch.randn((k, nz, nyy, nxx), generator=g)                    # the same on every rank
        lo, hi = P.member_slice(rank, world, k)
        var = P.scatter_grid(full[lo:hi].clone(), k, rank, world, nxb=nxb, stagger=stagger)
        t = P.local_index_tables(rank, world, nx, ny, nxb, 1)
        xi = torch.as_tensor(t["xloc_u" if stagger == 1 else "xloc"])
        yj = torch.as_tensor(t["yloc_v" if stagger == 2 else "yloc"])
        ok = ok and torch.equal(var, full.index_select(3, xi).index_select(2, yj))
        # "analysis": every rank adds its rank + 1 to its columns; the gathered grid must show the owner map
        mine = full[lo:hi].clone()
        P.gather_grid(var + (rank + 1), mine, k, rank, world, nxb=nxb, stagger=stagger)
        owner = torch.zeros((nyy, nxx))
        for r in range(world):
            tr = P.local_index_tables(r, world, nx, ny, nxb, 1)
            xr = torch.as_tensor(tr["xloc_u" if stagger == 1 else "xloc"])
            yr = torch.as_tensor(tr["yloc_v" if stagger == 2 else "yloc"])
            owner[yr[:, None], xr[None, :]] = r + 1
        ok = ok and torch.allclose(mine - full[lo:hi], owner.expand(hi - lo, nz, nyy, nxx))
        ok = ok and bool((owner > 0).all())
    out[rank] = int(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [8, 5])
def test_gloo_world2_scatter_gather_grid(k):
    """letkf_scatter_grid / letkf_gather_grid (module_mpi_util.f90:190-358) as one exchange per direction."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_a2a_worker, args=(2, port, k, out), nprocs=2, join=True)
    assert out[0] == 1 and out[1] == 1
