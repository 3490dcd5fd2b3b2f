#!/usr/bin/env python
"""Multi-GPU check of cycle.DeviceCycle: under torchrun every rank holds its members of the full grid, the cycle
runs with the NCCL scatter / gather exchanges, and the gathered result is compared on rank 0 with the
single-process letkf_driver mirror on the CPU oracle (tests/_driver_case.py).  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cwbnwp_letkf_b200 import cycle as CY  # noqa: E402
from cwbnwp_letkf_b200 import driver as D  # noqa: E402
from cwbnwp_letkf_b200 import host as H  # noqa: E402
from cwbnwp_letkf_b200 import partition as P  # noqa: E402
from _driver_case import KEYS_ALL, OracleBackend, VARS_ALL, copy_state, make_state, namelist  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sc, wrf, proj = make_state(k=8)
    eng = H.LetkfB200(sc.k, True, local)
    for o in sc.obs.values():
        eng.set_obs(o)
    lo, hi = P.member_slice(rank, world, sc.k)
    state = {key: torch.from_numpy(CY.to_member_major(wrf[key])[lo:hi].copy()).to(dev) for key in KEYS_ALL}
    geo = {g: wrf[g] for g in ("xlat", "xlon", "xlat_u", "xlon_u", "xlat_v", "xlon_v", "hgt")}
    cyc = CY.DeviceCycle(eng, namelist, proj, rank, world)
    cyc.run(state, geo, VARS_ALL)
    torch.cuda.synchronize()
    # collect every rank's members on rank 0
    ok, worst = True, 0.0
    full = {}
    for key in KEYS_ALL:
        t = P.allgather_members(state[key].contiguous(), sc.k, rank, world)
        full[key] = CY.from_member_major(t.cpu().numpy())
    if rank == 0:
        ref = copy_state(wrf)
        D.LetkfDriver(OracleBackend(sc), namelist, proj).run(ref, VARS_ALL)
        for key in KEYS_ALL:
            a, b = full[key], ref[key]
            same_nan = np.array_equal(np.isnan(a), np.isnan(b))
            good = ~np.isnan(b) & ~np.isnan(a)
            err = float(np.abs(a[good] - b[good]).max() / np.abs(b[good]).max())
            worst = max(worst, err)
            ok = ok and same_nan and err <= 5e-7
        print(json.dumps({"ok": bool(ok), "world": world, "max_rel": worst, "ms_exchange": cyc.ms_exchange,
                          "ms_analysis": cyc.ms_analysis, "variables": len(VARS_ALL)}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
