"""GPU tier: one full analysis cycle (all 16 variables of input.nml:7) with the fields resident in HBM
(cwbnwp_letkf_b200.cycle.DeviceCycle: device-side letkf_scatter_grid / letkf_gather_grid, ensemble-mean height on
the device, the eight hydrometeor variables in ONE library pass) against the letkf_driver mirror running on the
CPU oracle (module_letkf_core.f90:21-297, module_mpi_util.f90:190-358,445-580)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from cwbnwp_letkf_b200 import cycle as CY
from cwbnwp_letkf_b200 import driver as D
from cwbnwp_letkf_b200 import host as H

from _driver_case import KEYS_ALL, OracleBackend, VARS_ALL, copy_state, make_state, namelist

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEO = ("xlat", "xlon", "xlat_u", "xlon_u", "xlat_v", "xlon_v", "hgt")


def _compare(got, ref, wrf):
    for key in KEYS_ALL:
        a, b = got[key], ref[key]
        assert np.array_equal(np.isnan(a), np.isnan(b)), key
        ok = ~np.isnan(b)
        scale = np.abs(b[ok]).max()
        assert np.abs(a[ok] - b[ok]).max() <= 5e-7 * scale, key
        untouched = (b == wrf[key]) | np.isnan(b)
        assert np.array_equal(a[untouched & ok], wrf[key][untouched & ok]), key


def test_device_cycle_matches_the_host_driver_on_the_oracle():
    import torch
    sc, wrf, proj = make_state()
    ref = copy_state(wrf)
    D.LetkfDriver(OracleBackend(sc), namelist, proj).run(ref, VARS_ALL)
    eng = H.LetkfB200(sc.k)
    for o in sc.obs.values():
        eng.set_obs(o)
    dev = torch.device("cuda", 0)
    state = {key: torch.from_numpy(CY.to_member_major(wrf[key])).to(dev) for key in KEYS_ALL}
    geo = {g: wrf[g] for g in GEO}
    cyc = CY.DeviceCycle(eng, namelist, proj)
    log = cyc.run(state, geo, VARS_ALL)
    assert [n for n, _ in log] == VARS_ALL
    got = {key: CY.from_member_major(state[key].cpu().numpy()) for key in KEYS_ALL}
    _compare(got, ref, wrf)
    # the eight hydrometeor variables went through one pass: they report the same stats object
    st = [s for n, s in log if n in ("QRAIN", "QNHAIL")]
    assert st[0] is st[1]
    # ... and single-variable passes give the same bits
    state2 = {key: torch.from_numpy(CY.to_member_major(wrf[key])).to(dev) for key in KEYS_ALL}
    CY.DeviceCycle(eng, namelist, proj, batch=False).run(state2, geo, VARS_ALL)
    for key in KEYS_ALL:
        a, b = state[key].cpu().numpy(), state2[key].cpu().numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), key
    eng.finalize()


def test_device_cycle_two_ranks_equal_one():
    """letkf_scatter_grid / letkf_gather_grid over NCCL: needs two GPUs on the box (skipped otherwise; the
    exchange itself is covered on the CPU by tests/test_partition.py with gloo)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "cycle_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["world"] == 2, out
