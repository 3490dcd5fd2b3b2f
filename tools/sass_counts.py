#!/usr/bin/env python
"""Counts of the SASS mnemonics that identify the hardware paths a kernel uses (B200_PROFILING.md table), per
kernel of libletkf_b200.so: DMMA (mma.sync.m8n8k4.f64, the FP64 tensor pipe), DFMA, UBLKCP / SYNCS (TMA bulk
copies + mbarriers), SHFL, LDS, LDG, BAR; and UTC*MMA / LDTM / STTM / UTMALDG (tcgen05 / TMEM / tensor-map TMA),
which this FP64 path cannot use (tcgen05.mma has no f64 kind).  Usage: tools/sass_counts.py [out.json]"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from tree_stamp import tree_stamp  # noqa: E402

LIB = os.path.join(ROOT, "cwbnwp_letkf_b200", "_build", "libletkf_b200.so")
OPS = ["DMMA", "DFMA", "DADD", "DMUL", "FFMA", "HMMA", "UBLKCP", "SYNCS", "SHFL", "LDS", "STS", "LDG", "STG", "BAR",
       "MUFU", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("lk::", "")
            name = re.sub(r"\(.*", "", name)
            cur = out.setdefault(name, {"instructions": 0})
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur["instructions"] += 1
            op = m.group(1).split(".")[0]
            if op in OPS:
                cur[op] = cur.get(op, 0) + 1
    res = {"library": os.path.relpath(LIB, ROOT), "arch": arch, "tree": tree_stamp(), "kernels": out,
           "total": {op: sum(k.get(op, 0) for k in out.values()) for op in OPS}}
    js = json.dumps(res, indent=1, sort_keys=True)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(js + "\n")
    print(json.dumps(res["total"]), arch)


if __name__ == "__main__":
    main()
