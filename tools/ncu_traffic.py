#!/usr/bin/env python
"""Turns the captures of profiles/regen.sh into committed evidence (runs here, no GPU):
  profiles/<tag>_ncu_<case>_summary.json : the counters of every captured launch (tools/ncu_summary.py)
  profiles/ncu_traffic.json              : DRAM bytes per unit of each stage's kernel, stamped with the tree
Usage: python tools/ncu_traffic.py r02"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
GB = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
STAGE = {"search_kernel": "search", "gram32_dmma_kernel": "gram", "gram_tma_kernel": "gram", "fcn32_kernel": "solve",
         "fcn_blk_kernel": "solve", "count_rows": "count_rows"}


def main():
    tag = sys.argv[1]
    stamp = open(os.path.join(OUT, tag + "_tree_stamp.txt")).read().strip()
    traffic = {"source": "profiles/regen.sh %s (ncu --set full --clock-control none, small cases)" % tag, "git": stamp,
               "stamp": stamp}
    for case in ("k32", "k256", "k96"):
        rep = os.path.join(OUT, "%s_%s_raw.csv" % (tag, case))
        if not os.path.exists(rep):
            rep = os.path.join(OUT, "%s_%s.ncu-rep" % (tag, case))
        if not os.path.exists(rep):
            continue
        summ = os.path.join(PROF, "%s_ncu_%s_summary.json" % (tag, case))
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, summ],
                       stdout=subprocess.DEVNULL, check=True)
        launches = json.load(open(summ))
        plain = json.load(open(os.path.join(OUT, "%s_%s_plain.json" % (tag, case))))
        k = plain["config"]["members"]
        # one captured launch = one pipeline chunk: at most 2^18 grid points (the k = 32 case has five chunks)
        npts = min(int(round(plain["points_analysed"] / max(plain["analysed_fraction"], 1e-9))), 1 << 18)
        units = min(plain["points_analysed"], int(round(npts * plain["analysed_fraction"])))
        rows = plain["rows_per_analysed_point"]
        alg = {"gram": (4 * k + 8) * rows + 8 * k * k + 8 * k,           # gathered rows + C, b written
               "solve": 8 * k * k + 8 * k + 8 * k,                      # C, b read + xb in / xa out (one field)
               "search": None, "count_rows": None}
        entry = {}
        for L in launches:
            stage = next((v for kk, v in STAGE.items() if kk in L["kernel"]), None)
            if stage is None:
                continue
            byt = L["dram_read"] * GB[L["dram_read_unit"]] + L["dram_write"] * GB[L["dram_write_unit"]]
            per = units if stage in ("gram", "solve") else npts
            e = {"kernel": L["kernel"], "dram_bytes_per_launch": byt, "units_per_launch": per, "bytes_per_unit": byt / per,
                 "algorithmic_bytes_per_unit": alg[stage], "duration_ms": L["duration"] * (1.0 if L["duration_unit"] == "ms" else 1e-3),
                 "fp64_pipe_pct": L.get("fp64_pipe_pct"), "tensor_cycles_pct": L.get("tensor_cycles_pct"),
                 "issue_active_pct": L.get("issue_active_pct"), "warps_active_pct": L.get("warps_active_pct"),
                 "l2_hit_pct": L.get("l2_hit_pct"), "regs": L.get("regs")}
            if alg[stage]:
                e["traffic_over_algorithmic"] = e["bytes_per_unit"] / alg[stage]
            # keep the largest launch of a stage (the search runs once per tree)
            if stage not in entry or byt > entry[stage]["dram_bytes_per_launch"]:
                entry[stage] = e
        traffic["k%d" % k] = entry
    with open(os.path.join(PROF, "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps({kk: {s: round(v["bytes_per_unit"]) for s, v in vv.items()} for kk, vv in traffic.items()
                      if isinstance(vv, dict)}))


if __name__ == "__main__":
    main()
