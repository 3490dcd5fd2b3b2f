#!/bin/bash
# Regenerates the ncu evidence of a round on a GPU box.  Usage (from the repo root, under gpurun):
#     bash profiles/regen.sh r02
# For each of three small cases it runs the plain command (must exit 0), then -- only then -- the launch list and
# one `ncu --set full` capture of the hot kernels (recipe: /opt/skills/guides/B200_PROFILING.md).  The captures
# carry the stamp of the kernel sources (tools/tree_stamp.py); `python tools/ncu_traffic.py r02` afterwards (no
# GPU needed) turns the .ncu-rep files into profiles/ncu_traffic.json + per-kernel summaries, and bench.py only
# reports capture-derived numbers whose stamp equals the tree it runs from.
# Small grids on purpose: every ncu replay pass saves / restores the buffers a kernel overwrites.
set -e
TAG=${1:-r02}
OUT=gpurun_out
python tools/tree_stamp.py > $OUT/${TAG}_tree_stamp.txt
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
C32="$B --nx 150 --ny 150 --nz 50"
C256="$B --members 256 --nx 12 --ny 12 --nz 20"
C96="$B --members 96 --nx 32 --ny 32 --nz 20"
# ---- k = 32 (config M at 1/9 of the columns: one pipeline chunk = 2^18 units)
$C32 > $OUT/${TAG}_k32_plain.json 2> $OUT/${TAG}_k32_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_k32_launches.csv $C32 > $OUT/${TAG}_k32_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"search_kernel|gram32_dmma_kernel|fcn32_kernel|count_rows" -s 8 -c 8 -o $OUT/${TAG}_k32 -f $C32 > $OUT/${TAG}_k32_ncu.log 2>&1
# ---- k = 256 (config L)
$C256 > $OUT/${TAG}_k256_plain.json 2> $OUT/${TAG}_k256_plain.err
ncu --set full --clock-control none --import-source on -k regex:"gram_tma_kernel|fcn_blk_kernel" -s 2 -c 2 -o $OUT/${TAG}_k256 -f $C256 > $OUT/${TAG}_k256_ncu.log 2>&1
# ---- k = 96 (the shipped nmember)
$C96 > $OUT/${TAG}_k96_plain.json 2> $OUT/${TAG}_k96_plain.err
ncu --set full --clock-control none --import-source on -k regex:"gram_tma_kernel|fcn_blk_kernel" -s 2 -c 2 -o $OUT/${TAG}_k96 -f $C96 > $OUT/${TAG}_k96_ncu.log 2>&1
# gpurun brings back at most 64 MiB: export the raw counter pages (and the per-instruction SASS pages of the two
# solve kernels) as CSV and drop the .ncu-rep files
for c in k32 k256 k96; do
  ncu -i $OUT/${TAG}_$c.ncu-rep --page raw --csv > $OUT/${TAG}_${c}_raw.csv 2>/dev/null
done
ncu -i $OUT/${TAG}_k32.ncu-rep --page source --csv --print-source sass -k regex:fcn32_kernel > $OUT/${TAG}_k32_fcn32_sass.csv 2>/dev/null || true
ncu -i $OUT/${TAG}_k256.ncu-rep --page source --csv --print-source sass -k regex:fcn_blk_kernel > $OUT/${TAG}_k256_fcn_blk_sass.csv 2>/dev/null || true
rm -f $OUT/${TAG}_*.ncu-rep
gzip -f $OUT/${TAG}_*_sass.csv
ls -la $OUT/${TAG}_*
