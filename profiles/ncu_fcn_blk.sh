#!/bin/bash
# ncu --set full of fcn_blk_kernel on a small grid (every replay pass saves / restores the C matrices the
# kernel overwrites: keep them small).  Usage: profiles/ncu_fcn_blk.sh <tag> [members] [nx]
set -e
TAG=${1:-r02_fcnblk}
K=${2:-256}
NX=${3:-12}
CMD1="python bench.py --members $K --steps 1 --warmup 1 --nx $NX --ny $NX --nz 20 --no-cpu-baseline --no-e2e --no-secondary"
$CMD1 > gpurun_out/${TAG}_k${K}_plain.json 2> gpurun_out/${TAG}_k${K}_plain.err
ncu --set full --clock-control none --import-source on -k regex:"fcn_blk_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_k${K} -f $CMD1 > gpurun_out/${TAG}_k${K}_ncu.log 2>&1
