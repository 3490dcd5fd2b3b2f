#!/bin/bash
# ncu --set full of fcn_blk_kernel on small grids (every replay pass saves / restores the C matrices the
# kernel overwrites: keep them small).  Usage: profiles/ncu_fcn_blk.sh <tag>
set -e
TAG=${1:-r02_fcnblk}
CMD1="python bench.py --members 256 --steps 1 --warmup 1 --nx 12 --ny 12 --nz 20 --no-cpu-baseline --no-e2e"
CMD2="python bench.py --members 96 --steps 1 --warmup 1 --nx 32 --ny 32 --nz 20 --no-cpu-baseline --no-e2e"
$CMD1 > gpurun_out/${TAG}_k256_plain.json 2> gpurun_out/${TAG}_k256_plain.err
ncu --set full --clock-control none --import-source on -k regex:"fcn_blk_kernel" -s 1 -c 1 -o gpurun_out/${TAG}_k256 -f $CMD1 > gpurun_out/${TAG}_k256_ncu.log 2>&1
$CMD2 > gpurun_out/${TAG}_k96_plain.json 2> gpurun_out/${TAG}_k96_plain.err
ncu --set full --clock-control none --import-source on -k regex:"fcn_blk_kernel|gram_tma_kernel" -s 2 -c 2 -o gpurun_out/${TAG}_k96 -f $CMD2 > gpurun_out/${TAG}_k96_ncu.log 2>&1
