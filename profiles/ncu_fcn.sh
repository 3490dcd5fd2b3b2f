#!/bin/bash
# ncu --set full of the matrix-function solve kernels.  Usage: profiles/ncu_fcn.sh <tag>
#   k = 256 on a 32x32x50 grid (51 200 units) and k = 32 on 150x150x50; each plain command must exit 0 first.
set -e
TAG=${1:-r02_fcn}
CMD1="python bench.py --members 256 --steps 1 --warmup 1 --nx 32 --ny 32 --nz 50 --no-cpu-baseline --no-e2e"
CMD2="python bench.py --steps 1 --warmup 1 --nx 150 --ny 150 --nz 50 --no-cpu-baseline --no-e2e"
$CMD1 > gpurun_out/${TAG}_k256_plain.json 2> gpurun_out/${TAG}_k256_plain.err
ncu --set full --clock-control none --import-source on -k regex:"fcn_blk_kernel|gram_tma_kernel" -s 2 -c 2 -o gpurun_out/${TAG}_k256 -f $CMD1 > gpurun_out/${TAG}_k256_ncu.log 2>&1
$CMD2 > gpurun_out/${TAG}_k32_plain.json 2> gpurun_out/${TAG}_k32_plain.err
ncu --set full --clock-control none --import-source on -k regex:"fcn32_kernel|search_kernel" -s 4 -c 3 -o gpurun_out/${TAG}_k32 -f $CMD2 > gpurun_out/${TAG}_k32_ncu.log 2>&1
