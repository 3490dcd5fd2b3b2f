#!/bin/bash
# ncu on a small 256-member case (24x24x50 grid, full obs set).  Usage: profiles/ncu_k256.sh <tag> [kernel regex]
# 1. the plain command must exit 0; 2. launch list (durations only); 3. --set full capture of the large-k kernels.
set -e
TAG=${1:-r01e_k256}
RE=${2:-eig_blk_kernel|gram_tma_kernel}
CMD="python bench.py --members 256 --steps 1 --warmup 1 --nx 24 --ny 24 --nz 50 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 1 -c 2 -o gpurun_out/${TAG} -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
