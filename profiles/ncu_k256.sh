#!/bin/bash
# Full ncu capture of the large-k kernels on a small 256-member case (24x24x50 grid, full obs set).
# Usage: profiles/ncu_k256.sh <tag>.  The plain command must exit 0 before the capture.
set -e
TAG=${1:-r01e_k256}
CMD="python bench.py --members 256 --steps 1 --warmup 1 --nx 24 --ny 24 --nz 50 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --set full --clock-control none --import-source on -k regex:'eig_blk_kernel|gram_tma_kernel' -s 2 -c 2 -o gpurun_out/${TAG} -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
