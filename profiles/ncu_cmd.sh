#!/bin/bash
# ncu recipe for this repo (run under gpurun on one B200).  Usage: profiles/ncu_cmd.sh <tag>
# 1. plain run must exit 0; 2. launch list (durations only); 3. --set full capture of the top kernels.
set -e
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1 --nx 150 --ny 150 --nz 50 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'eig32_chain_kernel|gram32_dmma_kernel|search_kernel|transform32_kernel' -s 12 -c 8 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
