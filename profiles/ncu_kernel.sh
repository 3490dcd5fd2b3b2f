#!/bin/bash
# Full ncu capture of kernels matching a regex on the small bench case.  Usage: ncu_kernel.sh <regex> <tag> [skip] [count]
set -e
RE=$1; TAG=$2; SKIP=${3:-2}; CNT=${4:-1}
CMD="python bench.py --steps 1 --warmup 1 --nx 150 --ny 150 --nz 50 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/${TAG} -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
