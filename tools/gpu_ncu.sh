#!/bin/bash
# tools/gpu_ncu.sh TAG REGEX [skip] [count] [particles]: launch list + one ncu --set full capture of the kernels matching REGEX
set -u
mkdir -p gpurun_out
TAG=$1; RE=$2; SKIP=${3:-15}; CNT=${4:-6}; P=${5:-1.0e7}
CMD="timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --state lattice --particles $P"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
