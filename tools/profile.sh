#!/bin/bash
# Profiling pass run on the GPU box:  gpurun -- bash tools/profile.sh <tag>
# 1) plain run must exit 0, 2) ncu launch list (durations), 3) ncu --set full of the dominant kernels.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 20 -c 2 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dwconv7 -s 8 -c 1 -o gpurun_out/${TAG}_dwconv $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "dwconv full rc=$?"
ls -la gpurun_out/
