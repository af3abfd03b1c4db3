#!/bin/bash
# Profiling pass run on the GPU box:  gpurun -- bash tools/profile.sh <tag>
# 1) plain run must exit 0, 2) ncu launch list (durations) of the SAME command, 3) ncu --set full of the dominant
# kernels: the tcgen05 GEMM (an encoder pointwise GEMM), the skinny recurrent GEMM, the depthwise-conv+LN kernel.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 40 -c 2 -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_skinny -s 30 -c 2 -o gpurun_out/${TAG}_skinny $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "skinny full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dwconv7 -s 20 -c 1 -o gpurun_out/${TAG}_dwconv $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
echo "dwconv full rc=$?"
ls -la gpurun_out/ | grep ${TAG}
