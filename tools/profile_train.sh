set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/r01u_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r01u_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r01u_launches.csv $CMD > gpurun_out/r01u_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_cluster -s 30 -c 1 -o gpurun_out/r01u_attn $CMD > gpurun_out/r01u_ncu2.log 2>&1
echo "attn full rc=$?"
