# Round-2 ncu evidence (run under gpurun, one GPU): per-kernel metrics for EVERY kernel of one eager LSTM train step, one
# Transformer train step and a few beam-search steps, plus `--set full` captures of the top kernels.
# Summaries: python tools/ncu_summary.py metrics <csv> / full <ncu-rep>  ->  profiles/r02_*.txt
# usage: bash tools/profile_r02.sh [tag]      (tag: file-name prefix, default r02)
set -u
T=${1:-r02}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum"
for mode in lstm_step transformer_step beam; do
  python tools/profile_step.py $mode > gpurun_out/${T}_${mode}_plain.log 2>&1 || { echo "plain $mode failed"; tail -5 gpurun_out/${T}_${mode}_plain.log; exit 1; }
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_${mode}_metrics.csv \
      python tools/profile_step.py $mode > gpurun_out/${T}_${mode}_ncu.log 2>&1
  echo "metrics $mode rc=$?"
done
full() {   # full <name> <kernel regex> <skip> <count> <mode>
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 \
      -o gpurun_out/${T}_$1_full -f python tools/profile_step.py $5 > gpurun_out/${T}_full_$1.log 2>&1
  echo "full $1 rc=$?"
}
full lstm_persist lstm_tf_ 0 2 lstm_step
full gemm gemm_tn_kernel 40 2 lstm_step
full dwconv dwconv7_ln_kernel_v 10 2 lstm_step
full mha_tc mha_tc_ 0 2 transformer_step
ls -la gpurun_out | grep ${T}_ | head -30
