# Round-2 ncu evidence (run under gpurun, one GPU): per-kernel metrics for EVERY kernel of one eager LSTM train step, one
# Transformer train step and a few beam-search steps, plus `--set full` captures of the top kernels.
# Summaries: python tools/ncu_summary.py metrics <csv> / full <ncu-rep>  ->  profiles/r02_*.txt
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum"
for mode in lstm_step transformer_step beam; do
  python tools/profile_step.py $mode > gpurun_out/r02_${mode}_plain.log 2>&1 || { echo "plain $mode failed"; tail -5 gpurun_out/r02_${mode}_plain.log; exit 1; }
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_${mode}_metrics.csv \
      python tools/profile_step.py $mode > gpurun_out/r02_${mode}_ncu.log 2>&1
  echo "metrics $mode rc=$?"
done
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lstm_tf_ -c 2 \
    -o gpurun_out/r02_lstm_persist_full python tools/profile_step.py lstm_step > gpurun_out/r02_full1.log 2>&1
echo "full lstm rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tn_kernel -s 40 -c 2 \
    -o gpurun_out/r02_gemm_full python tools/profile_step.py lstm_step > gpurun_out/r02_full2.log 2>&1
echo "full gemm rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dwconv7_ln_kernel_v2 -s 10 -c 1 \
    -o gpurun_out/r02_dwconv_full python tools/profile_step.py lstm_step > gpurun_out/r02_full3.log 2>&1
echo "full dwconv rc=$?"
ls -la gpurun_out | grep r02_ | head -20
