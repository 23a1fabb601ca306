# Profiling recipe of round 2 (run under gpurun, one GPU): per workload the plain command first, then the ncu launch list of
# the same command and ONE full capture of the two hot kernels (B200_PROFILING.md).  The bench is run call by call
# (--no-graph) so that every kernel launch is visible to ncu.
set -x
mkdir -p gpurun_out
for WL in config5 config4; do
  CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra --no-graph --verify-windows 0"
  $CMD > gpurun_out/r02_plain_$WL.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_${WL}_launches.csv $CMD > gpurun_out/r02_ncu_launch_$WL.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"k1_fused|k3_finish|k_finalize_counts" -s 9 -c 3 -o gpurun_out/r02_$WL $CMD > gpurun_out/r02_ncu_full_$WL.log 2>&1
  tail -1 gpurun_out/r02_ncu_full_$WL.log
done
