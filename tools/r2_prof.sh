# ncu captures of the fused count kernel + finish kernel (config 4 and config 5), after the same plain command exits 0
mkdir -p gpurun_out
for wl in config4 config5; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu"
  $CMD > gpurun_out/r2p_plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"k1_fused|k3_finish" -s 6 -c 2 -o gpurun_out/r2p_$wl $CMD > gpurun_out/r2p_ncu_$wl.log 2>&1
  tail -2 gpurun_out/r2p_ncu_$wl.log
done
