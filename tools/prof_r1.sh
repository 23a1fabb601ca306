# Profiling recipe of round 1 (run under gpurun, one GPU): the plain command first, then the ncu launch list and ONE full
# capture of the two hot kernels of the same command (B200_PROFILING.md).  usage: bash tools/prof_r1.sh [workload] [snps] [tag]
set -x
WL=${1:-config5}; SNPS=${2:-0}; TAG=${3:-c5}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-e2e --no-cpu"
if [ "$SNPS" != "0" ]; then CMD="$CMD --snps $SNPS"; fi
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k1_genotypes|k3_score_small" -s 6 -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
