# round 2, first GPU call: GPU test-suite with the fused scan on by default, then the bench (config 5 / config 4) and A/B arms
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_tests.log 2>&1; tail -15 gpurun_out/r2a_tests.log
python bench.py --no-cpu --steps 10 > gpurun_out/r2a_c5.json 2> gpurun_out/r2a_c5.err; tail -3 gpurun_out/r2a_c5.err
python bench.py --workload config4 --no-cpu --steps 20 --no-e2e > gpurun_out/r2a_c4.json 2> gpurun_out/r2a_c4.err; tail -3 gpurun_out/r2a_c4.err
for arm in NO_FUSE K1_OLD REC_WIDE NO_POS_TMA; do
  for wl in config5 config4; do
    env TDSFS_$arm=1 timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 10 > gpurun_out/r2a_${wl}_$arm.json 2> gpurun_out/r2a_${wl}_$arm.err
  done
done
python tools/show_bench.py gpurun_out/r2a_c5.json gpurun_out/r2a_c4.json gpurun_out/r2a_config*_*.json
