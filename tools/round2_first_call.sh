# First GPU call of round 2 (everything below was written without a GPU at the end of round 1; run under gpurun):
#   1. the regular GPU suite and the default bench (sanity of the committed state)
#   2. configs[2] (sims replicates, batched) -- never measured in round 1
#   3. the experimental pipelined scorer: opt-in parity test, then the bench with it switched on (TDSFS_PIPELINE=1)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_tests.log 2>&1; tail -3 gpurun_out/r2_tests.log
python bench.py --no-cpu > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
python tools/bench_sims.py > gpurun_out/r2_sims.json 2> gpurun_out/r2_sims.err; tail -1 gpurun_out/r2_sims.json; tail -2 gpurun_out/r2_sims.err
( TDSFS_TEST_PIPELINE=1 timeout 200 python -m pytest tests/test_gpu_pipeline_experimental.py -m gpu -x -q ) > gpurun_out/r2_pipe_tests.log 2>&1; tail -5 gpurun_out/r2_pipe_tests.log
for ch in 4 8; do
  TDSFS_PIPELINE=1 TDSFS_PIPELINE_CHUNKS=$ch timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_pipe_c5_$ch.json 2> gpurun_out/r2_pipe_c5_$ch.err
done
python tools/show_bench.py gpurun_out/r2_bench_c5.json gpurun_out/r2_pipe_c5_4.json gpurun_out/r2_pipe_c5_8.json
