# quick GPU check used during development: parity tests, then both synthetic configs (device-resident + e2e)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --workload config4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -3 gpurun_out/bench_c5.err
