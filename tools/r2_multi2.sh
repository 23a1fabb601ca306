# 2-GPU call: sharded-scan parity tests (NCCL and peer-memory exchange), then the 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2m_tests2.log 2>&1; tail -6 gpurun_out/r2m_tests2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; tail -2 gpurun_out/r2m_bench_n2.err
python tools/show_bench.py gpurun_out/r2m_bench_n2.json
