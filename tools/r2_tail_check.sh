# 2-GPU box: multi-GPU tests, then the tail diagnostic (tools/r2_tail_diag.sh)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r2u_tests2.log 2>&1; head -3 gpurun_out/r2u_tests2.log | cut -c1-200
bash tools/r2_tail_diag.sh
