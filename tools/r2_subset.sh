# last check of the round (3 GPU-minutes left): fused-path and C-ABI parity tests + smoke on the shipped binary
mkdir -p gpurun_out
( time timeout 140 python -m pytest tests/test_gpu_fused.py tests/test_gpu_capi_parity.py tests/test_gpu_fullsize.py -m gpu -x -q ) > gpurun_out/r2v_tests.log 2>&1; tail -4 gpurun_out/r2v_tests.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
