# 8-GPU call: bench at N = 8 and 4 on the same box (strong scaling of the 50 M-SNP workload), tail vs separate exchange kernel, NCCL arm
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --e2e-steps 2 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err; tail -1 gpurun_out/r2m_bench_n8.err | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2m_bench_n4.json 2> gpurun_out/r2m_bench_n4.err; tail -1 gpurun_out/r2m_bench_n4.err | cut -c1-200
TDSFS_NO_TAIL=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e --verify-windows 0 > gpurun_out/r2m_bench_n8_notail.json 2> gpurun_out/r2m_bench_n8_notail.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e --verify-windows 0 --exchange nccl > gpurun_out/r2m_bench_n8_nccl.json 2> gpurun_out/r2m_bench_n8_nccl.err
python bench.py --gpus 1 --steps 20 --no-cpu --no-extra --no-e2e --verify-windows 0 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
python tools/show_bench.py gpurun_out/r2m_bench_n1.json gpurun_out/r2m_bench_n4.json gpurun_out/r2m_bench_n8.json gpurun_out/r2m_bench_n8_notail.json gpurun_out/r2m_bench_n8_nccl.json
python - <<'PY'
import json
for n in ("1","4","8","8_notail","8_nccl"):
    try:
        j=json.load(open(f"gpurun_out/r2m_bench_n{n}.json")); v=j["verify"]
        print(n, "%.2f G/s"%(j["value"]/1e9), "%.4f ms"%j["ms_per_step"], v["snp_count_sum_equals_S"], v["int_checksum"], v["T2D_milli_sum"], v["T1D_p1_milli_sum"], v["T1D_p2_milli_sum"], v.get("oracle",{}).get("ok"), "launches", j["gpu_launches_per_step"])
    except Exception as e: print(n, "ERR", e)
PY
