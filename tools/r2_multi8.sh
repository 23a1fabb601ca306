# 8-GPU call: bench at N = 8, 4, 2, 1 on the same box (strong scaling of the 50 M-SNP workload), peer-memory and NCCL exchange
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 --no-cpu --e2e-steps 2 > gpurun_out/r2m_bench_n$n.json 2> gpurun_out/r2m_bench_n$n.err; tail -1 gpurun_out/r2m_bench_n$n.err
done
python bench.py --gpus 1 --steps 20 --no-cpu --no-extra --no-e2e > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --no-e2e --exchange nccl > gpurun_out/r2m_bench_n8_nccl.json 2> gpurun_out/r2m_bench_n8_nccl.err
python tools/show_bench.py gpurun_out/r2m_bench_n1.json gpurun_out/r2m_bench_n2.json gpurun_out/r2m_bench_n4.json gpurun_out/r2m_bench_n8.json gpurun_out/r2m_bench_n8_nccl.json
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        j=json.load(open(f"gpurun_out/r2m_bench_n{n}.json")); v=j["verify"]
        print(n, v["snp_count_sum_equals_S"], v["int_checksum"], v["T2D_milli_sum"], v["T1D_p1_milli_sum"], v["T1D_p2_milli_sum"], v.get("oracle",{}).get("ok"))
    except Exception as e: print(n, "ERR", e)
PY
