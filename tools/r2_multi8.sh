# 8-GPU call: bench at N = 8, 4, 2 on the same box (strong scaling of the 50 M-SNP workload)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 --no-cpu --e2e-steps 2 > gpurun_out/r2m_bench_n$n.json 2> gpurun_out/r2m_bench_n$n.err; tail -1 gpurun_out/r2m_bench_n$n.err | cut -c1-200
done
python bench.py --gpus 1 --steps 20 --no-cpu --no-extra --no-e2e > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
python tools/show_bench.py gpurun_out/r2m_bench_n1.json gpurun_out/r2m_bench_n2.json gpurun_out/r2m_bench_n4.json gpurun_out/r2m_bench_n8.json
python - <<'PY'
import json
for n in ("1","2","4","8"):
    try:
        j=json.load(open(f"gpurun_out/r2m_bench_n{n}.json")); v=j["verify"]
        print(n, "%.2f G/s"%(j["value"]/1e9), "%.4f ms"%j["ms_per_step"], v["snp_count_sum_equals_S"], v["int_checksum"], v["T2D_milli_sum"], v["T1D_p1_milli_sum"], v["T1D_p2_milli_sum"], v.get("oracle",{}).get("ok"), "launches", j["gpu_launches_per_step"], "e2e", j["e2e"] and "%.3g"%j["e2e"]["value"])
    except Exception as e: print(n, "ERR", e)
PY
