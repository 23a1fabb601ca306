# 8-GPU box, shipped binary: the N = 8 bench line only (strong scaling of config 5)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu --e2e-steps 2 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err; tail -1 gpurun_out/r2n_bench_n8.err | cut -c1-200
python tools/show_bench.py gpurun_out/r2n_bench_n8.json
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2n_bench_n8.json")); v=j["verify"]
print("%.2f G/s"%(j["value"]/1e9), "%.4f ms"%j["ms_per_step"], v["snp_count_sum_equals_S"], v["int_checksum"], v["T2D_milli_sum"], v["T1D_p1_milli_sum"], v["T1D_p2_milli_sum"], v.get("oracle",{}).get("ok"), "e2e", j["e2e"] and "%.3g"%j["e2e"]["value"])
PY
