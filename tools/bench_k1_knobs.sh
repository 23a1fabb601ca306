# K1 ring-geometry sweep (device-resident bench, no e2e / CPU arms): warps x stages per warp x blocks per tile
run() {  # tag workload env...
  tag=$1; wl=$2; shift 2
  env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/k1_${tag}.json 2> gpurun_out/k1_${tag}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/k1_${tag}.json")); k = j["roofline"]["kernel_ms_all"]
    print("${tag} ($wl $*) ms/step %.4f k1 %.4f k3 %.4f" % (j["ms_per_step"], k["k1_count"], k["k3_small"]))
except Exception as e:
    print("${tag} ERR", e)
PY
}
run c5_w6t2 config5 TDSFS_K1_WARPS=6 TDSFS_K1_TILE=2
run c5_w10t2 config5 TDSFS_K1_WARPS=10 TDSFS_K1_TILE=2
run c5_w6t3 config5 TDSFS_K1_WARPS=6 TDSFS_K1_TILE=3
run c5_w5t3 config5 TDSFS_K1_WARPS=5 TDSFS_K1_TILE=3
run c5_w4t4 config5 TDSFS_K1_WARPS=4 TDSFS_K1_TILE=4
run c5_w7t2 config5 TDSFS_K1_WARPS=7 TDSFS_K1_TILE=2
run c5_w8t2_contig config5 TDSFS_K1_WARPS=8 TDSFS_K1_TILE=2 TDSFS_K1_INTERLEAVE=0
TDSFS_K1_WARPS=6 TDSFS_K1_TILE=3 timeout 60 python -m pytest tests/test_gpu_capi_parity.py -m gpu -x -q -k "genotype_scan_vs_oracle or large_windows" 2>&1 | tail -1
TDSFS_K1_WARPS=6 TDSFS_K1_TILE=2 timeout 60 python -m pytest tests/test_gpu_capi_parity.py -m gpu -x -q -k "genotype_scan_vs_oracle or large_windows" 2>&1 | tail -1
