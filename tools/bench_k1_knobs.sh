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
run c5_w12d1t1 config5 TDSFS_K1_WARPS=12 TDSFS_K1_DEPTH=1
run c5_w16d1t1 config5 TDSFS_K1_WARPS=16 TDSFS_K1_DEPTH=1
run c5_w14d1t1 config5 TDSFS_K1_WARPS=14 TDSFS_K1_DEPTH=1
run c5_w10d1t1 config5 TDSFS_K1_WARPS=10 TDSFS_K1_DEPTH=1
run c5_w16d1t1_contig config5 TDSFS_K1_WARPS=16 TDSFS_K1_DEPTH=1 TDSFS_K1_INTERLEAVE=0
run c5_w12d1t2 config5 TDSFS_K1_WARPS=12 TDSFS_K1_DEPTH=1 TDSFS_K1_TILE=2
run c5_w8d1t2 config5 TDSFS_K1_WARPS=8 TDSFS_K1_DEPTH=1 TDSFS_K1_TILE=2
run c5_w12d2t1 config5 TDSFS_K1_WARPS=12 TDSFS_K1_DEPTH=2
run c4_w12d2t2 config4 TDSFS_K1_WARPS=12 TDSFS_K1_DEPTH=2
run c4_w16d1t2 config4 TDSFS_K1_WARPS=16 TDSFS_K1_DEPTH=1
run c4_w16d2t1 config4 TDSFS_K1_WARPS=16 TDSFS_K1_DEPTH=2 TDSFS_K1_TILE=1
run c4_w16d1t3 config4 TDSFS_K1_WARPS=16 TDSFS_K1_DEPTH=1 TDSFS_K1_TILE=3
run c4_w16d1t4 config4 TDSFS_K1_WARPS=16 TDSFS_K1_DEPTH=1 TDSFS_K1_TILE=4
run c4_w12d1t4 config4 TDSFS_K1_WARPS=12 TDSFS_K1_DEPTH=1 TDSFS_K1_TILE=4
run c4_w12d1t3 config4 TDSFS_K1_WARPS=12 TDSFS_K1_DEPTH=1 TDSFS_K1_TILE=3
run c4_w16d1t2_contig config4 TDSFS_K1_WARPS=16 TDSFS_K1_DEPTH=1 TDSFS_K1_INTERLEAVE=0
