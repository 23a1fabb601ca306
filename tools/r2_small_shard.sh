# the shard of one rank of an 8-GPU run (6.25 M SNPs of config 5) on one GPU: count-kernel warps and finish grid at that size
mkdir -p gpurun_out
run() {  # tag env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --workload config5 --snps 6250000 --no-cpu --no-e2e --no-extra --verify-windows 0 --steps 50 --warmup 5 > gpurun_out/r2s_$tag.json 2> gpurun_out/r2s_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r2s_{tag}.json")); k=j["roofline"]["kernel_ms_all"]
    print(f"{tag:10s} step {j['ms_per_step']:.4f} ms  k1 {k['k1_count']:.4f} fin {k['finalize']:.4f} k3 {k['k3_small']:.4f} windows {j['config']['windows']}")
except Exception as e:
    print(tag, "ERR", e, open(f"gpurun_out/r2s_{tag}.err").read()[-300:])
PY
}
for rep in 1 2; do
  run w14_$rep X=1
  run w13_$rep TDSFS_K1_WARPS=13
  run w12_$rep TDSFS_K1_WARPS=12
  run w11_$rep TDSFS_K1_WARPS=11
  run w10_$rep TDSFS_K1_WARPS=10
done
