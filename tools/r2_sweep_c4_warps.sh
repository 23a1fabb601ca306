# config 4, final count kernel: warps per CTA (a warp's range is ~7 windows long at 24 warps: 25,030 windows / 3,552 warps)
mkdir -p gpurun_out
run() {  # tag env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --workload config4 --no-cpu --no-e2e --no-extra --verify-windows 0 --steps 40 --warmup 5 > gpurun_out/r2w_$tag.json 2> gpurun_out/r2w_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r2w_{tag}.json")); k=j["roofline"]["kernel_ms_all"]
    print(f"{tag:10s} step {j['ms_per_step']:.4f} ms  {j['value']/1e9:6.2f} G/s  k1 {k['k1_count']:.4f} fin {k['finalize']:.4f} k3 {k['k3_small']:.4f}")
except Exception as e:
    print(tag, "ERR", e, open(f"gpurun_out/r2w_{tag}.err").read()[-300:])
PY
}
for rep in 1 2; do
  for w in 24 23 22 21 20 18; do run w${w}_$rep TDSFS_K1_WARPS=$w; done
done
