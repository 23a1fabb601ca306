# round 2, call 4: whole GPU suite on the streaming finish kernel + nearest-boundary ranges, default bench line, warp sweep
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_tests.log 2>&1; tail -6 gpurun_out/r2d_tests.log
( time python bench.py --steps 20 ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2d_bench.json") if l.startswith("{")][0])
    print("c5 value %.2f G/s step %.4f ms"%(j["value"]/1e9,j["ms_per_step"]), {k:round(v,4) for k,v in j["roofline"]["kernel_ms_all"].items()}, "e2e", j["e2e"] and j["e2e"]["value"], "whole frac", j["roofline"]["whole_step"]["frac"])
    print("verify", json.dumps(j["verify"])[:700])
    x=j["extra"]; print("c4 %.2f G/s %.4f ms"%(x["config4"]["value"]/1e9, x["config4"]["ms_per_step"]), {k:round(v,4) for k,v in x["config4"]["roofline"]["kernel_ms_all"].items()}, x["config4"]["verify"].get("snp_count_sum_equals_S"))
except Exception as e: print("ERR", e)
PY
run() {  # workload tag env...
  wl=$1; tag=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-extra --verify-windows 0 --steps 20 > gpurun_out/r2d_${wl}_$tag.json 2> gpurun_out/r2d_${wl}_$tag.err
  python - "$wl" "$tag" <<'PY'
import json,sys
wl,tag=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r2d_{wl}_{tag}.json")); k=j["roofline"]["kernel_ms_all"]
    print(f"{wl} {tag:14s} step {j['ms_per_step']:.4f} ms  {j['value']/1e9:6.2f} G/s  k1 {k['k1_count']:.4f} fin {k['finalize']:.4f} k3 {k['k3_small']:.4f} launches {j['gpu_launches_per_step']}")
except Exception as e:
    print(wl, tag, "ERR", e, open(f"gpurun_out/r2d_{wl}_{tag}.err").read()[-300:])
PY
}
for w in 12 14; do run config5 w${w}t1 TDSFS_K1_WARPS=$w TDSFS_K1_TILE=1; done
for w in 16 17 20 24; do run config4 w${w}t1 TDSFS_K1_WARPS=$w TDSFS_K1_TILE=1; done
for w in 16 17; do run config4 w${w}t2 TDSFS_K1_WARPS=$w TDSFS_K1_TILE=2; done
