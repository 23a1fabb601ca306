# round 2, call 7: whole suite (graph step included), default bench line with the graph, no-graph A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2g_tests.log 2>&1; tail -5 gpurun_out/r2g_tests.log
( time python bench.py --steps 20 ) > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2g_bench.json") if l.startswith("{")][0])
    print("c5 value %.2f G/s step %.4f ms"%(j["value"]/1e9,j["ms_per_step"]), {k:round(v,4) for k,v in j["roofline"]["kernel_ms_all"].items()}, "e2e", j["e2e"] and j["e2e"]["value"], "whole frac %.3f"%j["roofline"]["whole_step"]["frac"], "launches", j["gpu_launches_per_step"])
    v=j["verify"]; print("verify", v["snp_count_sum_equals_S"], v["int_checksum"], v["T2D_milli_sum"], v["oracle"]["ok"], v["oracle"]["max_rel_err"])
    x=j["extra"]; print("c4 %.2f G/s %.4f ms"%(x["config4"]["value"]/1e9, x["config4"]["ms_per_step"]), {k:round(v,4) for k,v in x["config4"]["roofline"]["kernel_ms_all"].items()}, x["config4"]["verify"].get("snp_count_sum_equals_S"))
    print(j["cpu_baseline"])
except Exception as e: print("ERR", e)
PY
run() {  # workload tag args...
  wl=$1; tag=$2; shift 2
  timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-extra --verify-windows 0 --steps 20 "$@" > gpurun_out/r2g_${wl}_$tag.json 2> gpurun_out/r2g_${wl}_$tag.err
  python - "$wl" "$tag" <<'PY'
import json,sys
wl,tag=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r2g_{wl}_{tag}.json")); k=j["roofline"]["kernel_ms_all"]
    print(f"{wl} {tag:18s} step {j['ms_per_step']:.4f} ms  {j['value']/1e9:6.2f} G/s  k1 {k['k1_count']:.4f} fin {k['finalize']:.4f} k3 {k['k3_small']:.4f} launches {j['gpu_launches_per_step']}")
except Exception as e:
    print(wl, tag, "ERR", e, open(f"gpurun_out/r2g_{wl}_{tag}.err").read()[-300:])
PY
}
for rep in 1 2; do for wl in config5 config4; do run $wl graph_$rep; run $wl nograph_$rep --no-graph; done; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_reference.json 2> gpurun_out/r2g_reference.err; tail -c 600 gpurun_out/r2g_reference.json
