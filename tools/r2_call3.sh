# round 2, call 3: fused tests, full default bench line (verify + config4 + extras), warp sweep incl. 20/24 warps, ncu of both configs
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_capi_parity.py tests/test_gpu_fullsize.py -m gpu -x -q ) > gpurun_out/r2c_tests.log 2>&1; tail -6 gpurun_out/r2c_tests.log
( time python bench.py --steps 10 ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2c_bench.json") if l.startswith("{")][0])
    print("c5 value %.2f G/s step %.4f ms"%(j["value"]/1e9,j["ms_per_step"]), j["roofline"]["kernel_ms_all"], "e2e", j["e2e"] and j["e2e"]["value"])
    print("verify", json.dumps(j["verify"])[:600])
    x=j["extra"]; print("c4 %.2f G/s %.4f ms"%(x["config4"]["value"]/1e9, x["config4"]["ms_per_step"]), x["config4"]["roofline"]["kernel_ms_all"])
    print("sims", json.dumps(x.get("config3_sims_batch"))[:400]); print("ecb", json.dumps(x.get("config1_2_ecb"))[:700])
except Exception as e: print("ERR", e)
PY
run() {  # workload tag env...
  wl=$1; tag=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-extra --verify-windows 0 --steps 10 > gpurun_out/r2c_${wl}_$tag.json 2> gpurun_out/r2c_${wl}_$tag.err
  python - "$wl" "$tag" <<'PY'
import json,sys
wl,tag=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r2c_{wl}_{tag}.json")); k=j["roofline"]["kernel_ms_all"]
    print(f"{wl} {tag:14s} step {j['ms_per_step']:.4f} ms  {j['value']/1e9:6.2f} G/s  k1 {k['k1_count']:.4f} fin {k['finalize']:.4f} k3 {k['k3_small']:.4f} launches {j['gpu_launches_per_step']}")
except Exception as e:
    print(wl, tag, "ERR", e, open(f"gpurun_out/r2c_{wl}_{tag}.err").read()[-300:])
PY
}
for w in 12 13 14; do run config5 w${w}t1 TDSFS_K1_WARPS=$w TDSFS_K1_TILE=1; done
for w in 16 20 24; do for t in 1 2; do run config4 w${w}t${t} TDSFS_K1_WARPS=$w TDSFS_K1_TILE=$t; done; done
for wl in config4 config5; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra --verify-windows 0"
  $CMD > gpurun_out/r2c_plain_$wl.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"k1_fused|k3_finish" -s 6 -c 2 -o gpurun_out/r2c_$wl $CMD > gpurun_out/r2c_ncu_$wl.log 2>&1
  tail -1 gpurun_out/r2c_ncu_$wl.log
done
