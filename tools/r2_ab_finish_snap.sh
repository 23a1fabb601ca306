# round 2, call 5: suite (both finish variants for the fused tests), then same-box A/B of finish variant x boundary snapping
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2e_tests.log 2>&1; tail -4 gpurun_out/r2e_tests.log
( TDSFS_FINISH_STREAM=1 timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -x -q ) > gpurun_out/r2e_tests_stream.log 2>&1; tail -3 gpurun_out/r2e_tests_stream.log
run() {  # workload tag env...
  wl=$1; tag=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-extra --verify-windows 0 --steps 20 > gpurun_out/r2e_${wl}_$tag.json 2> gpurun_out/r2e_${wl}_$tag.err
  python - "$wl" "$tag" <<'PY'
import json,sys
wl,tag=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r2e_{wl}_{tag}.json")); k=j["roofline"]["kernel_ms_all"]
    print(f"{wl} {tag:18s} step {j['ms_per_step']:.4f} ms  {j['value']/1e9:6.2f} G/s  k1 {k['k1_count']:.4f} fin {k['finalize']:.4f} k3 {k['k3_small']:.4f} launches {j['gpu_launches_per_step']}")
except Exception as e:
    print(wl, tag, "ERR", e, open(f"gpurun_out/r2e_{wl}_{tag}.err").read()[-300:])
PY
}
for rep in 1 2; do
  for wl in config5 config4; do
    run $wl base_$rep X=1
    run $wl stream_$rep TDSFS_FINISH_STREAM=1
    run $wl snapback_$rep TDSFS_SNAP_BACK=1
    run $wl old_$rep TDSFS_K1_OLD=1
  done
done
run config4 w24t1 TDSFS_K1_WARPS=24 TDSFS_K1_TILE=1
run config4 w24t1_stream TDSFS_K1_WARPS=24 TDSFS_K1_TILE=1 TDSFS_FINISH_STREAM=1
