mkdir -p gpurun_out
( TDSFS_NO_POS_TMA=1 timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -x -q ) > gpurun_out/r2l_tests_nopostma.log 2>&1; tail -2 gpurun_out/r2l_tests_nopostma.log
run() {  # workload tag env...
  wl=$1; tag=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --no-cpu --no-e2e --no-extra --verify-windows 16 --steps 20 > gpurun_out/r2l_${wl}_$tag.json 2> gpurun_out/r2l_${wl}_$tag.err
  python - "$wl" "$tag" <<'PY'
import json,sys
wl,tag=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r2l_{wl}_{tag}.json")); k=j["roofline"]["kernel_ms_all"]; v=j["verify"]
    print(f"{wl} {tag:10s} step {j['ms_per_step']:.4f} ms  {j['value']/1e9:6.2f} G/s  k1 {k['k1_count']:.4f} fin {k['finalize']:.4f} k3 {k['k3_small']:.4f} verify {v['snp_count_sum_equals_S']} {v['int_checksum']} {v['T2D_milli_sum']} {v['oracle']['ok']}")
except Exception as e:
    print(wl, tag, "ERR", e, open(f"gpurun_out/r2l_{wl}_{tag}.err").read()[-300:])
PY
}
for rep in 1 2; do for wl in config5 config4; do run $wl tma_$rep X=1; run $wl ldg_$rep TDSFS_NO_POS_TMA=1; done; done
