# scorer variants side by side (device-resident bench, no e2e / CPU arms): TDSFS_K3_MODE = walk | incr2d | incr1d
for wl in config5 config4; do
  for m in walk incr2d incr1d; do
    TDSFS_K3_MODE=$m python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/var_${wl}_$m.json 2> gpurun_out/var_${wl}_$m.err
    python - <<PY
import json
try:
    j = json.load(open("gpurun_out/var_${wl}_$m.json")); k = j["roofline"]["kernel_ms_all"]
    print("$wl $m ms/step %.4f k3_small %.4f k1 %.4f" % (j["ms_per_step"], k["k3_small"], k["k1_count"]))
except Exception as e:
    print("$wl $m ERR", e)
PY
  done
done
