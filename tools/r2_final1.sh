# final 1-GPU call of round 2: whole GPU suite, smoke, the default bench line and the reference arm exactly as the driver runs them
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2z_tests.log 2>&1; tail -4 gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/r2z_reference.json 2> gpurun_out/r2z_reference.err
( time python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -4 gpurun_out/r2z_bench.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r2z_bench.json") if l.startswith("{")][0])
    print("c5 value %.2f G/s step %.4f ms"%(j["value"]/1e9,j["ms_per_step"]), {k:round(v,4) for k,v in j["roofline"]["kernel_ms_all"].items()}, "e2e %.3g"%j["e2e"]["value"], "K1 frac %.3f whole frac %.3f"%(j["roofline"]["frac"], j["roofline"]["whole_step"]["frac"]), "launches", j["gpu_launches_per_step"], "traffic", j["roofline"]["traffic"])
    v=j["verify"]; print("verify", v["snp_count_sum_equals_S"], v["int_checksum"], v["T2D_milli_sum"], v["oracle"]["ok"], v["oracle"]["max_rel_err"], "e2e same:", j["e2e"]["results_equal_device_resident_run"])
    x=j["extra"]; print("c4 %.2f G/s %.4f ms"%(x["config4"]["value"]/1e9, x["config4"]["ms_per_step"]), {k:round(v,4) for k,v in x["config4"]["roofline"]["kernel_ms_all"].items()}, x["config4"]["verify"].get("snp_count_sum_equals_S"))
    print("sims", json.dumps(x.get("config3_sims_batch"))[:300]); print("ecb", json.dumps(x.get("config1_2_ecb"))[:300])
    print(j["cpu_baseline"]["value"], j["clocks"])
    r=json.loads([l for l in open("gpurun_out/r2z_reference.json") if l.startswith("{")][0]); print("reference", r["value"], r["cpu_baseline"]["cores"])
except Exception as e: print("ERR", e)
PY
