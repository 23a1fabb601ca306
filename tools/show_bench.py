import json, sys
for f in sys.argv[1:] or ('gpurun_out/bench_c4.json', 'gpurun_out/bench_c5.json'):
    try:
        j = json.load(open(f)); r = j['roofline']
        print(f, 'value %.4g SNPs/s' % j['value'], 'ms/step %.3f' % j['ms_per_step'], 'K1 frac %.3f' % r['frac'], 'step frac %.3f' % r['whole_step']['frac'],
              {k: round(v, 4) for k, v in r['kernel_ms_all'].items()}, 'e2e %.3g' % (j['e2e'] or {}).get('value', 0), j['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
