import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[:300]); continue
    print('==', f, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'roof', round(d['roofline']['achieved'], 1),
          round(d['roofline']['frac'], 3), 'match pairs/s', round(d['match']['pairs_per_s']), 'clk', d.get('clocks'))
    print('  kernels', {k: round(v, 4) for k, v in d['kernels_ms_per_image'].items()})
    print('  layers', {k: round(v, 4) for k, v in d['layers_ms_per_image'].items()})
    if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline'])
