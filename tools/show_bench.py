import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[:300]); continue
    print('==', f, 'value', round(d['value'], 1), 'e2e(single call)', round(d['e2e']['value'], 1),
          'e2e(batched)', round(d['e2e'].get('batched', {}).get('value', 0), 1), 'roof', round(d['roofline']['achieved'], 1),
          round(d['roofline']['frac'], 3), 'clk', d.get('clocks'))
    print('  kernels', {k: round(v, 4) for k, v in d['kernels_ms_per_image'].items()})
    print('  layers', {k: round(v, 4) for k, v in d['layers_ms_per_image'].items()})
    m = d['match']
    print('  match pairs/s', round(m['pairs_per_s']), 'ms/pair', round(m['ms_per_pair'], 4), 'kernel_ms', m.get('kernel_ms'), 'prep', m.get('prep_kernel_ms'),
          'roof', m.get('roofline', {}).get('frac'))
    for k in ('grouped', 'e2e', 'one_to_many', 'cpu_baseline'):
        if k in m: print('   ', k, m[k])
    for k in ('pairs', 'sweep', 'other_modes', 'cpu_baseline'):
        if k in d: print(' ', k, d[k])
