#!/bin/bash
# bench.py one-liners: the fields that matter, one row per workload
for wl in "$@"; do
python bench.py --steps 5 --no-extras --workload $wl 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    r=d['roofline']
    print('$wl', 'value=%.0f GF' % d['value'], 'ms=%.3f' % d['ms_per_step'], 'e2e=%.0f GF' % d['e2e']['value'], 'roofline=%.3f of %.1f %s' % (r['frac'], r['peak'], r['unit']), 'cpu=%.0f GF x%d cores' % (d['cpu_baseline']['value'], d['cpu_baseline']['cores']), d['kernel'], 'launches', d['gpu_launches'])
"
done
