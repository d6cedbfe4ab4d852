"""Print the interesting part of a bench.py JSON line read from stdin (dev helper)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ''
d = json.loads(sys.stdin.read().strip().split('\n')[-1])


def show(d, tag):
    r = d.get('roofline') or {}
    sh = r.get('kernel_time_shares', {})
    print(tag, 'frames/s %.0f  e2e %.0f  ms/step %.2f  loss %s  launches %s' % (
        d['value'], (d.get('e2e') or {}).get('value', float('nan')), d['ms_per_step'], d.get('loss'), d.get('gpu_launches')))
    print('   shares(ms):', {k: round(v * d['ms_per_step'], 2) for k, v in sh.items() if v > 0.004})
    if r:
        print('   roofline: %s frac %.4f us/step %s traffic %s' % (r.get('kernel'), r.get('frac') or 0,
                                                                   r.get('us_per_serial_step_all', r.get('avg_launch_ms')), r.get('traffic')))
    for k in ('roofline_gemm', 'cpu_baseline', 'ctc_loss_delta_vs_cpu', 'strong', 'allreduce', 'clocks'):
        if d.get(k):
            v = dict(d[k])
            for drop in ('cuda', 'cpu_fp64', 'what', 'sample', 'note'):
                v.pop(drop, None)
            print('   %s: %s' % (k, v))


show(d, tag)
if isinstance(d.get('las'), dict):
    if 'value' in d['las']:
        show(d['las'], 'las')
    else:
        print('las:', d['las'])
for k, v in (d.get('decode') or {}).items():
    if isinstance(v, dict):
        print('decode/%s:' % k, {a: b for a, b in v.items() if a not in ('workload',)})
