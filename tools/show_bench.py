"""Print the interesting part of a bench.py JSON line read from stdin (dev helper)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ''
d = json.loads(sys.stdin.read().strip().split('\n')[-1])
r = d.get('roofline') or {}
sh = r.get('kernel_time_shares', {})
print(tag, 'frames/s %.0f  ms/step %.1f  loss %.6f' % (d['value'], d['ms_per_step'], d.get('loss', float('nan'))),
      {k: round(v * d['ms_per_step'], 1) for k, v in sh.items() if v > 0.004})
