"""Golden vectors for `bucket_boundaries` produced by THE REFERENCE'S OWN CODE.

nabu/processing/input_pipeline.py cannot be imported under Python 3 (print statements, TensorFlow imports), but
`bucket_boundaries` (:176-202) is plain Python: this script cuts the function's source text out of
/root/reference, rewrites its one py2 print statement, exec's it and freezes its outputs on seeded histograms.
Run here (the GPU box has no /root/reference):  python tests/golden/make_bucket_golden.py
"""
import json
import os
import re

import numpy as np

SRC = '/root/reference/nabu/processing/input_pipeline.py'
text = open(SRC).read()
body = text[text.index('def bucket_boundaries('):]
body = re.sub(r"print '([^']*)' % \(\s*([^)]*)\)", r"print('\1' % (\2))", body, flags=re.S)
ns = {}
exec(body, ns)
ref = ns['bucket_boundaries']

cases = []
rng = np.random.default_rng(2024)
for n, nb in [(50, 4), (200, 16), (1501, 16), (30, 8), (12, 3), (400, 2)]:
    hist = rng.integers(0, 20, size=n).astype(np.float64)
    hist[:rng.integers(1, max(2, n // 4))] = 0          # no utterances shorter than some minimum, as in real data
    cases.append({'histogram': hist.tolist(), 'numbuckets': nb, 'boundaries': [int(b) for b in ref(hist, nb)]})
# a peaked histogram and a sparse one
h = np.zeros(120); h[40:60] = np.arange(20); h[100] = 500
cases.append({'histogram': h.tolist(), 'numbuckets': 6, 'boundaries': [int(b) for b in ref(h, 6)]})
h = np.zeros(64); h[[5, 17, 33, 63]] = [3, 1, 4, 1]
cases.append({'histogram': h.tolist(), 'numbuckets': 5, 'boundaries': [int(b) for b in ref(h, 5)]})
# more buckets than the data can fill: the boundaries run past the end of the histogram, one bin per bucket
h = np.zeros(15); h[[9, 11, 12, 14]] = 1
cases.append({'histogram': h.tolist(), 'numbuckets': 16, 'boundaries': [int(b) for b in ref(h, 16)]})
h = np.zeros(6); h[3] = 7
cases.append({'histogram': h.tolist(), 'numbuckets': 9, 'boundaries': [int(b) for b in ref(h, 9)]})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'bucket_boundaries.json')
json.dump(cases, open(out, 'w'))
print('wrote %d cases to %s' % (len(cases), out))
