"""Summarise a NABU_REC_TRACE dump (cl_common.cuh): per-phase durations and who is last.

usage: python tools/trace_report.py gpurun_out/trace.fwd_tc.bin [n_ctas]
Phases: 0 step start, 1 producers' counters seen, 2 first block landed, 3 product done, 4 scatter issued,
5 cluster barrier passed, 6 pointwise done, 7 fences done, 8 CTA barrier passed, 9 counter released."""
import sys

import numpy as np

N, PH = 8, 10
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, N, PH)
n = int(sys.argv[2]) if len(sys.argv) > 2 else int((a[:, 0, 0] != 0).sum())
a = a[:n].astype(np.float64)
t0 = a[:, :, 0].min(axis=0)                       # earliest step start per step
names = ['start', 'cnt', 'land', 'mma', 'scat', 'cbar', 'pw', 'fence', 'bar', 'red']
print('CTAs %d; step period (us): %s' % (n, np.round(np.diff(a[:, :, 9].max(axis=0)) / 1e3, 2)))
d = np.diff(a, axis=2)                            # [cta][step][phase-1]
print('phase durations (us), mean over CTAs and steps / max over CTAs (mean over steps):')
for i in range(PH - 1):
    print('  %-6s->%-6s  mean %6.2f   max-cta %6.2f   min-cta %6.2f' % (
        names[i], names[i + 1], d[:, :, i].mean() / 1e3, d[:, :, i].mean(axis=1).max() / 1e3, d[:, :, i].mean(axis=1).min() / 1e3))
rel = (a - t0[None, :, None]) / 1e3
print('time since the earliest step start (us), mean over steps: per phase min / mean / max over CTAs')
for i in range(PH):
    m = rel[:, :, i].mean(axis=1)
    print('  %-6s  min %6.2f (cta %3d)  mean %6.2f  max %6.2f (cta %3d)' % (names[i], m.min(), m.argmin(), m.mean(), m.max(), m.argmax()))
last = a[:, :, 9].argmax(axis=0)
print('last CTA to release per step:', last.tolist())
