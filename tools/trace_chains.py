"""Per-chain view of a NABU_REC_TRACE dump of the chain recurrences (slot = chain * grid + CTA).

Needs a DIAGNOSTIC build in which thread 0 of EVERY chain stamps (a CH_STAMP macro next to CL_STAMP in cl_common.cuh with
slot = chain * gridDim.x + blockIdx.x, TRACE_CTAS = 512); the committed kernels stamp chain 0 only, because the extra
stamps cost 0.6 us per backward time step (profiles/r2h_recurrence_experiments.md).  Kept as the reader of such dumps.


usage: python tools/trace_chains.py gpurun_out/trace.bwd8c.bin <grid CTAs> <chains>
Shows, per chain, when each phase happens inside a step period (mean over CTAs and steps, relative to the earliest
step start of chain 0) and the phase durations: do the chains of a CTA run in lock step (and collide on the tensor pipe /
the DSMEM port) or staggered?"""
import sys

import numpy as np

N, PH = 8, 10
grid, nch = int(sys.argv[2]), int(sys.argv[3])
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, N, PH)[:grid * nch].astype(np.float64).reshape(nch, grid, N, PH)
names = ['start', 'cnt', 'land', 'mma', 'scat', 'cbar', 'pw']
t0 = a[0, :, :, 0].min(axis=0)                       # earliest start of chain 0 per step
period = np.diff(a[:, :, :, 0], axis=2).mean()
print('period %.2f us' % (period / 1e3))
print('phase times (us) relative to chain 0\'s earliest start, mean over CTAs and steps 1..N-1:')
for c in range(nch):
    rel = (a[c] - t0[None, :, None]) / 1e3
    print('  chain %d: ' % c + '  '.join('%s %6.2f' % (names[i], rel[:, 1:, i].mean()) for i in range(7)))
print('phase durations (us):')
for c in range(nch):
    d = np.diff(a[c], axis=2) / 1e3
    print('  chain %d: ' % c + '  '.join('%s>%s %5.2f' % (names[i], names[i + 1], d[:, :, i].mean()) for i in range(6)))
# within one CTA: offsets of the chains' starts against chain 0 (mean / sd over CTAs and steps)
for c in range(1, nch):
    off = (a[c, :, :, 0] - a[0, :, :, 0]) / 1e3
    print('  start(chain %d) - start(chain 0) inside a CTA: mean %.2f  sd %.2f us' % (c, off.mean(), off.std()))
