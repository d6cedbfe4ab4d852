"""One BLSTM layer forward + backward at cfg-3 width (D = 1024, H = 512) for small per-GPU batches (the strong-scaling
split of the 128-utterance minibatch): which recurrence kernels win at B = 16 / 32 / 64?  Run once per kernel choice
(NABU_REC_FWD / NABU_REC_BWD are read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nabu_b200 import engine  # noqa: E402


def run(B, T, D, H):
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(1)

    class V(object):
        def __init__(self, shape):
            self.data = (torch.rand(shape, generator=g) * 0.1 - 0.05).to(dev).requires_grad_(True)
            self.grad = torch.zeros(shape, device=dev)
    vs = [V((D + H, 4 * H)), V((4 * H,)), V((D + H, 4 * H)), V((4 * H,))]
    x = (torch.randn((B, T, D), generator=g) * 0.3).to(dev).requires_grad_(True)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    dy = torch.randn((B, T, 2 * H), generator=g).to(dev)
    for it in range(2):
        y = engine.blstm(x, lens, vs[0], vs[1], vs[2], vs[3], H)
        y.backward(dy)
        engine.side_join()
    torch.cuda.synchronize()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record()
    y = engine.blstm(x, lens, vs[0], vs[1], vs[2], vs[3], H)
    e1.record()
    y.backward(dy)
    engine.side_join()
    e2.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), e1.elapsed_time(e2)


if __name__ == '__main__':
    tag = 'fwd=%s bwd=%s' % (os.environ.get('NABU_REC_FWD', 'auto'), os.environ.get('NABU_REC_BWD', 'auto'))
    T = 600
    for B in (16, 32, 64, 128):
        f, b = run(B, T, 1024, 512)
        print('%s B=%3d: layer fwd %.2f ms (%.2f us/step incl. GEMM)  bwd %.2f ms (%.2f us/step incl. GEMMs)'
              % (tag, B, f, f * 1e3 / T, b, b * 1e3 / T), flush=True)
