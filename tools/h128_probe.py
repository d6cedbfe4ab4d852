"""Times one BLSTM layer forward + backward at num_units = 128 (what every shipped recipe uses) on the kernels the
library picks by itself and with NABU_PAD_UNITS=1 (zero-padded to the 256-unit tcgen05 kernels)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nabu_b200 import engine  # noqa: E402


def run(B, T, D, H, pad):
    os.environ['NABU_PAD_UNITS'] = '1' if pad else '0'
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(1)

    class V(object):
        def __init__(self, shape):
            self.data = (torch.rand(shape, generator=g) * 0.2 - 0.1).to(dev).requires_grad_(True)
            self.grad = torch.zeros(shape, device=dev)
    vs = [V((D + H, 4 * H)), V((4 * H,)), V((D + H, 4 * H)), V((4 * H,))]
    x = torch.randn((B, T, D), generator=g).to(dev).requires_grad_(True)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    dy = torch.randn((B, T, 2 * H), generator=g).to(dev)
    for it in range(3):
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
    return e0.elapsed_time(e1), e1.elapsed_time(e2), float(y.abs().sum())


if __name__ == '__main__':
    for B, T, D, H in ((16, 800, 40, 128), (128, 800, 40, 128), (128, 800, 256, 128), (64, 800, 512, 192)):
        for pad in (0, 1):
            f, b, chk = run(B, T, D, H, pad)
            print('B=%d T=%d D=%d H=%d pad=%d: fwd %.2f ms (%.2f us/step)  bwd %.2f ms (%.2f us/step)  checksum %.4f'
                  % (B, T, D, H, pad, f, f * 1e3 / T, b, b * 1e3 / T, chk), flush=True)
