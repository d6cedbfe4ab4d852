"""Per-kernel times of nabu_gemm (precision 2 = pre-split fp16 path) on a few layer-sized shapes (dev helper).

usage: python tools/gemm_bench.py [precision]"""
import ctypes
import json
import sys

import torch

from nabu_b200 import engine, lib as L

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib = L.load()
shapes = [(0, 192000, 2048, 1024), (0, 192000, 2048, 64), (0, 192000, 2048, 2048), (1, 192000, 1024, 2048),
          (2, 1024, 2048, 192000), (0, 16384, 2048, 1024)]
for mode, M, N, K in shapes:
    a = torch.randn((K, M) if mode == 2 else (M, K), device='cuda')
    b = torch.randn((N, K) if mode == 1 else (K, N), device='cuda')
    c = torch.zeros((M, N), device='cuda')
    bias = torch.randn(N, device='cuda')
    for it in range(3):
        if it == 1:
            lib.nabu_profile_enable(1)
        engine.gemm(mode, a, b, M, N, K, a.shape[1], b.shape[1], N, C=c, bias=bias, precision=prec)
    torch.cuda.synchronize()
    lib.nabu_profile_enable(0)
    buf = ctypes.create_string_buffer(65536)
    lib.nabu_profile_collect(buf, 65536)
    prof = json.loads(buf.value.decode())
    tf = 2.0 * M * N * K / 1e12
    line = {k: round(v[1] / v[0], 3) for k, v in prof.items()}
    g = [v[1] / v[0] for k, v in prof.items() if k.startswith('gemm')][0]
    print('mode %d M %d N %d K %d: %s  -> %.0f TFLOP/s (fp32-equivalent)' % (mode, M, N, K, line, tf / g * 1e3))
    del a, b, c
