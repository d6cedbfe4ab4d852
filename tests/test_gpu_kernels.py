"""GPU parity tests, kernel level: every C-ABI entry point against the NumPy oracle on the same
seeded inputs.  Tolerance: 1e-4 relative (BASELINE.json north_star) for fp32 values, written next
to each assert; integer outputs are compared exactly."""
import numpy as np
import pytest
import torch

import oracle as O
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('shape', [(128, 128, 64), (200, 72, 40), (37, 29, 515), (1000, 256, 129), (5, 3, 7)])
def test_gemm(mode, shape):
    from nabu_b200 import engine
    M, N, K = shape
    rng = np.random.default_rng(0)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N))
    bias = rng.standard_normal(N); C0 = rng.standard_normal((M, N))
    ref = 0.5 * A @ B + 2.0 * C0 + bias
    a = dev(A if mode != 2 else A.T, torch.float32)
    b = dev(B if mode != 1 else B.T, torch.float32)
    c = dev(C0, torch.float32)
    engine.gemm(mode, a, b, M, N, K, a.shape[1], b.shape[1], N, C=c, alpha=0.5, beta=2.0,
                bias=dev(bias, torch.float32))
    assert rel_err(c.cpu().numpy(), ref) < 2e-5


def test_gemm_large_splitk():
    from nabu_b200 import engine
    rng = np.random.default_rng(1)
    R, M, N = 20000, 40, 256
    A = rng.standard_normal((R, M)).astype(np.float32); B = rng.standard_normal((R, N)).astype(np.float32)
    c = engine.gemm(2, dev(A), dev(B), M, N, R, M, N, N)
    assert rel_err(c.cpu().numpy(), A.astype(np.float64).T @ B.astype(np.float64)) < 2e-5


def _blstm_case(B, T, D, H, ragged, seed, yT=None, need_dx=True):
    from nabu_b200 import lib as L
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    lens = rng.integers(max(1, T // 2), T + 1, size=B).astype(np.int32) if ragged else np.full(B, T, np.int32)
    if ragged:
        lens[0] = T
    p = O.init_blstm_params(rng, D, H)
    yT = T if yT is None else yT
    dy = rng.standard_normal((B, yT, 2 * H)).astype(np.float32)
    y_ref, cache = O.blstm_fwd(x, lens, p, np.float64)
    dx_ref, g_ref = O.blstm_bwd(cache, dy[:, :T].astype(np.float64))

    lib = L.load()
    xd, ld = dev(x), dev(lens)
    pd = {k: dev(v) for k, v in p.items()}
    y = torch.full((B, yT, 2 * H), 7.0, device='cuda')
    gates = torch.empty((2, B, T, 4 * H), device='cuda'); cells = torch.empty((2, B, T, H), device='cuda')
    nws = lib.nabu_blstm_workspace_bytes(B, T, D, H)
    assert nws > 0
    ws = torch.empty(nws, dtype=torch.uint8, device='cuda')
    L.check(lib.nabu_blstm_fwd(L.ptr(xd), L.ptr(ld), B, T, D, H, L.ptr(pd['fw_kernel']), L.ptr(pd['fw_bias']),
                               L.ptr(pd['bw_kernel']), L.ptr(pd['bw_bias']), L.ptr(y), yT, L.ptr(gates),
                               L.ptr(cells), L.ptr(ws), nws, L.stream()), 'fwd')
    torch.cuda.synchronize()
    yh = y.cpu().numpy()
    assert rel_err(yh[:, :T], y_ref) < TOL                       # 1e-4 relative fp32
    assert np.all(yh[:, T:] == 0)
    for b in range(B):
        assert np.all(yh[b, lens[b]:] == 0)
    dyd = dev(dy)
    dx = torch.empty_like(xd) if need_dx else None
    gk = {k: torch.full_like(v, 3.0) for k, v in pd.items()}
    L.check(lib.nabu_blstm_bwd(L.ptr(xd), L.ptr(ld), B, T, D, H, L.ptr(pd['fw_kernel']), L.ptr(pd['bw_kernel']),
                               L.ptr(y), yT, L.ptr(gates), L.ptr(cells), L.ptr(dyd), L.ptr(dx),
                               L.ptr(gk['fw_kernel']), L.ptr(gk['fw_bias']), L.ptr(gk['bw_kernel']),
                               L.ptr(gk['bw_bias']), L.ptr(ws), nws, L.stream()), 'bwd')
    torch.cuda.synchronize()
    if need_dx:
        assert rel_err(dx.cpu().numpy(), dx_ref) < TOL
    for k in g_ref:
        assert rel_err(gk[k].cpu().numpy(), g_ref[k]) < TOL, k


@pytest.mark.parametrize('B,T,D,H,ragged', [
    (3, 7, 5, 64, True),         # tiny batch, TBT=1, hs=2
    (20, 12, 40, 64, True),      # TBT=2
    (40, 9, 24, 128, True),      # TBT=4, hs=2
    (100, 6, 40, 256, True),     # TBT=8, hs=4, padded batch tile
    (130, 5, 16, 64, False),     # two batch tiles
    (16, 10, 40, 512, True),     # hs=8 (cfg-3 width)
    (4, 1, 8, 64, False),        # T=1
    (128, 48, 40, 512, True),    # cfg-3 width and batch: every row of the tcgen05 cluster kernels, flag parity wraps 12 times
    (77, 33, 24, 256, True),     # H=256 variants (clusters of 4 forward, 2 clusters of 8 per direction backward), odd batch
    (200, 21, 40, 512, True),    # more than 128 rows at cfg-3 width: two batch tiles of the tcgen05 recurrences (128 + 72)
    (260, 9, 24, 256, True),     # three tiles (128 + 128 + 4)
    (48, 30, 40, 512, True),     # chains of the backward recurrence: 2 x 32 rows
    (9, 17, 40, 256, True),      # one chain of 16 rows
])
def test_blstm_fwd_bwd(B, T, D, H, ragged):
    _blstm_case(B, T, D, H, ragged, seed=B * 1000 + T)


def test_blstm_padded_output_and_no_dx():
    _blstm_case(6, 7, 12, 64, True, seed=5, yT=8, need_dx=False)


def test_blstm_h1024_sequential_directions():
    _blstm_case(8, 4, 32, 1024, True, seed=11)


@pytest.mark.parametrize('B,T,D', [(32, 24, 40), (40, 11, 64), (20, 9, 2048)])
def test_blstm_h1024_tensor_core_chains(B, T, D):
    """configs[4] width (DBLSTM 6 x 1024): the chain kernels with the directions one after the other, forward weights in
    TMEM, backward weights hi in TMEM / lo in shared memory, backward batch tiles of 32 rows."""
    _blstm_case(B, T, D, 1024, True, seed=B + T)


@pytest.mark.parametrize('B,T,V,L,ragged', [(4, 30, 6, 7, True), (32, 200, 29, 20, True), (2, 5, 3, 2, False),
                                            (3, 40, 29, 0, False), (8, 1500, 29, 150, True)])
def test_ctc(B, T, V, L, ragged):
    from nabu_b200 import engine
    rng = np.random.default_rng(B + T)
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    lens = rng.integers(max(2 * L + 1, T // 2), T + 1, size=B).astype(np.int32) if ragged else np.full(B, T, np.int32)
    Lp = max(L, 1)
    labels = rng.integers(0, V - 1, size=(B, Lp)).astype(np.int32)
    if L > 1:
        labels[0, 1] = labels[0, 0]          # a repeat: needs the blank-between-repeats rule
    ll = rng.integers(max(L // 2, 0), L + 1, size=B).astype(np.int32) if (ragged and L > 0) else np.full(B, L, np.int32)
    loss_ref, grad_ref = O.ctc_loss_and_grad(logits, lens, labels, ll, dtype=np.float64)
    args = (dev(logits), dev(lens), dev(labels), dev(ll))
    loss, grad = engine.ctc_loss_per_utt(*args, want_grad=True)
    e_loss = rel_err(loss.cpu().numpy(), loss_ref)
    e_grad = np.abs(grad.cpu().numpy() - grad_ref).max()
    print('ctc B=%d T=%d: loss rel err %.3g, grad abs err %.3g' % (B, T, e_loss, e_grad))
    assert e_loss < 1e-5, e_loss         # far inside the 1e-4 relative bar
    assert e_grad < 1e-4, e_grad         # posteriors (values in [0,1]): absolute 1e-4


def test_ctc_infeasible_is_inf():
    from nabu_b200 import engine
    logits = np.zeros((1, 3, 4), np.float32)
    labels = np.array([[0, 0, 1, 2]], np.int32)
    args = (dev(logits), dev(np.array([3], np.int32)), dev(labels), dev(np.array([4], np.int32)))
    loss, grad = engine.ctc_loss_per_utt(*args, want_grad=True)
    assert np.isinf(loss.cpu().numpy()[0]) and np.all(grad.cpu().numpy() == 0)


def test_linear_and_ce():
    from nabu_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(3)
    N, D, V = 333, 96, 29
    x = rng.standard_normal((N, D)).astype(np.float32)
    p = O.init_linear_params(rng, D, V)
    p['biases'] = rng.standard_normal(V).astype(np.float32)
    dy = rng.standard_normal((N, V)).astype(np.float32)
    y_ref = O.linear_fwd(x, p)
    dx_ref, g_ref = O.linear_bwd(x, p, dy.astype(np.float64))
    xd, Wd, bd, dyd = dev(x), dev(p['weights']), dev(p['biases']), dev(dy)
    y = torch.empty((N, V), device='cuda'); dx = torch.empty_like(xd); dW = torch.empty_like(Wd); db = torch.empty_like(bd)
    ws = torch.empty(lib.nabu_gemm_workspace_bytes(), dtype=torch.uint8, device='cuda')
    L.check(lib.nabu_linear_fwd(L.ptr(xd), N, D, V, L.ptr(Wd), L.ptr(bd), L.ptr(y), None, 0, L.stream()), 'lf')
    L.check(lib.nabu_linear_bwd(L.ptr(xd), N, D, V, L.ptr(Wd), L.ptr(dyd), L.ptr(dx), L.ptr(dW), L.ptr(db),
                                L.ptr(ws), ws.numel(), L.stream()), 'lb')
    assert rel_err(y.cpu().numpy(), y_ref) < TOL
    assert rel_err(dx.cpu().numpy(), dx_ref) < TOL
    assert rel_err(dW.cpu().numpy(), g_ref['weights']) < TOL
    assert rel_err(db.cpu().numpy(), g_ref['biases']) < TOL
    # masked cross-entropy
    B, U = 7, 11
    logits = rng.standard_normal((B, U, V)).astype(np.float32) * 3
    tl = rng.integers(1, U + 1, size=B).astype(np.int32); tl[0] = U
    tg = rng.integers(0, V, size=(B, U)).astype(np.int32)
    loss_ref, d_ref = O.average_cross_entropy(logits, tg, tl, tl)
    loss = torch.empty(B, device='cuda'); grad = torch.empty((B, U, V), device='cuda')
    lgd, tgd, tld = dev(logits), dev(tg), dev(tl)        # keep the device buffers alive across the call
    L.check(lib.nabu_masked_ce_fwd_bwd(L.ptr(lgd), L.ptr(tgd), U, L.ptr(tld), L.ptr(tld), B, U, V,
                                       1.0 / B, L.ptr(loss), L.ptr(grad), L.stream()), 'ce')
    assert abs(loss.mean().item() - loss_ref) / abs(loss_ref) < TOL
    assert rel_err(grad.cpu().numpy(), d_ref) < TOL


@pytest.mark.parametrize('N,D,V', [(333, 256, 29), (64 * 5 + 1, 1024, 29), (7, 512, 32), (1000, 1024, 5), (4099, 512, 30)])
def test_linear_output_layer_with_few_units(N, D, V):
    """dnn_decoder.py:53-57 at the shapes linear_skinny.cu takes (V <= 32, D % 256 == 0): ragged N, every row tile / row
    chunk boundary, against the fp64 oracle; twice, bit-identical (fixed summation order)."""
    from nabu_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(N + D + V)
    x = rng.standard_normal((N, D)).astype(np.float32)
    p = O.init_linear_params(rng, D, V)
    p['biases'] = rng.standard_normal(V).astype(np.float32)
    dy = rng.standard_normal((N, V)).astype(np.float32)
    y_ref = O.linear_fwd(x, p)
    dx_ref, g_ref = O.linear_bwd(x, p, dy.astype(np.float64))
    xd, Wd, bd, dyd = dev(x), dev(p['weights']), dev(p['biases']), dev(dy)
    ws = torch.empty(lib.nabu_gemm_workspace_bytes(), dtype=torch.uint8, device='cuda')
    outs = []
    for _ in range(2):
        y = torch.full((N, V), 7.0, device='cuda'); dx = torch.empty_like(xd)
        dW = torch.full_like(Wd, 7.0); db = torch.empty_like(bd)
        L.check(lib.nabu_linear_fwd(L.ptr(xd), N, D, V, L.ptr(Wd), L.ptr(bd), L.ptr(y), None, 0, L.stream()), 'lf')
        L.check(lib.nabu_linear_bwd(L.ptr(xd), N, D, V, L.ptr(Wd), L.ptr(dyd), L.ptr(dx), L.ptr(dW), L.ptr(db),
                                    L.ptr(ws), ws.numel(), L.stream()), 'lb')
        outs.append((y.cpu().numpy(), dW.cpu().numpy()))
    assert rel_err(outs[0][0], y_ref) < TOL
    assert rel_err(outs[0][1], g_ref['weights']) < TOL
    assert rel_err(dx.cpu().numpy(), dx_ref) < TOL
    assert rel_err(db.cpu().numpy(), g_ref['biases']) < TOL
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_clip_adam_matches_tf_formula():
    from nabu_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(4)
    n = 10007
    th = rng.standard_normal(n).astype(np.float32); g = (rng.standard_normal(n) * 2).astype(np.float32)
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    thd, md, vd = dev(th), dev(m), dev(v)
    for t in range(1, 4):
        gd = dev(g * t)
        L.check(lib.nabu_clip_adam_step(L.ptr(thd), L.ptr(gd), L.ptr(md), L.ptr(vd), n, 1e-3, t, 0.9, 0.999, 1e-8,
                                        1.0, 1.0, L.stream()), 'adam')
        th, m, v = O.tf_adam_clip(th, g * t, m, v, 1e-3, t, dtype=np.float64)
    # fp32 (1 - beta2) carries a 1.3e-5 relative rounding, in TF's kernel exactly as here
    assert rel_err(thd.cpu().numpy(), th) < 1e-5
    assert rel_err(md.cpu().numpy(), m) < 1e-5 and rel_err(vd.cpu().numpy(), v) < 5e-5


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('shape', [(128, 256, 32), (256, 512, 64), (1000, 260, 132), (300, 520, 40), (132, 257 * 4, 1028),
                                   (4096, 2048, 1024)])
def test_gemm_tensor_core_3xtf32(mode, shape):
    """tcgen05 path: fp32-grade accuracy (hi/lo TF32 split), every operand layout, ragged tiles."""
    from nabu_b200 import engine
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32); C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = 0.5 * A.astype(np.float64) @ B.astype(np.float64) + 2.0 * C0 + bias
    a = dev(A if mode != 2 else np.ascontiguousarray(A.T))
    b = dev(B if mode != 1 else np.ascontiguousarray(B.T))
    c = dev(C0)
    engine.gemm(mode, a, b, M, N, K, a.shape[1], b.shape[1], N, C=c, alpha=0.5, beta=2.0, bias=dev(bias), precision=1)
    err = rel_err(c.cpu().numpy(), ref)
    print('gemm_tc mode %d %s rel err %.3g' % (mode, shape, err))
    assert err < 2e-5          # single-pass TF32 would be ~5e-4; fp32 FFMA is ~1e-6


def test_gemm_tensor_core_splitk_tn():
    from nabu_b200 import engine
    rng = np.random.default_rng(12)
    R, M, N = 40000, 256, 512
    A = rng.standard_normal((R, M)).astype(np.float32); B = rng.standard_normal((R, N)).astype(np.float32)
    c = engine.gemm(2, dev(A), dev(B), M, N, R, M, N, N, precision=1)
    assert rel_err(c.cpu().numpy(), A.astype(np.float64).T @ B.astype(np.float64)) < 2e-5


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('shape', [(128, 256, 64), (256, 512, 128), (1000, 260, 132), (300, 520, 40), (132, 257 * 4, 1028),
                                   (4096, 2048, 1024)])
def test_gemm_tensor_core_split_fp16(mode, shape):
    """tcgen05 kind::f16 on pre-split, power-of-two scaled operands: fp32-grade accuracy, every layout, ragged tiles."""
    from nabu_b200 import engine
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32); C0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = 0.5 * A.astype(np.float64) @ B.astype(np.float64) + 2.0 * C0 + bias
    a = dev(A if mode != 2 else np.ascontiguousarray(A.T))
    b = dev(B if mode != 1 else np.ascontiguousarray(B.T))
    c = dev(C0)
    engine.gemm(mode, a, b, M, N, K, a.shape[1], b.shape[1], N, C=c, alpha=0.5, beta=2.0, bias=dev(bias), precision=2)
    err = rel_err(c.cpu().numpy(), ref)
    print('gemm_h2 mode %d %s rel err %.3g' % (mode, shape, err))
    assert err < 2e-5


def test_gemm_split_fp16_dynamic_range():
    """Rows 1e-30 .. 1e+30 apart (gradients next to activations): the per-row power-of-two scales keep every output
    row at fp32-grade RELATIVE accuracy, which a plain fp16 cast or a global scale cannot."""
    from nabu_b200 import engine
    rng = np.random.default_rng(77)
    M, N, K = 512, 256, 320
    A = rng.standard_normal((M, K)).astype(np.float32)
    A *= (10.0 ** rng.uniform(-30, 30, size=(M, 1))).astype(np.float32)
    Bt = rng.standard_normal((N, K)).astype(np.float32)
    Bt *= (10.0 ** rng.uniform(-6, 6, size=(N, 1))).astype(np.float32)
    ref = A.astype(np.float64) @ Bt.astype(np.float64).T
    c = engine.gemm(1, dev(A), dev(Bt), M, N, K, K, K, N, precision=2).cpu().numpy().astype(np.float64)
    row_scale = np.abs(ref).max(axis=1, keepdims=True)
    col_scale = np.abs(ref / row_scale).max(axis=0, keepdims=True)
    err = np.abs((c - ref) / row_scale / col_scale).max()
    print('gemm_h2 dynamic range err', err)
    assert np.isfinite(c).all() and err < 2e-5


def test_gemm_split_fp16_splitk_tn():
    from nabu_b200 import engine
    rng = np.random.default_rng(12)
    R, M, N = 40000, 256, 512
    A = rng.standard_normal((R, M)).astype(np.float32) * 1e-4
    B = rng.standard_normal((R, N)).astype(np.float32) * 1e3
    c = engine.gemm(2, dev(A), dev(B), M, N, R, M, N, N, precision=2)
    assert rel_err(c.cpu().numpy(), A.astype(np.float64).T @ B.astype(np.float64)) < 2e-5


@pytest.mark.parametrize('B,T,H,pyramid', [
    (128, 40, 512, False),    # cfg-3 width: tcgen05 recurrences write the planes themselves
    (64, 37, 256, True),      # cfg-2 width, odd T: yT = T + 1, the stacked output and its planes feed the next layer
    (6, 9, 64, False),        # FFMA kernels: planes from the fixed-scale split pass
])
def test_blstm_operand_planes_travel_between_layers(B, T, H, pyramid):
    """nabu_blstm_fwd_planes / nabu_blstm_bwd_planes (include/nabu_b200.h): layer 1 writes the fp16 hi/lo operand planes
    of its output, layer 2 consumes them as x_planes (forward: input projection; backward: dKx), both layers hand dZ to
    their GEMMs as planes.  Results against the fp64 oracle of the two-layer stack."""
    from nabu_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(B + T)
    D = 40
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    lens = rng.integers(max(1, T // 2), T + 1, size=B).astype(np.int32)
    lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = 0
    steps = 2 if pyramid else 1
    p1 = O.init_blstm_params(rng, D, H)
    p2 = O.init_blstm_params(rng, 2 * H * steps, H)
    # oracle
    y1_ref, c1 = O.blstm_fwd(x, lens, p1)
    if pyramid:
        x2_ref, lens2 = O.pyramid_stack_fwd(y1_ref, lens, 2)
    else:
        x2_ref, lens2 = y1_ref, lens
    y2_ref, c2 = O.blstm_fwd(x2_ref, lens2, p2)
    dy2 = rng.standard_normal(y2_ref.shape).astype(np.float32)
    for b in range(B):
        dy2[b, lens2[b]:] = 0
    dx2_ref, g2_ref = O.blstm_bwd(c2, dy2.astype(np.float64))
    dy1_ref = O.pyramid_stack_bwd(dx2_ref, T, 2) if pyramid else dx2_ref
    dx1_ref, g1_ref = O.blstm_bwd(c1, dy1_ref)

    def layer_fwd(xd, ld, p, Tl, Dl, yT, x_planes):
        pd = {k: dev(v) for k, v in p.items()}
        y = torch.full((B, yT, 2 * H), 7.0, device='cuda')
        gates = torch.empty((2, B, Tl, 4 * H), device='cuda')
        cells = torch.empty((2, B, Tl, H), device='cuda')
        planes = torch.full((lib.nabu_blstm_planes_bytes(B, yT, H),), 0x55, dtype=torch.uint8, device='cuda')
        nws = lib.nabu_blstm_workspace_bytes(B, Tl, Dl, H)
        ws = torch.empty(nws, dtype=torch.uint8, device='cuda')
        L.check(lib.nabu_blstm_fwd_planes(L.ptr(xd), L.ptr(x_planes), L.ptr(ld), B, Tl, Dl, H, L.ptr(pd['fw_kernel']),
                                          L.ptr(pd['fw_bias']), L.ptr(pd['bw_kernel']), L.ptr(pd['bw_bias']), L.ptr(y),
                                          L.ptr(planes), yT, L.ptr(gates), L.ptr(cells), L.ptr(ws), nws, L.stream()), 'fwd')
        return dict(pd=pd, y=y, gates=gates, cells=cells, planes=planes, ws=ws, nws=nws, x=xd, len=ld, x_planes=x_planes,
                    T=Tl, D=Dl, yT=yT)

    def layer_bwd(st, dyd, need_dx):
        pd = st['pd']
        dx = torch.empty_like(st['x']) if need_dx else None
        gk = {k: torch.full_like(v, 3.0) for k, v in pd.items()}
        L.check(lib.nabu_blstm_bwd_planes(L.ptr(st['x']), L.ptr(st['x_planes']), L.ptr(st['len']), B, st['T'], st['D'], H,
                                          L.ptr(pd['fw_kernel']), L.ptr(pd['bw_kernel']), L.ptr(st['y']), L.ptr(st['planes']),
                                          st['yT'], L.ptr(st['gates']), L.ptr(st['cells']), L.ptr(dyd), L.ptr(dx),
                                          L.ptr(gk['fw_kernel']), L.ptr(gk['fw_bias']), L.ptr(gk['bw_kernel']),
                                          L.ptr(gk['bw_bias']), L.ptr(st['ws']), st['nws'], L.stream()), 'bwd')
        return dx, gk

    yT1 = (T + steps - 1) // steps * steps
    s1 = layer_fwd(dev(x), dev(lens), p1, T, D, yT1, None)
    # the planes hold y * 32 as hi + lo / 2048 (fp16), zero rows past the lengths and in the padding
    half = s1['planes'].numel() // 2
    hi = s1['planes'][:half].view(torch.float16)[:B * yT1 * 2 * H].float().view(B, yT1, 2 * H)
    lo = s1['planes'][half:].view(torch.float16)[:B * yT1 * 2 * H].float().view(B, yT1, 2 * H)
    rebuilt = ((hi.double() + lo.double() / 2048) / 32).cpu().numpy()
    assert np.abs(rebuilt - s1['y'].double().cpu().numpy()).max() < 2e-7
    assert np.all(rebuilt[:, T:] == 0) and np.all(rebuilt[1, lens[1]:] == 0)
    if pyramid:
        x2 = s1['y'].view(B, yT1 // 2, 4 * H)
        l2 = dev(lens2.astype(np.int32))
    else:
        x2, l2 = s1['y'], dev(lens)
    T2, D2 = x2.shape[1], x2.shape[2]
    s2 = layer_fwd(x2, l2, p2, T2, D2, T2, s1['planes'])
    assert rel_err(s2['y'].cpu().numpy(), y2_ref) < TOL
    dx2, g2 = layer_bwd(s2, dev(dy2), True)
    assert rel_err(dx2.cpu().numpy(), dx2_ref) < TOL
    for k in g2_ref:
        assert rel_err(g2[k].cpu().numpy(), g2_ref[k]) < TOL, ('layer 2', k)
    dy1 = dx2.view(B, yT1, 2 * H)
    dx1, g1 = layer_bwd(s1, dy1, True)
    torch.cuda.synchronize()
    assert rel_err(dx1.cpu().numpy(), dx1_ref) < TOL
    for k in g1_ref:
        assert rel_err(g1[k].cpu().numpy(), g1_ref[k]) < TOL, ('layer 1', k)
    if H < 256:
        return                                           # the fallback kernels overwrite the saved gates: no second pass
    # nabu_blstm_bwd_hints: layer 2 leaves max |dx| behind (from the dX contraction's epilogue), layer 1 takes it instead
    # of reading dy -- bit-identical results
    hint = torch.full((128,), -1, dtype=torch.int32, device='cuda')
    L.check(lib.nabu_blstm_bwd_hints(L.ptr(hint), None), 'hints')
    dx2b, g2b = layer_bwd(s2, dev(dy2), True)
    assert torch.equal(dx2b, dx2)
    h = hint.cpu().numpy()
    assert np.all(h[1:] == 0) and h[:1].view(np.float32)[0] == np.abs(dx2b.cpu().numpy()).max()
    L.check(lib.nabu_blstm_bwd_hints(None, L.ptr(hint)), 'hints')
    dx1b, g1b = layer_bwd(s1, dx2b.view(B, yT1, 2 * H), True)
    torch.cuda.synchronize()
    assert torch.equal(dx1b, dx1)
    for k in g1:
        assert torch.equal(g1b[k], g1[k]), ('layer 1 with the max |dy| hint', k)
    # a hint is consumed by ONE call: the next call computes max |dy| itself again
    dx1c, _ = layer_bwd(s1, dx2b.view(B, yT1, 2 * H), True)
    assert torch.equal(dx1c, dx1)


def test_out_of_range_labels_are_flagged_not_dereferenced():
    """a label that is not a class of the output (a reader's none-symbol -1, or output_dims too small) gives
    loss = +inf and a zero gradient for that utterance (TF raises / returns NaN) instead of indexing the logits out of
    bounds; the other utterances of the batch are unaffected."""
    from nabu_b200 import engine
    rng = np.random.default_rng(0)
    B, T, V, L = 3, 12, 5, 3
    logits = rng.standard_normal((B, T, V)).astype(np.float32)
    labels = np.array([[0, 1, 2], [1, -1, 0], [2, V - 1, 1]], np.int32)      # -1, and the blank index as a label
    lens, ll = np.full(B, T, np.int32), np.full(B, L, np.int32)
    loss, grad = engine.ctc_loss_per_utt(dev(logits), dev(lens), dev(labels), dev(ll), want_grad=True)
    loss, grad = loss.cpu().numpy(), grad.cpu().numpy()
    ref, _ = O.ctc_loss_and_grad(logits[:1], lens[:1], labels[:1], ll[:1])
    assert abs(loss[0] - ref[0]) < 1e-4 * abs(ref[0])
    assert np.isinf(loss[1]) and np.isinf(loss[2]) and np.all(grad[1:] == 0) and np.isfinite(grad[0]).all()
    # masked cross entropy
    U = 4
    lg = rng.standard_normal((B, U, V)).astype(np.float32)
    tg = np.array([[0, 1, 2, 3], [1, V, 0, 0], [2, -3, 1, 1]], np.int32)
    tl = np.full(B, U, np.int32)
    x = torch.tensor(lg, device='cuda', requires_grad=True)
    out = engine.masked_ce_mean(x, dev(tg), dev(tl), dev(tl))
    assert np.isinf(float(out))
