"""Row a8 through its own C-ABI entry points (nabu_attn_keys / nabu_attn_step_fwd / nabu_attn_step_bwd):
forward against the oracle's attention_step, backward against torch autograd over a float64 twin of the same formulas
(components/attention.py:142-240).  Includes the cfg-2 width (A = 256, E = 512, numfilt 10, filtersize 201, T' = 125)."""
import ctypes

import numpy as np
import pytest
import torch

import oracle as O
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _twin(query, prev, values, keys, mask, p, attention):
    """torch float64 restatement (independent index arithmetic for the 'same' convolution)"""
    q = query @ p['query_kernel']
    pre = q[:, None, :] + keys
    if attention == 'location_aware':
        Wc = p['conv_kernel']                                   # [k,1,F]
        k = Wc.shape[0]
        padl = (k - 1) // 2
        x = torch.nn.functional.pad(prev[:, None, :], (padl, k - 1 - padl))
        cf = torch.nn.functional.conv1d(x, Wc.permute(2, 1, 0)).transpose(1, 2)      # cross-correlation, like tf conv1d
        pre = pre + cf @ p['conv_dense_kernel']
    e = torch.tanh(pre) @ p['attention_v']
    e = torch.where(mask, e, torch.full_like(e, -float('inf')))
    alpha = torch.softmax(e, 1)
    ctx = torch.einsum('bt,bte->be', alpha, values)
    return alpha, ctx


@pytest.mark.parametrize('attention,B,Tm,E,H,numfilt,fs', [
    ('location_aware', 5, 13, 16, 8, 3, 5),
    ('vanilla', 4, 9, 24, 16, 0, 1),
    ('location_aware', 8, 125, 512, 256, 10, 201),             # cfg-2 width
])
def test_attention_entry_points(attention, B, Tm, E, H, numfilt, fs):
    from nabu_b200 import engine, lib as L
    lib = L.load()
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(B + Tm)
    V, NL = 7, 1
    p = O.init_speller_params(rng, V, E, H, NL, attention, max(numfilt, 1), fs)
    memory = rng.standard_normal((B, Tm, E)).astype(np.float32)
    mem_len = rng.integers(max(1, Tm // 2), Tm + 1, size=B).astype(np.int32)
    mem_len[0] = Tm
    query = rng.standard_normal((B, H)).astype(np.float32)
    prev = rng.random((B, Tm)).astype(np.float32)
    prev /= prev.sum(1, keepdims=True)
    dalpha = rng.standard_normal((B, Tm)).astype(np.float32)
    dctx = rng.standard_normal((B, E)).astype(np.float32)

    # ---- references -------------------------------------------------------------------------------------------
    values, keys, mask = O.attention_keys(memory, mem_len, p)
    ref_alpha, ref_ctx, _ = O.attention_step(query.astype(np.float64), prev.astype(np.float64), values, keys, mask, p,
                                             attention, np.float64)
    tp = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in p.items()
          if k in ('query_kernel', 'attention_v', 'conv_kernel', 'conv_dense_kernel', 'memory_kernel')}
    tq = torch.tensor(query.astype(np.float64), requires_grad=True)
    tprev = torch.tensor(prev.astype(np.float64), requires_grad=True)
    tvalues = torch.tensor(values, requires_grad=True)
    tkeys = torch.tensor(keys, requires_grad=True)
    ta, tc = _twin(tq, tprev, tvalues, tkeys, torch.tensor(mask), tp, attention)
    assert rel_err(ta.detach().numpy(), ref_alpha) < 1e-10 and rel_err(tc.detach().numpy(), ref_ctx) < 1e-10
    (ta * torch.tensor(dalpha.astype(np.float64))).sum().add((tc * torch.tensor(dctx.astype(np.float64))).sum()).backward()

    # ---- CUDA through the C ABI ----------------------------------------------------------------------------------
    desc = engine.speller_desc(B, Tm, E, V, H, NL, attention, numfilt, fs, 1)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device=dev)
    pk = L.SpellerParams()
    gk = L.SpellerParams()
    held = {}
    for name, field in (('memory_kernel', 'memory_kernel'), ('query_kernel', 'query_kernel'), ('attention_v', 'attention_v'),
                        ('conv_kernel', 'conv_kernel'), ('conv_dense_kernel', 'conv_dense_kernel'),
                        ('out_kernel', 'out_kernel'), ('out_bias', 'out_bias')):
        if name in p and p[name] is not None:
            held[name] = t(np.asarray(p[name], np.float32))
            held['g' + name] = torch.full_like(held[name], 7.0)
            setattr(pk, field, L.ptr(held[name]))
            setattr(gk, field, L.ptr(held['g' + name]))
    d_mem, d_len = t(memory), t(mem_len)
    d_values = torch.empty((B, Tm, E), device=dev)
    d_keys = torch.empty((B, Tm, H), device=dev)
    L.check(lib.nabu_attn_keys(ctypes.byref(desc), ctypes.byref(pk), L.ptr(d_mem), L.ptr(d_len), L.ptr(d_values),
                               L.ptr(d_keys), L.stream()), 'nabu_attn_keys')
    assert rel_err(d_values.cpu().numpy(), values) < 1e-6 and rel_err(d_keys.cpu().numpy(), keys) < TOL
    nws = lib.nabu_attn_workspace_bytes(ctypes.byref(desc), B)
    assert nws > 0
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    d_q, d_prev = t(query), t(prev)
    d_alpha = torch.empty((B, Tm), device=dev)
    d_ctx = torch.zeros((B, E), device=dev)
    F = numfilt if attention == 'location_aware' else 0
    q_save = torch.empty((B, H), device=dev)
    cf_save = torch.empty((B, Tm, max(F, 1)), device=dev)
    L.check(lib.nabu_attn_step_fwd(ctypes.byref(desc), ctypes.byref(pk), L.ptr(d_q), B, 1, L.ptr(d_keys), L.ptr(d_values),
                                   L.ptr(d_len), L.ptr(d_prev), L.ptr(d_alpha), L.ptr(d_ctx), L.ptr(q_save),
                                   L.ptr(cf_save), None, L.ptr(ws), nws, L.stream()), 'nabu_attn_step_fwd')
    assert rel_err(d_alpha.cpu().numpy(), ref_alpha) < TOL
    assert rel_err(d_ctx.cpu().numpy(), ref_ctx) < TOL
    assert np.all(d_alpha.cpu().numpy()[1, mem_len[1]:] == 0)

    dq = torch.empty((B, H), device=dev)
    dprev = torch.empty((B, Tm), device=dev)
    dkeys = torch.zeros((B, Tm, H), device=dev)
    dvalues = torch.zeros((B, Tm, E), device=dev)
    d_dalpha, d_dctx = t(dalpha), t(dctx)          # held: a temporary's memory would be recycled by the next allocation
    L.check(lib.nabu_attn_step_bwd(ctypes.byref(desc), ctypes.byref(pk), L.ptr(d_q), B, L.ptr(d_keys), L.ptr(d_values),
                                   L.ptr(d_len), L.ptr(d_prev), L.ptr(d_alpha), L.ptr(q_save), L.ptr(cf_save), None,
                                   L.ptr(d_dalpha), L.ptr(d_dctx), L.ptr(dq), L.ptr(dprev), L.ptr(dkeys), L.ptr(dvalues),
                                   ctypes.byref(gk), L.ptr(ws), nws, L.stream()), 'nabu_attn_step_bwd')
    torch.cuda.synchronize()
    assert rel_err(dq.cpu().numpy(), tq.grad.numpy()) < TOL
    assert rel_err(dkeys.cpu().numpy(), tkeys.grad.numpy()) < TOL
    assert rel_err(dvalues.cpu().numpy(), tvalues.grad.numpy()) < TOL
    assert rel_err(held['gquery_kernel'].cpu().numpy(), tp['query_kernel'].grad.numpy()) < TOL
    assert rel_err(held['gattention_v'].cpu().numpy(), tp['attention_v'].grad.numpy()) < TOL
    if attention == 'location_aware':
        # d(align_prev) arrives through the location features; the incoming dalign_new is the gradient wrt align_new and
        # has been consumed by the softmax backward, so what is left in the buffer is d(align_prev) alone
        assert rel_err(dprev.cpu().numpy(), tprev.grad.numpy()) < TOL
        assert rel_err(held['gconv_kernel'].cpu().numpy(), tp['conv_kernel'].grad.numpy()) < TOL
        assert rel_err(held['gconv_dense_kernel'].cpu().numpy(), tp['conv_dense_kernel'].grad.numpy()) < TOL
    else:
        assert np.all(dprev.cpu().numpy() == 0)
