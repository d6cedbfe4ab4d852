"""Golden-fixture tests.  tests/golden/*.npz are frozen outputs of independent witnesses (torch CPU
ops, brute force; see tests/golden/make_golden.py) -- the reference ships none.  CPU: the oracle must
reproduce them.  GPU: the CUDA kernels must reproduce them too (1e-4 relative; ids exact)."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from tests.util import rel_err

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return dict(np.load(os.path.join(G, name)))


def test_oracle_ctc_golden():
    g = load('ctc.npz')
    loss, grad = O.ctc_loss_and_grad(g['logits'], g['lens'], g['labels'], g['ll'])
    assert rel_err(loss, g['loss']) < 1e-12 and np.abs(grad - g['grad']).max() < 1e-12


def test_oracle_blstm_golden():
    g = load('blstm.npz')
    p = {k: g[k] for k in ('fw_kernel', 'fw_bias', 'bw_kernel', 'bw_bias')}
    y, cache = O.blstm_fwd(g['x'], g['lens'], p)
    dx, gr = O.blstm_bwd(cache, g['dy'].astype(np.float64))
    assert rel_err(y, g['y']) < 1e-6 and rel_err(dx, g['dx']) < 1e-6
    for k, r in (('fw_kernel', 'dkf'), ('fw_bias', 'dbf'), ('bw_kernel', 'dkb'), ('bw_bias', 'dbb')):
        assert rel_err(gr[k], g[r]) < 1e-6, k


def _speller_params(g):
    return {k[2:]: g[k] for k in g if k.startswith('p_')}


def test_oracle_speller_golden():
    g = load('speller.npz')
    p = _speller_params(g)
    logits, ctx = O.speller_fwd(g['memory'], g['mem_lens'], g['targets'], g['tl'], p, 'location_aware', 2)
    dmem, gr = O.speller_bwd(ctx, g['dlog'].astype(np.float64))
    assert np.abs(logits - g['logits']).max() < 1e-12 and np.abs(dmem - g['dmemory']).max() < 1e-10
    for k in gr:
        assert np.abs(gr[k] - g['g_' + k]).max() < 1e-10, k


def test_oracle_ctc_beam_golden():
    g = load('ctc_beam.npz')
    for n in range(g['logits'].shape[0]):
        path, _ = O.ctc_beam_search(g['logits'][n], g['logits'].shape[1], 1000, merge_repeated=False)
        assert list(path) == list(g['best'][n, :g['best_len'][n]])


# ------------------------------------------------------------------------------------------------
# the same fixtures through the CUDA kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_ctc_golden():
    from nabu_b200 import engine
    g = load('ctc.npz')
    args = [torch.tensor(g[k]).cuda() for k in ('logits', 'lens', 'labels', 'll')]
    loss, grad = engine.ctc_loss_per_utt(*args, want_grad=True)
    assert rel_err(loss.cpu().numpy(), g['loss']) < 1e-5
    assert np.abs(grad.cpu().numpy() - g['grad']).max() < 1e-5


@pytest.mark.gpu
def test_cuda_blstm_golden():
    from nabu_b200 import lib as L
    g = load('blstm.npz')
    lib = L.load()
    B, T, D = g['x'].shape
    H = g['fw_bias'].shape[0] // 4
    d = {k: torch.tensor(g[k]).cuda() for k in ('x', 'lens', 'dy', 'fw_kernel', 'fw_bias', 'bw_kernel', 'bw_bias')}
    y = torch.empty((B, T, 2 * H), device='cuda')
    gates = torch.empty((2, B, T, 4 * H), device='cuda')
    cells = torch.empty((2, B, T, H), device='cuda')
    nws = lib.nabu_blstm_workspace_bytes(B, T, D, H)
    ws = torch.empty(nws, dtype=torch.uint8, device='cuda')
    L.check(lib.nabu_blstm_fwd(L.ptr(d['x']), L.ptr(d['lens']), B, T, D, H, L.ptr(d['fw_kernel']), L.ptr(d['fw_bias']),
                               L.ptr(d['bw_kernel']), L.ptr(d['bw_bias']), L.ptr(y), T, L.ptr(gates), L.ptr(cells),
                               L.ptr(ws), nws, L.stream()), 'fwd')
    assert rel_err(y.cpu().numpy(), g['y']) < 1e-4
    dx = torch.empty_like(d['x'])
    gk = {k: torch.empty_like(d[k]) for k in ('fw_kernel', 'fw_bias', 'bw_kernel', 'bw_bias')}
    L.check(lib.nabu_blstm_bwd(L.ptr(d['x']), L.ptr(d['lens']), B, T, D, H, L.ptr(d['fw_kernel']),
                               L.ptr(d['bw_kernel']), L.ptr(y), T, L.ptr(gates), L.ptr(cells), L.ptr(d['dy']),
                               L.ptr(dx), L.ptr(gk['fw_kernel']), L.ptr(gk['fw_bias']), L.ptr(gk['bw_kernel']),
                               L.ptr(gk['bw_bias']), L.ptr(ws), nws, L.stream()), 'bwd')
    assert rel_err(dx.cpu().numpy(), g['dx']) < 1e-4
    for k, r in (('fw_kernel', 'dkf'), ('fw_bias', 'dbf'), ('bw_kernel', 'dkb'), ('bw_bias', 'dbb')):
        assert rel_err(gk[k].cpu().numpy(), g[r]) < 1e-4, k


@pytest.mark.gpu
def test_cuda_speller_golden():
    from nabu_b200 import engine
    from tests.test_gpu_speller import _svars, _grads
    g = load('speller.npz')
    p = _speller_params(g)
    dev = torch.device('cuda', 0)
    sv = _svars(p, 'location_aware', 2, dev)
    mem = torch.tensor(g['memory'], device=dev, requires_grad=True)
    V, H = p['out_bias'].shape[0], p['query_kernel'].shape[0]
    logits = engine.speller(mem, torch.tensor(g['mem_lens'], device=dev), torch.tensor(g['targets'], device=dev),
                            torch.tensor(g['tl'], device=dev), sv, V, H, 2, 'location_aware', 3, 5)
    assert rel_err(logits.detach().cpu().numpy(), g['logits']) < 1e-4
    logits.backward(torch.tensor(g['dlog'], device=dev))
    assert rel_err(mem.grad.cpu().numpy(), g['dmemory']) < 1e-4
    for k, v in _grads(sv, 2, 'location_aware').items():
        assert rel_err(v, g['g_' + k]) < 1e-4, k


@pytest.mark.gpu
def test_cuda_ctc_beam_golden():
    from nabu_b200 import engine
    g = load('ctc_beam.npz')
    N, T, V = g['logits'].shape
    ids, lens, _ = engine.ctc_beam_search(torch.tensor(g['logits']).cuda(),
                                          torch.full((N,), T, dtype=torch.int32).cuda(), 1000, False)
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    for n in range(N):
        assert lens[n] == g['best_len'][n] and np.array_equal(ids[n, :lens[n]], g['best'][n, :lens[n]])
