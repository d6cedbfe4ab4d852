"""Parity at BASELINE.json's full sizes (cfg-3: DBLSTM 5x512 + CTC, 128 x 1500 x 40) through size-independent properties
-- the NumPy oracle would take minutes there -- plus one independent witness that is fast enough (torch's CPU CTC).

 * batch-permutation equivariance: utterances are independent through the encoder, so permuting the batch permutes
   outputs and input gradients BIT-EXACTLY and leaves the (summed) weight gradients unchanged up to summation order;
 * masking: outputs are exactly zero past every utterance's length, and so are dx and the CTC gradient;
 * padding idempotence: the same utterances in another padded batch give the same outputs on the common part (bit for bit
   when both batches select the same kernel variant, to fp32 rounding across the 128-row and the small-batch kernels);
 * CTC: loss and gradient at 128 x 1500 x 29 against torch.nn.functional.ctc_loss (CPU, fp64, blank = V-1), gradient
   rows sum to zero over the labels;
 * determinism: two identical train steps from the same state give the bit-identical loss and parameters.
"""
import numpy as np
import pytest
import torch

import oracle as O
from tests.util import make_conf, synthetic_ctc_batch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _blstm_fwd_bwd(x, lens, p, dy, yT=None):
    from nabu_b200 import lib as L
    lib = L.load()
    B, T, D = x.shape
    H = p['fw_bias'].shape[0] // 4
    yT = T if yT is None else yT
    xd, ld = dev(x), dev(lens)
    pd = {k: dev(v) for k, v in p.items()}
    y = torch.empty((B, yT, 2 * H), device='cuda')
    gates = torch.empty((2, B, T, 4 * H), device='cuda')
    cells = torch.empty((2, B, T, H), device='cuda')
    nws = lib.nabu_blstm_workspace_bytes(B, T, D, H)
    ws = torch.empty(nws, dtype=torch.uint8, device='cuda')
    L.check(lib.nabu_blstm_fwd(L.ptr(xd), L.ptr(ld), B, T, D, H, L.ptr(pd['fw_kernel']), L.ptr(pd['fw_bias']),
                               L.ptr(pd['bw_kernel']), L.ptr(pd['bw_bias']), L.ptr(y), yT, L.ptr(gates), L.ptr(cells),
                               L.ptr(ws), nws, L.stream()), 'fwd')
    out = {'y': y.cpu().numpy()}
    if dy is not None:
        dyd = dev(dy)
        dx = torch.empty_like(xd)
        gk = {k: torch.empty_like(v) for k, v in pd.items()}
        L.check(lib.nabu_blstm_bwd(L.ptr(xd), L.ptr(ld), B, T, D, H, L.ptr(pd['fw_kernel']), L.ptr(pd['bw_kernel']),
                                   L.ptr(y), yT, L.ptr(gates), L.ptr(cells), L.ptr(dyd), L.ptr(dx),
                                   L.ptr(gk['fw_kernel']), L.ptr(gk['fw_bias']), L.ptr(gk['bw_kernel']),
                                   L.ptr(gk['bw_bias']), L.ptr(ws), nws, L.stream()), 'bwd')
        torch.cuda.synchronize()
        out['dx'] = dx.cpu().numpy()
        out.update({'d' + k: v.cpu().numpy() for k, v in gk.items()})
    return out


def test_blstm_full_size_permutation_masking_padding():
    """One cfg-3 layer (B=128, T=1500, D=1024 -> H=512) forward and backward."""
    B, T, D, H = 128, 1500, 1024, 512
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((B, T, D)) * 0.5).astype(np.float32)
    lens = rng.integers(int(0.6 * T), T + 1, size=B).astype(np.int32)
    lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = 0
    p = O.init_blstm_params(rng, D, H)
    dy = rng.standard_normal((B, T, 2 * H)).astype(np.float32)
    for b in range(B):
        dy[b, lens[b]:] = 0
    a = _blstm_fwd_bwd(x, lens, p, dy)
    # masking
    for b in range(0, B, 9):
        assert np.all(a['y'][b, lens[b]:] == 0) and np.all(a['dx'][b, lens[b]:] == 0)
    assert np.all(np.isfinite(a['y'])) and np.abs(a['y']).max() < 1.0
    # permutation equivariance (bit-exact per utterance; weight gradients are sums over utterances)
    perm = rng.permutation(B)
    b2 = _blstm_fwd_bwd(x[perm], lens[perm], p, dy[perm])
    assert np.array_equal(b2['y'], a['y'][perm])
    assert np.array_equal(b2['dx'], a['dx'][perm])
    for k in ('dfw_kernel', 'dbw_kernel', 'dfw_bias', 'dbw_bias'):
        scale = np.abs(a[k]).max()
        # 192 000-term fp32 sums in a different order: sqrt(N) * 2^-24 * |partial sums| ~ 1e-4 of the largest entry
        assert np.abs(b2[k] - a[k]).max() <= 2e-4 * scale, k
    # padding idempotence: first 16 utterances alone, padded to a longer T (forward only; yT > T exercises the pad rows)
    sub = np.argsort(lens)[:16]
    Tsub = int(lens[sub].max())
    c = _blstm_fwd_bwd(np.ascontiguousarray(x[sub, :Tsub]), lens[sub], p, None, yT=Tsub + 4)
    # (a batch of 16 runs the small-batch chain kernels, the batch of 128 the 128-row kernel: the same arithmetic in
    # another summation order inside the tensor core, so the outputs agree to fp32 rounding, not bit for bit)
    assert np.abs(c['y'][:, :Tsub] - a['y'][sub, :Tsub]).max() < 2e-6
    assert np.all(c['y'][:, Tsub:] == 0)
    # ... and bit for bit when both batches run the same kernel (two sub-batches of 16)
    sub2 = np.argsort(lens)[8:24]
    T2 = int(lens[sub2].max())
    d = _blstm_fwd_bwd(np.ascontiguousarray(x[sub2, :T2]), lens[sub2], p, None)
    common = [i for i in sub2 if i in set(sub.tolist())]
    for i in common:
        ia, ib = list(sub).index(i), list(sub2).index(i)
        assert np.array_equal(c['y'][ia, :lens[i]], d['y'][ib, :lens[i]])


def test_ctc_full_size_against_torch_cpu():
    from nabu_b200 import engine
    B, T, V, L = 128, 1500, 29, 150
    rng = np.random.default_rng(3)
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    lens = rng.integers(int(0.6 * T), T + 1, size=B).astype(np.int32)
    labels = rng.integers(0, V - 1, size=(B, L)).astype(np.int32)
    ll = np.maximum(lens // 10, 1).astype(np.int32)
    loss, grad = engine.ctc_loss_per_utt(dev(logits), dev(lens), dev(labels), dev(ll), want_grad=True)
    loss, grad = loss.cpu().numpy(), grad.cpu().numpy()
    lt = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    lp = torch.log_softmax(lt, dim=-1).transpose(0, 1)
    ref = torch.nn.functional.ctc_loss(lp, torch.tensor(labels, dtype=torch.long), torch.tensor(lens, dtype=torch.long),
                                       torch.tensor(ll, dtype=torch.long), blank=V - 1, reduction='none')
    ref.sum().backward()
    assert np.abs(loss - ref.detach().numpy()).max() / np.abs(ref.detach().numpy()).max() < 1e-5
    assert np.abs(grad - lt.grad.numpy()).max() < 1e-4
    for b in range(0, B, 7):
        assert np.all(grad[b, lens[b]:] == 0)
        assert np.abs(grad[b, :lens[b]].sum(axis=-1)).max() < 1e-4      # softmax minus posteriors: rows sum to zero


def test_cfg3_train_step_is_deterministic():
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    dev_ = torch.device('cuda', 0)
    B, T, D, H, NL, V = 128, 300, 40, 512, 5, 29
    mconf = ('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = dblstm\nnum_units = %d\n'
             'num_layers = %d\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n' % (V - 1, H, NL))
    tconf = '[trainer]\ntrainer = standard\nloss = CTC\ntrainlabels = 1\ntargets = text\n'
    x, lens, labels, ll = synthetic_ctc_batch(B, T, D, V, ragged=True)
    batch = ({'features': dev(x)}, {'features': dev(lens)}, {'text': dev(labels)}, {'text': dev(ll)})
    runs = []
    for _ in range(2):
        tr = trainer_factory.factory('standard')(make_conf(tconf), None, make_conf(mconf), None, None, None, 0,
                                                 device=dev_, seed=9)
        tr.num_steps = 100
        tr.model.build({'features': D}, dev_)
        losses = [float(tr.update(*batch)[0]) for _ in range(2)]
        torch.cuda.synchronize()
        runs.append((losses, tr.model.store.to_numpy()))
    assert runs[0][0] == runs[1][0] and np.isfinite(runs[0][0]).all()
    for k in runs[0][1]:
        assert np.array_equal(runs[0][1][k], runs[1][1][k]), k
