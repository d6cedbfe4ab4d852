"""CPU tests that pin the oracle (it has no reference golden vectors to pin it -- PARITY UNPINNED):
independent witnesses are torch CPU ops, brute-force enumeration and closed forms."""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle as O


def test_ctc_matches_torch_and_brute_force():
    rng = np.random.default_rng(0)
    B, T, V, L = 4, 30, 6, 7
    logits = rng.normal(size=(B, T, V))
    lens = np.array([30, 25, 28, 14])
    labels = rng.integers(0, V - 1, size=(B, L))
    ll = np.array([7, 5, 6, 7])
    loss, grad = O.ctc_loss_and_grad(logits, lens, labels, ll)
    lt = torch.tensor(logits, requires_grad=True)
    tl = F.ctc_loss(F.log_softmax(lt, -1).transpose(0, 1), torch.tensor(labels), torch.tensor(lens), torch.tensor(ll),
                    blank=V - 1, reduction='none')
    tl.sum().backward()
    assert np.abs(loss - tl.detach().numpy()).max() < 1e-10
    assert np.abs(grad - lt.grad.numpy()).max() < 1e-10
    for lab in ([0, 1], [1, 1], [0], []):
        lg = rng.normal(size=(5, 3))
        l1, _ = O.ctc_loss_and_grad(lg[None], [5], [lab + [0]], [len(lab)])
        assert abs(l1[0] - O.ctc_brute_force(lg, lab)) < 1e-10


def test_ctc_infeasible_raises():
    with pytest.raises(ValueError):
        O.ctc_loss_and_grad(np.zeros((1, 3, 4)), [3], [[0, 0, 1]], [3])


def _tf_to_torch_lstm(k, b, D):
    i, j, f, o = np.split(k, 4, 1)
    bi, bj, bf, bo = np.split(b, 4)
    w = np.concatenate([i, f, j, o], 1).T
    return w[:, :D], w[:, D:], np.concatenate([bi, bf + 1, bj, bo])


def test_blstm_matches_torch_packed_lstm():
    rng = np.random.default_rng(1)
    B, T, D, H = 3, 9, 5, 4
    x = rng.normal(size=(B, T, D))
    lens = np.array([9, 6, 3])
    p = O.init_blstm_params(rng, D, H, np.float64)
    y, cache = O.blstm_fwd(x, lens, p)
    lstm = torch.nn.LSTM(D, H, batch_first=True, bidirectional=True).double()
    with torch.no_grad():
        for sfx, d in (('', 'fw'), ('_reverse', 'bw')):
            wi, wh, bb = _tf_to_torch_lstm(p[d + '_kernel'], p[d + '_bias'], D)
            getattr(lstm, 'weight_ih_l0' + sfx).copy_(torch.tensor(wi))
            getattr(lstm, 'weight_hh_l0' + sfx).copy_(torch.tensor(wh))
            getattr(lstm, 'bias_ih_l0' + sfx).copy_(torch.tensor(bb))
            getattr(lstm, 'bias_hh_l0' + sfx).zero_()
    xt = torch.tensor(x, requires_grad=True)
    pk = torch.nn.utils.rnn.pack_padded_sequence(xt, torch.tensor(lens), batch_first=True)
    out, _ = lstm(pk)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=T)
    assert np.abs(out.detach().numpy() - y).max() < 1e-12
    dy = rng.normal(size=y.shape)
    (out * torch.tensor(dy)).sum().backward()
    dx, g = O.blstm_bwd(cache, dy)
    assert np.abs(dx - xt.grad.numpy()).max() < 1e-12
    gw = lstm.weight_hh_l0.grad.numpy()
    i, f, gg, o = np.split(gw, 4, 0)
    assert np.abs(np.concatenate([i, gg, f, o], 0).T - g['fw_kernel'][D:]).max() < 1e-12
    gw = lstm.weight_ih_l0_reverse.grad.numpy()
    i, f, gg, o = np.split(gw, 4, 0)
    assert np.abs(np.concatenate([i, gg, f, o], 0).T - g['bw_kernel'][:D]).max() < 1e-12


def test_pyramid_stack():
    x = np.arange(2 * 5 * 3, dtype=np.float64).reshape(2, 5, 3)
    y, l = O.pyramid_stack_fwd(x, [5, 3], 2)
    assert y.shape == (2, 3, 6) and list(l) == [3, 2]
    assert np.array_equal(y[0, 0], np.concatenate([x[0, 0], x[0, 1]]))
    assert np.array_equal(y[0, 2], np.concatenate([x[0, 4], np.zeros(3)]))
    assert np.array_equal(O.pyramid_stack_bwd(y, 5, 2), x)


def _torch_speller(memory, mem_lens, targets, tl, p, attention, NL, probability_fn='softmax', window=None, ids=None,
                   drop=None):
    """Independent torch-autograd twin of the Speller forward (used only to check the oracle)."""
    B, Tm, E = memory.shape
    V = p['out_bias'].shape[0]
    H = p['query_kernel'].shape[0]
    U = int(tl.max())
    mask = torch.arange(Tm)[None, :] < torch.tensor(mem_lens)[:, None]
    values = memory * mask[:, :, None]
    keys = values @ p['memory_kernel']
    if ids is None:
        ids = torch.cat([torch.full((B, 1), V - 1, dtype=torch.long), torch.tensor(targets, dtype=torch.long)[:, :U]], 1)
    h = [torch.zeros(B, H, dtype=torch.float64) for _ in range(NL)]
    c = [torch.zeros(B, H, dtype=torch.float64) for _ in range(NL)]
    att = torch.zeros(B, E, dtype=torch.float64)
    al = torch.zeros(B, Tm, dtype=torch.float64)
    if attention == 'windowed':
        al[:, 0] = 1
    outs = []
    for u in range(U):
        act = (u < torch.tensor(tl))[:, None]
        inp = torch.cat([F.one_hot(ids[:, u], V).double(), att], 1)
        nh, nc = [], []
        for l in range(NL):
            z = torch.cat([inp, h[l]], 1) @ p['cell_%d_kernel' % l] + p['cell_%d_bias' % l]
            i, j, f, o = z.chunk(4, 1)
            cn = c[l] * torch.sigmoid(f + 1) + torch.sigmoid(i) * torch.tanh(j)
            hn = torch.tanh(cn) * torch.sigmoid(o)
            nh.append(hn)
            nc.append(cn)
            inp = hn if drop is None else hn * drop[1][u][l] / drop[0]      # output dropout; the state keeps hn
        pre = (inp @ p['query_kernel'])[:, None, :] + keys
        if attention == 'location_aware':
            ksz = p['conv_kernel'].shape[0]
            padl = (ksz - 1) // 2
            padded = F.pad(al[:, None, :], (padl, ksz - 1 - padl))
            cf = F.conv1d(padded, p['conv_kernel'].permute(2, 1, 0))            # [B,F,Tm]
            pre = pre + cf.transpose(1, 2) @ p['conv_dense_kernel']
        e = torch.tanh(pre) @ p['attention_v']
        if attention == 'windowed':                  # written independently of the oracle: explicit index arithmetic
            L, R = window
            half = (torch.cumsum(al.detach(), 1) > 0.5)
            win = torch.zeros_like(half)
            for t in range(Tm):
                left = half[:, t + L + 1] if t + L + 1 < Tm else torch.ones(B, dtype=torch.bool)
                right = half[:, t - R] if t - R >= 0 else torch.zeros(B, dtype=torch.bool)
                win[:, t] = left ^ right
            e = e.masked_fill(~win, float('-inf'))
        e = e.masked_fill(~mask, float('-inf'))
        if probability_fn == 'softmax':
            a = torch.softmax(e, 1)
        else:
            a = torch.sigmoid(e)                     # sigmoid(-inf) = 0 on the masked positions
            if probability_fn == 'normalized_sigmoid':
                a = a / a.sum(1, keepdim=True)
        ctx = torch.einsum('bt,bte->be', a, values)
        lg = torch.cat([inp, ctx], 1) @ p['out_kernel'] + p['out_bias']
        outs.append(torch.where(act, lg, torch.zeros_like(lg)))
        h = [torch.where(act, a_, b_) for a_, b_ in zip(nh, h)]
        c = [torch.where(act, a_, b_) for a_, b_ in zip(nc, c)]
        att = torch.where(act, ctx, att)
        al = torch.where(act, a, al)
    return torch.stack(outs, 1)


@pytest.mark.parametrize('probability_fn', ['softmax', 'sigmoid', 'normalized_sigmoid'])
@pytest.mark.parametrize('attention', ['vanilla', 'location_aware', 'windowed'])
def test_speller_oracle_matches_torch_autograd(attention, probability_fn):
    rng = np.random.default_rng(2)
    B, Tm, E, V, H, NL, U = 4, 9, 6, 5, 4, 2, 5
    p = O.init_speller_params(rng, V, E, H, NL, attention, 3, 4, np.float64)
    for k in p:
        if k.endswith('bias'):
            p[k] = rng.normal(size=p[k].shape) * 0.1
    memory = rng.normal(size=(B, Tm, E))
    mem_lens = np.array([9, 7, 5, 9])
    tl = np.array([5, 3, 1, 4])
    targets = rng.integers(0, V, size=(B, U))
    dlog = rng.normal(size=(B, U, V))
    for b in range(B):
        dlog[b, tl[b]:] = 0
    window = (2, 3) if attention == 'windowed' else None
    logits, ctx = O.speller_fwd(memory, mem_lens, targets, tl, p, attention, NL, probability_fn=probability_fn,
                                window=window)
    dmem, g = O.speller_bwd(ctx, dlog)
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    tm = torch.tensor(memory, requires_grad=True)
    tlog = _torch_speller(tm, mem_lens, targets, tl, tp, attention, NL, probability_fn, window)
    assert np.abs(tlog.detach().numpy() - logits).max() < 1e-12
    (tlog * torch.tensor(dlog)).sum().backward()
    assert np.abs(tm.grad.numpy() - dmem).max() < 1e-10
    for k in g:
        if tp[k].grad is not None:
            assert np.abs(tp[k].grad.numpy() - g[k]).max() < 1e-10, k


def test_average_cross_entropy_matches_torch():
    rng = np.random.default_rng(3)
    B, U, V = 5, 7, 6
    logits = rng.normal(size=(B, U, V))
    tl = np.array([7, 3, 1, 5, 6])
    tg = rng.integers(0, V, size=(B, U))
    loss, d = O.average_cross_entropy(logits, tg, tl, tl)
    lt = torch.tensor(logits, requires_grad=True)
    ce = F.cross_entropy(lt.reshape(-1, V), torch.tensor(tg).reshape(-1), reduction='none').reshape(B, U)
    m = (torch.arange(U)[None, :] < torch.tensor(tl)[:, None]).double()
    ref = ((ce * m).sum(1) / torch.tensor(tl).double()).mean()
    ref.backward()
    assert abs(loss - ref.item()) < 1e-12 and np.abs(d - lt.grad.numpy()).max() < 1e-12


def test_tf_adam_closed_form():
    th = np.array([1.0, -2.0, 0.5])
    g = np.array([0.3, -5.0, 0.0])
    th1, m1, v1 = O.tf_adam_clip(th, g, np.zeros(3), np.zeros(3), 1e-3, 1, dtype=np.float64)
    gc = np.clip(g, -1, 1)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(th1, th - lr_t * (0.1 * gc) / (np.sqrt(0.001 * gc * gc) + 1e-8), rtol=1e-12)
    assert abs(O.exponential_decay(1e-3, 50, 100, 0.1) - 1e-3 * 0.1 ** 0.5) < 1e-15


def test_ctc_beam_search_against_exhaustive_prefix_search():
    """Beam wide enough to be exhaustive: the best path must be the arg-max over all label sequences of
    the total (unnormalised, max-shifted like TF's Step) path mass, computed by brute force."""
    rng = np.random.default_rng(5)
    T, V = 5, 4
    for trial in range(5):
        logits = rng.normal(size=(T, V)).astype(np.float32) * 2
        path, neg = O.ctc_beam_search(logits, T, beam_width=1000, merge_repeated=False)
        x = logits - logits.max(1, keepdims=True)
        mass = {}
        for al in itertools.product(range(V), repeat=T):
            lab, prev = [], None
            for s in al:
                if s != prev and s != V - 1:
                    lab.append(s)
                prev = s
            mass[tuple(lab)] = np.logaddexp(mass.get(tuple(lab), -np.inf), x[np.arange(T), list(al)].sum())
        best_lab = max(mass, key=mass.get)
        assert tuple(path) == best_lab
        assert abs(-neg - mass[best_lab]) < 1e-3


def test_ctc_beam_search_merge_repeated_collapses():
    logits = np.full((4, 3), -5.0, np.float32)
    logits[0, 0] = logits[2, 0] = 5.0      # a, blank, a, blank -> "a a"
    logits[1, 2] = logits[3, 2] = 5.0
    p_keep, _ = O.ctc_beam_search(logits, 4, 10, merge_repeated=False)
    p_merge, _ = O.ctc_beam_search(logits, 4, 10, merge_repeated=True)
    assert list(p_keep) == [0, 0] and list(p_merge) == [0]


def test_las_beam_search_wide_beam_beats_greedy():
    rng = np.random.default_rng(6)
    B, Tm, E, V, H, NL = 2, 6, 4, 4, 4, 1
    p = O.init_speller_params(rng, V, E, H, NL, 'vanilla', 1, 1)
    p['out_bias'] = np.array([0, 0, 0, 1.5], np.float32)
    memory = rng.normal(size=(B, Tm, E)).astype(np.float32)
    lens = np.array([6, 4])
    seqs, lengths, scores, aligns = O.las_beam_search(memory, lens, p, 4, 8, 'vanilla', NL, 0.0, 1.0)
    assert seqs.shape[0] == B and seqs.shape[1] == 4
    assert np.all(scores[:, 0] >= scores[:, 1])
    values, keys, mask = O.attention_keys(memory, lens, p, np.float32)
    st = O.speller_zero_state(B, Tm, E, H, NL, np.float32)
    ids = np.full(B, V - 1)
    lp = np.zeros(B)
    done = np.zeros(B, bool)
    for _ in range(8):
        lg, st, _ = O.speller_step(ids, st, values, keys, mask, p, 'vanilla', np.float32)
        l = O.log_softmax(lg.astype(np.float32))
        ids = l.argmax(1)
        lp = np.where(done, lp, lp + l.max(1))
        done |= ids == V - 1
    assert np.all(scores[:, 0] >= lp - 1e-4)
    assert aligns.shape == (B, 4, seqs.shape[2], Tm)


def test_edit_distance():
    assert O.edit_distance([1, 2, 3], [1, 3]) == 1 and O.edit_distance([], [1, 2]) == 2


def test_speller_dropout_and_scheduled_sampling_oracle_matches_torch():
    """Rows a6/a7 with the stochastic parts on: DropoutWrapper(output_keep_prob) on every LSTM layer and
    ScheduledEmbeddingTrainingHelper.  The masks and draws come from the counter-based generator the kernels use, so
    the oracle is deterministic; the torch twin is fed the same masks and the token ids the oracle ended up feeding."""
    rng = np.random.default_rng(5)
    B, Tm, E, V, H, NL, U = 5, 8, 6, 6, 4, 2, 6
    keep, sp, seed = 0.6, 0.5, 1234
    p = O.init_speller_params(rng, V, E, H, NL, 'location_aware', 3, 4, np.float64)
    memory = rng.normal(size=(B, Tm, E))
    mem_lens = np.array([8, 6, 5, 8, 7])
    tl = np.array([6, 3, 1, 5, 6])
    targets = rng.integers(0, V, size=(B, U))
    dlog = rng.normal(size=(B, U, V))
    for b in range(B):
        dlog[b, tl[b]:] = 0
    logits, ctx = O.speller_fwd(memory, mem_lens, targets, tl, p, 'location_aware', NL, dropout_keep=keep,
                                sample_prob=sp, seed=seed)
    teacher = np.concatenate([np.full((B, 1), V - 1), targets[:, :U]], 1)
    assert (ctx['ids_in'][:, :U] != teacher[:, :U]).any()                    # some tokens were sampled
    dmem, g = O.speller_bwd(ctx, dlog)
    masks = [[torch.tensor(O.speller_dropout_mask(seed, l, u, B, H, keep)) for l in range(NL)] for u in range(U)]
    frac = np.mean([m.numpy().mean() for ms in masks for m in ms])
    assert abs(frac - keep) < 0.15
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    tm = torch.tensor(memory, requires_grad=True)
    tlog = _torch_speller(tm, mem_lens, targets, tl, tp, 'location_aware', NL, ids=torch.tensor(ctx['ids_in']),
                          drop=(keep, masks))
    assert np.abs(tlog.detach().numpy() - logits).max() < 1e-12
    (tlog * torch.tensor(dlog)).sum().backward()
    assert np.abs(tm.grad.numpy() - dmem).max() < 1e-10
    for k in g:
        if tp[k].grad is not None:
            assert np.abs(tp[k].grad.numpy() - g[k]).max() < 1e-10, k


def test_counter_rng_is_uniform_and_sampling_follows_the_softmax():
    u = O.rng_uniform(7, 3, np.arange(200000), 11)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3 and abs((u < 0.25).mean() - 0.25) < 5e-3
    hist, _ = np.histogram(u, bins=16, range=(0, 1))
    assert np.abs(hist / len(u) - 1 / 16).max() < 3e-3
    logits = np.tile(np.log(np.array([[0.5, 0.25, 0.125, 0.125]])), (100000, 1))
    ids = O.speller_sample_ids(99, 0, logits, 1.0)
    freq = np.bincount(ids, minlength=4) / len(ids)
    assert np.abs(freq - [0.5, 0.25, 0.125, 0.125]).max() < 5e-3
    assert (O.speller_sample_ids(99, 0, logits, 0.0) == -1).all()
    assert abs((O.speller_sample_ids(99, 1, logits, 0.3) >= 0).mean() - 0.3) < 5e-3
