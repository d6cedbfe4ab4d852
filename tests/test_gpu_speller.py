"""GPU parity tests for the attention decoder (rows a6-a8, a10, a13): nabu_speller_fwd/bwd and
nabu_las_beam_search against the NumPy oracle, plus the LAS model through the Trainer."""
import numpy as np
import pytest
import torch

import oracle as O
from tests.util import make_conf, rel_err, synthetic_las_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4      # BASELINE.json north_star: 1e-4 relative fp32


def _svars(p, attention, NL, dev):
    """engine.SpellerVars over plain tensors (with grad buffers) from an oracle parameter dict."""
    from nabu_b200 import engine

    class V(object):
        def __init__(self, a):
            self.data = torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev).requires_grad_(True)
            self.grad = torch.full_like(self.data, 5.0).detach()
    ks = [V(p['cell_%d_kernel' % l]) for l in range(NL)]
    bs = [V(p['cell_%d_bias' % l]) for l in range(NL)]
    ck = V(p['conv_kernel']) if attention == 'location_aware' else None
    dk = V(p['conv_dense_kernel']) if attention == 'location_aware' else None
    return engine.SpellerVars(ks, bs, V(p['memory_kernel']), V(p['query_kernel']), V(p['attention_v']), ck, dk,
                              V(p['out_kernel']), V(p['out_bias']))


def _grads(sv, NL, attention):
    g = {'memory_kernel': sv.memory_kernel, 'query_kernel': sv.query_kernel, 'attention_v': sv.attention_v,
         'out_kernel': sv.out_kernel, 'out_bias': sv.out_bias}
    for l in range(NL):
        g['cell_%d_kernel' % l] = sv.cell_kernels[l]
        g['cell_%d_bias' % l] = sv.cell_biases[l]
    if attention == 'location_aware':
        g['conv_kernel'] = sv.conv_kernel
        g['conv_dense_kernel'] = sv.conv_dense_kernel
    return {k: v.grad.cpu().numpy() for k, v in g.items()}


@pytest.mark.parametrize('attention,B,Tm,E,V,H,NL,U,numfilt,fs', [
    ('location_aware', 5, 13, 16, 7, 8, 2, 6, 3, 5),
    ('vanilla', 4, 9, 24, 9, 16, 2, 5, 0, 1),
    ('location_aware', 3, 20, 32, 30, 32, 1, 7, 10, 21),
    ('location_aware', 70, 11, 16, 12, 24, 3, 4, 4, 4),     # > 64 rows (two row tiles), even filter
    ('location_aware+normalized_sigmoid', 6, 15, 16, 8, 16, 2, 6, 3, 5),   # row f4: attention.py:41-55
    ('location_aware+sigmoid', 6, 15, 16, 8, 16, 2, 6, 3, 5),
    ('vanilla+normalized_sigmoid', 4, 9, 24, 9, 16, 2, 5, 0, 1),
    ('vanilla+sigmoid', 4, 9, 24, 9, 16, 1, 5, 0, 1),
    ('windowed', 6, 17, 16, 8, 16, 2, 7, 2, 3),             # row f4: attention.py:294-396, left = 2, right = 3
    ('windowed+normalized_sigmoid', 5, 23, 16, 8, 16, 2, 6, 4, 5),
])
def test_speller_fwd_bwd(attention, B, Tm, E, V, H, NL, U, numfilt, fs):
    from nabu_b200 import engine
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(B * 100 + U)
    full_attention = attention
    attention, _, prob_fn = attention.partition('+')
    prob_fn = prob_fn or 'softmax'
    window = (numfilt, fs) if attention == 'windowed' else None
    p = O.init_speller_params(rng, V, E, H, NL, attention, max(numfilt, 1), fs)
    for k in p:
        if k.endswith('bias'):
            p[k] = (rng.standard_normal(p[k].shape) * 0.1).astype(np.float32)
    memory = rng.standard_normal((B, Tm, E)).astype(np.float32)
    mem_len = rng.integers(max(1, Tm // 2), Tm + 1, size=B).astype(np.int32)
    mem_len[0] = Tm
    tl = rng.integers(1, U + 1, size=B).astype(np.int32)
    tl[0] = U
    targets = rng.integers(0, V, size=(B, U)).astype(np.int32)
    dlog = rng.standard_normal((B, U, V)).astype(np.float32)
    for b in range(B):
        dlog[b, tl[b]:] = 0
    ref_logits, ctx = O.speller_fwd(memory, mem_len, targets, tl, p, attention, NL, np.float64, probability_fn=prob_fn,
                                    window=window)
    ref_dmem, ref_g = O.speller_bwd(ctx, dlog.astype(np.float64))

    sv = _svars(p, attention, NL, dev)
    mem_d = torch.tensor(memory, device=dev, requires_grad=True)
    logits = engine.speller(mem_d, torch.tensor(mem_len, device=dev), torch.tensor(targets, device=dev),
                            torch.tensor(tl, device=dev), sv, V, H, NL, full_attention, numfilt, fs)
    assert rel_err(logits.detach().cpu().numpy(), ref_logits) < TOL
    logits.backward(torch.tensor(dlog, device=dev))
    assert rel_err(mem_d.grad.cpu().numpy(), ref_dmem) < TOL
    got = _grads(sv, NL, attention)
    for k, v in got.items():
        assert rel_err(v, ref_g[k]) < TOL, k


@pytest.mark.parametrize('attention,B,W,Tm,E,V,H,NL,max_steps,lp', [
    ('location_aware', 3, 4, 12, 16, 7, 8, 2, 9, 1.0),
    ('vanilla', 2, 3, 8, 8, 5, 8, 1, 6, 0.0),
    ('location_aware', 4, 16, 25, 32, 30, 32, 2, 20, 1.0),   # beam 16 like the LAS recipe
    ('location_aware+normalized_sigmoid', 3, 4, 12, 16, 7, 8, 2, 9, 1.0),
    ('vanilla+sigmoid', 2, 3, 8, 8, 5, 8, 1, 6, 0.0),
    ('windowed', 3, 4, 14, 16, 7, 8, 2, 9, 1.0),
])
def test_las_beam_search_ids_bit_exact(attention, B, W, Tm, E, V, H, NL, max_steps, lp):
    from nabu_b200 import engine
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(B * 10 + W)
    full_attention = attention
    attention, _, prob_fn = attention.partition('+')
    numfilt, fs = (3, 5) if attention == 'location_aware' else ((2, 3) if attention == 'windowed' else (0, 1))
    window = (numfilt, fs) if attention == 'windowed' else None
    p = O.init_speller_params(rng, V, E, H, NL, attention, max(numfilt, 1), fs)
    p['out_bias'] = rng.standard_normal(V).astype(np.float32)
    p['out_bias'][V - 1] += 1.0       # make EOS likely enough that hypotheses finish
    memory = rng.standard_normal((B, Tm, E)).astype(np.float32)
    mem_len = rng.integers(max(1, Tm // 2), Tm + 1, size=B).astype(np.int32)
    mem_len[0] = Tm
    ref = O.las_beam_search(memory, mem_len, p, W, max_steps, attention, NL, lp, 1.0, np.float32,
                            probability_fn=prob_fn or 'softmax', window=window)
    sv = _svars(p, attention, NL, dev)
    got = engine.las_beam_search(torch.tensor(memory, device=dev), torch.tensor(mem_len, device=dev), sv, V, H, NL,
                                 full_attention, numfilt, fs, W, max_steps, lp, 1.0)
    seqs, lens, scores, aligns = [g.cpu().numpy() for g in got]
    assert seqs.shape == ref[0].shape, (seqs.shape, ref[0].shape)      # same number of loop iterations
    assert np.array_equal(seqs, ref[0])                                 # token ids: bit-exact
    assert np.array_equal(lens, ref[1])
    fin = np.isfinite(ref[2])
    assert np.array_equal(np.isfinite(scores), fin)
    assert rel_err(scores[fin], ref[2][fin]) < TOL
    assert np.abs(aligns - ref[3]).max() < 1e-4


def _las_trainer(dev, H, NL, V, dec_H, numfilt, fs):
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = listener\n'
                      'num_units = %d\nnum_layers = %d\npyramid_steps = 2\ninput_noise = 0\ndropout = 1\n[decoder]\n'
                      'decoder = speller\nnum_layers = 2\nnum_units = %d\ndropout = 1\nattention = location_aware\n'
                      'numfilt = %d\nfiltersize = %d\nsample_prob = 0\n' % (V - 1, H, NL, dec_H, numfilt, fs))
    tconf = make_conf('[trainer]\ntrainer = standard\nloss = average_cross_entropy\ntrainlabels = 1\ntargets = text\n')
    tr = trainer_factory.factory('standard')(tconf, None, mconf, None, None, None, 0, device=dev, seed=9)
    tr.num_steps = 100
    return tr


def las_oracle_params(params, NL, inp='features'):
    layers = []
    for l in range(NL + 1):
        mid = 'BLSTM/' if l < NL else ''
        base = 'Listener/%s/layer%d/%sbidirectional_rnn/%%s/layer_norm_basic_lstm_cell/%%s' % (inp, l, mid)
        layers.append({'fw_kernel': params[base % ('fw', 'kernel')], 'fw_bias': params[base % ('fw', 'bias')],
                       'bw_kernel': params[base % ('bw', 'kernel')], 'bw_bias': params[base % ('bw', 'bias')]})
    s = 'Speller/decoder/attention_wrapper/'
    sp = {'memory_kernel': params['Speller/memory_layer/kernel'],
          'query_kernel': params[s + 'location_aware_attention/query_layer/kernel'],
          'attention_v': params[s + 'location_aware_attention/attention_v'],
          'conv_kernel': params[s + 'location_aware_attention/conv1d/kernel'],
          'conv_dense_kernel': params[s + 'location_aware_attention/process_conv_features/kernel'],
          'out_kernel': params['Speller/decoder/dense/kernel'], 'out_bias': params['Speller/decoder/dense/bias']}
    l = 0
    while s + 'multi_rnn_cell/cell_%d/lstm_cell/kernel' % l in params:
        sp['cell_%d_kernel' % l] = params[s + 'multi_rnn_cell/cell_%d/lstm_cell/kernel' % l]
        sp['cell_%d_bias' % l] = params[s + 'multi_rnn_cell/cell_%d/lstm_cell/bias' % l]
        l += 1
    return layers, sp


def test_las_train_step_matches_oracle():
    """Listener (2 pBLSTM + BLSTM, odd T so the pyramid pads) + Speller + average_cross_entropy."""
    dev = torch.device('cuda', 0)
    B, T, D, H, NL, V, U = 6, 37, 40, 64, 2, 12, 7
    tr = _las_trainer(dev, H, NL, V, 64, 4, 7)
    tr.model.build({'features': D}, dev)
    store = tr.model.store
    x, lens, targets, tl = synthetic_las_batch(B, T, D, V, U, ragged=True)
    params = store.to_numpy()
    batch = ({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)},
             {'text': torch.from_numpy(targets).to(dev)}, {'text': torch.from_numpy(tl).to(dev)})
    loss, _ = tr.update(*batch)
    layers, sp = las_oracle_params(params, NL)
    enc, elens, caches = O.listener_fwd(x, lens, layers, 2)
    logits, ctx = O.speller_fwd(enc, elens, targets, tl, sp, 'location_aware', 2)
    ref_loss, dlogits = O.average_cross_entropy(logits, targets, tl, tl)
    dmem, gsp = O.speller_bwd(ctx, dlogits)
    _, glayers = O.listener_bwd(caches, dmem, 2)
    assert abs(float(loss) - ref_loss) / abs(ref_loss) < TOL
    grads = store.grads_numpy()
    glayers_got, gsp_got = las_oracle_params(grads, NL)
    for k in gsp:
        assert rel_err(gsp_got[k], gsp[k]) < 5 * TOL, k
    for l in range(NL + 1):
        for k in glayers[l]:
            assert rel_err(glayers_got[l][k], glayers[l][k]) < 5 * TOL, (l, k)


@pytest.mark.parametrize('attention,keep,sp,numfilt,fs', [
    ('location_aware', 0.5, 0.0, 3, 5),      # output dropout only (the LAS/TIMIT recipe: dropout = 0.5)
    ('vanilla', 1.0, 0.4, 0, 1),             # scheduled sampling only
    ('location_aware', 0.7, 0.1, 3, 5),      # both (sample_prob = 0.1 is the recipe default)
])
def test_speller_dropout_and_scheduled_sampling(attention, keep, sp, numfilt, fs):
    """Rows a6/a7 with the stochastic parts on (speller.py:37-41 DropoutWrapper(output_keep_prob), rnn_decoder.py:59-64
    ScheduledEmbeddingTrainingHelper).  TF's random streams cannot be reproduced; the kernels draw from a counter-based
    generator that the oracle restates, so for one seed the masks and the sampled tokens are THE SAME on both sides and
    logits / gradients are compared as usual.  The oracle itself is pinned by the torch twin (tests/test_oracle.py)."""
    from nabu_b200 import engine
    dev = torch.device('cuda', 0)
    B, Tm, E, V, H, NL, U, seed = 9, 14, 16, 8, 16, 2, 8, 4242
    rng = np.random.default_rng(int(keep * 100 + sp * 10))
    p = O.init_speller_params(rng, V, E, H, NL, attention, max(numfilt, 1), fs)
    memory = rng.standard_normal((B, Tm, E)).astype(np.float32)
    mem_len = rng.integers(Tm // 2, Tm + 1, size=B).astype(np.int32)
    tl = rng.integers(1, U + 1, size=B).astype(np.int32)
    tl[0] = U
    targets = rng.integers(0, V, size=(B, U)).astype(np.int32)
    dlog = rng.standard_normal((B, U, V)).astype(np.float32)
    for b in range(B):
        dlog[b, tl[b]:] = 0
    ref_logits, ctx = O.speller_fwd(memory, mem_len, targets, tl, p, attention, NL, np.float64, dropout_keep=keep,
                                    sample_prob=sp, seed=seed)
    ref_dmem, ref_g = O.speller_bwd(ctx, dlog.astype(np.float64))
    if sp > 0:
        teacher = np.concatenate([np.full((B, 1), V - 1), targets], 1)[:, :U]
        assert (ctx['ids_in'][:, :U] != teacher).any()
    sv = _svars(p, attention, NL, dev)
    mem_d = torch.tensor(memory, device=dev, requires_grad=True)
    logits = engine.speller(mem_d, torch.tensor(mem_len, device=dev), torch.tensor(targets, device=dev),
                            torch.tensor(tl, device=dev), sv, V, H, NL, attention, numfilt, fs, dropout_keep=keep,
                            sample_prob=sp, seed=seed)
    assert rel_err(logits.detach().cpu().numpy(), ref_logits) < TOL
    logits.backward(torch.tensor(dlog, device=dev))
    assert rel_err(mem_d.grad.cpu().numpy(), ref_dmem) < TOL
    for k, v in _grads(sv, NL, attention).items():
        assert rel_err(v, ref_g[k]) < TOL, k
    # a different seed gives different masks / samples
    other = engine.speller(mem_d.detach(), torch.tensor(mem_len, device=dev), torch.tensor(targets, device=dev),
                           torch.tensor(tl, device=dev), sv, V, H, NL, attention, numfilt, fs, dropout_keep=keep,
                           sample_prob=sp, seed=seed + 1)
    assert rel_err(other.detach().cpu().numpy(), ref_logits) > 1e-3


def test_las_recipe_settings_train():
    """The shipped LAS/TIMIT recipe's stochastic settings (config/recipes/LAS/TIMIT/model.cfg: input_noise = 0.6,
    dropout = 0.5 in listener and speller, sample_prob = 0.1) run through Trainer.update and learn."""
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    dev = torch.device('cuda', 0)
    V, D = 12, 40
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = listener\n'
                      'num_units = 64\nnum_layers = 2\npyramid_steps = 2\ninput_noise = 0.6\ndropout = 0.5\n[decoder]\n'
                      'decoder = speller\nnum_layers = 2\nnum_units = 32\ndropout = 0.5\nattention = vanilla\n'
                      'sample_prob = 0.1\n' % (V - 1))
    tconf = make_conf('[trainer]\ntrainer = standard\nloss = average_cross_entropy\ntrainlabels = 1\ntargets = text\n'
                      'initial_learning_rate = 3e-3\n')
    tr = trainer_factory.factory('standard')(tconf, None, mconf, None, None, None, 0, device=dev, seed=1)
    tr.num_steps = 1000
    tr.model.build({'features': D}, dev)
    x, lens, targets, tl = synthetic_las_batch(8, 48, D, V, 9, ragged=True)
    batch = ({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)},
             {'text': torch.from_numpy(targets).to(dev)}, {'text': torch.from_numpy(tl).to(dev)})
    losses = [float(tr.update(*batch)[0]) for _ in range(60)]
    assert np.isfinite(losses).all()
    assert np.mean(losses[-10:]) < 0.8 * np.mean(losses[:5]), (losses[:5], losses[-10:])
