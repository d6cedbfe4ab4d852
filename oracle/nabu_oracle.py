"""NumPy restatement of nabu's per-utterance training / decoding hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Pinned against the
reference's own Python executed over a restated TF-1.8 API
(tests/golden/tf18shim_cases); TensorFlow's kernels themselves are restated,
never run (no reference golden vectors exist; TF-1.8 is external).  Every function cites the
reference call site it follows (paths relative to /root/reference/nabu) and,
where the arithmetic is TensorFlow's, the TF-1.8 op it restates (SURVEY.md
appendix B).

All functions take a ``dtype`` (np.float64 = "truth", np.float32 = the
reference's working precision) and are written for clarity, not speed.
"""
from __future__ import annotations

import math
import numpy as np

__all__ = [
    'sigmoid', 'glorot_uniform', 'lstm_dir_fwd', 'lstm_dir_bwd', 'blstm_fwd',
    'blstm_bwd', 'pyramid_stack_fwd', 'pyramid_stack_bwd', 'dblstm_fwd',
    'dblstm_bwd', 'listener_fwd', 'listener_bwd', 'linear_fwd', 'linear_bwd',
    'log_softmax', 'ctc_loss_and_grad', 'ctc_brute_force', 'ctc_loss_mean',
    'average_cross_entropy', 'speller_fwd', 'speller_bwd', 'speller_step',
    'speller_zero_state', 'attention_keys', 'attention_step', 'attention_window', 'rng_u32',
    'rng_uniform', 'speller_dropout_mask', 'speller_sample_ids', 'tf_adam_clip',
    'exponential_decay', 'ctc_beam_search', 'las_beam_search',
    'edit_distance', 'init_blstm_params', 'init_speller_params',
    'init_linear_params',
]


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------

def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def glorot_uniform(rng, shape, dtype=np.float32):
    """tf.glorot_uniform_initializer (TF-1.8 get_variable default, appendix B1).

    1-D shapes use fan_in = fan_out = shape[0] (so LayerNormBasicLSTMCell's
    bias is *not* zero-initialised)."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:  # conv kernels [k, in, out]
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(dtype)


def init_blstm_params(rng, D, H, dtype=np.float32):
    """One BLSTM layer's variables (components/layer.py:35-42): per direction
    kernel[(D+H),4H] and bias[4H], both glorot-uniform (appendix B1)."""
    return {
        'fw_kernel': glorot_uniform(rng, (D + H, 4 * H), dtype),
        'fw_bias': glorot_uniform(rng, (4 * H,), dtype),
        'bw_kernel': glorot_uniform(rng, (D + H, 4 * H), dtype),
        'bw_bias': glorot_uniform(rng, (4 * H,), dtype),
    }


def init_linear_params(rng, D, V, dtype=np.float32):
    """tf.contrib.layers.linear (models/ed_decoders/dnn_decoder.py:53-57):
    glorot weights, zero biases (appendix B11)."""
    return {'weights': glorot_uniform(rng, (D, V), dtype),
            'biases': np.zeros((V,), dtype)}


# --------------------------------------------------------------------------
# a1: BLSTM  (components/layer.py:8-51; TF LayerNormBasicLSTMCell(layer_norm=
# False) under bidirectional_dynamic_rnn, appendix B1/B2)
# --------------------------------------------------------------------------

def lstm_dir_fwd(x, lens, kernel, bias, reverse, dtype=np.float64):
    """One direction of components/layer.py:45-47.

    z=[x,h].K+b; i,j,f,o=split(z); c'=c*sig(f+1)+sig(i)*tanh(j);
    h'=tanh(c')*sig(o).  Past len[b]: output 0, state frozen.  The backward
    direction runs over reverse_sequence(x, len): step s touches t=len-1-s.
    Returns y[B,T,H] and a cache for lstm_dir_bwd."""
    x = np.asarray(x, dtype)
    kernel = np.asarray(kernel, dtype)
    bias = np.asarray(bias, dtype)
    lens = np.asarray(lens, np.int64)
    B, T, D = x.shape
    H = kernel.shape[1] // 4
    Kx, Kh = kernel[:D], kernel[D:]
    gx = (x.reshape(B * T, D) @ Kx + bias).reshape(B, T, 4 * H)
    h = np.zeros((B, H), dtype)
    c = np.zeros((B, H), dtype)
    y = np.zeros((B, T, H), dtype)
    gates = np.zeros((B, T, 4 * H), dtype)   # activated i, g, f, o
    cs = np.zeros((B, T, H), dtype)
    ar = np.arange(B)
    for s in range(T):
        valid = s < lens
        if not valid.any():
            break
        t = np.where(valid, (lens - 1 - s) if reverse else s, 0)
        z = gx[ar, t] + h @ Kh
        i = sigmoid(z[:, :H])
        g = np.tanh(z[:, H:2 * H])
        f = sigmoid(z[:, 2 * H:3 * H] + dtype(1.0))
        o = sigmoid(z[:, 3 * H:])
        c_new = c * f + i * g
        h_new = np.tanh(c_new) * o
        vm = valid[:, None]
        c = np.where(vm, c_new, c)
        h = np.where(vm, h_new, h)
        vb = ar[valid]
        y[vb, t[valid]] = h_new[valid]
        gates[vb, t[valid]] = np.concatenate([i, g, f, o], 1)[valid]
        cs[vb, t[valid]] = c_new[valid]
    cache = dict(x=x, lens=lens, kernel=kernel, reverse=reverse, y=y,
                 gates=gates, cs=cs)
    return y, cache


def lstm_dir_bwd(cache, dy):
    """BPTT for lstm_dir_fwd.  Returns dx, dkernel, dbias."""
    x, lens, kernel = cache['x'], cache['lens'], cache['kernel']
    reverse, y, gates, cs = (cache['reverse'], cache['y'], cache['gates'],
                             cache['cs'])
    dtype = x.dtype.type
    B, T, D = x.shape
    H = kernel.shape[1] // 4
    Kx, Kh = kernel[:D], kernel[D:]
    dy = np.asarray(dy, dtype)
    dZ = np.zeros((B, T, 4 * H), dtype)
    hprev_all = np.zeros((B, T, H), dtype)
    dh_rec = np.zeros((B, H), dtype)
    dc_rec = np.zeros((B, H), dtype)
    ar = np.arange(B)
    for s in range(int(min(T, lens.max(initial=0))) - 1, -1, -1):
        valid = s < lens
        t = np.where(valid, (lens - 1 - s) if reverse else s, 0)
        tp = t + 1 if reverse else t - 1          # previous step's time index
        has_prev = valid & (s > 0)
        tpc = np.clip(tp, 0, T - 1)
        c_prev = np.where(has_prev[:, None], cs[ar, tpc], 0)
        h_prev = np.where(has_prev[:, None], y[ar, tpc], 0)
        gt = gates[ar, t]
        i, g, f, o = gt[:, :H], gt[:, H:2 * H], gt[:, 2 * H:3 * H], gt[:, 3 * H:]
        tc = np.tanh(cs[ar, t])
        dh = dy[ar, t] + dh_rec
        do = dh * tc
        dc = dc_rec + dh * o * (1 - tc * tc)
        di = dc * g
        dg = dc * i
        df = dc * c_prev
        dz = np.concatenate([di * i * (1 - i), dg * (1 - g * g),
                             df * f * (1 - f), do * o * (1 - o)], 1)
        vm = valid[:, None]
        dz = np.where(vm, dz, 0)
        dh_rec = np.where(vm, dz @ Kh.T, dh_rec)
        dc_rec = np.where(vm, dc * f, dc_rec)
        vb = ar[valid]
        dZ[vb, t[valid]] = dz[valid]
        hprev_all[vb, t[valid]] = h_prev[valid]
    dZ2 = dZ.reshape(B * T, 4 * H)
    dKx = x.reshape(B * T, D).T @ dZ2
    dKh = hprev_all.reshape(B * T, H).T @ dZ2
    db = dZ2.sum(0)
    dx = (dZ2 @ Kx.T).reshape(B, T, D)
    return dx, np.concatenate([dKx, dKh], 0), db


def blstm_fwd(x, lens, p, dtype=np.float64):
    """components/layer.py:8-51: concat((out_fw, out_bw), 2)."""
    yf, cf = lstm_dir_fwd(x, lens, p['fw_kernel'], p['fw_bias'], False, dtype)
    yb, cb = lstm_dir_fwd(x, lens, p['bw_kernel'], p['bw_bias'], True, dtype)
    return np.concatenate([yf, yb], 2), (cf, cb)


def blstm_bwd(cache, dy):
    cf, cb = cache
    H = cf['y'].shape[2]
    dxf, dkf, dbf = lstm_dir_bwd(cf, dy[:, :, :H])
    dxb, dkb, dbb = lstm_dir_bwd(cb, dy[:, :, H:])
    return dxf + dxb, {'fw_kernel': dkf, 'fw_bias': dbf,
                       'bw_kernel': dkb, 'bw_bias': dbb}


# --------------------------------------------------------------------------
# a2: pyramid_stack (components/ops.py:6-60)
# --------------------------------------------------------------------------

def pyramid_stack_fwd(x, lens, numsteps):
    """Zero-pad T to a multiple of numsteps, concat numsteps consecutive frames
    on the feature axis (frame i of each group first), len -> ceil(len/n)
    computed in float32 like ops.py:56-58."""
    B, T, C = x.shape
    Tp = int(math.ceil(T / numsteps) * numsteps)
    xp = np.concatenate([x, np.zeros((B, Tp - T, C), x.dtype)], 1)
    out = xp.reshape(B, Tp // numsteps, numsteps * C)
    new_lens = np.ceil(np.asarray(lens, np.float32) / np.float32(numsteps)
                       ).astype(np.int32)
    return out, new_lens


def pyramid_stack_bwd(dout, T, numsteps):
    B, T2, C2 = dout.shape
    return dout.reshape(B, T2 * numsteps, C2 // numsteps)[:, :T]


# --------------------------------------------------------------------------
# a4 / a3: DBLSTM and Listener encoders (models/ed_encoders/dblstm.py:34-57,
# listener.py:37-72) at is_training=False / input_noise=0 / dropout=1
# --------------------------------------------------------------------------

def dblstm_fwd(x, lens, layers, dtype=np.float64):
    caches = []
    h = np.asarray(x, dtype)
    for p in layers:
        h, c = blstm_fwd(h, lens, p, dtype)
        caches.append(c)
    return h, np.asarray(lens, np.int32), caches


def dblstm_bwd(caches, dy):
    grads = []
    for c in reversed(caches):
        dy, g = blstm_bwd(c, dy)
        grads.append(g)
    return dy, grads[::-1]


def listener_fwd(x, lens, layers, pyramid_steps=2, dtype=np.float64):
    """layers = num_layers pBLSTM params + 1 final BLSTM params."""
    caches = []
    h = np.asarray(x, dtype)
    lens = np.asarray(lens, np.int32)
    for p in layers[:-1]:
        y, c = blstm_fwd(h, lens, p, dtype)
        T = y.shape[1]
        h, lens = pyramid_stack_fwd(y, lens, pyramid_steps)
        caches.append((c, T))
    h, c = blstm_fwd(h, lens, layers[-1], dtype)
    caches.append((c, None))
    return h, lens, caches


def listener_bwd(caches, dy, pyramid_steps=2):
    grads = []
    c, _ = caches[-1]
    dy, g = blstm_bwd(c, dy)
    grads.append(g)
    for c, T in reversed(caches[:-1]):
        dy = pyramid_stack_bwd(dy, T, pyramid_steps)
        dy, g = blstm_bwd(c, dy)
        grads.append(g)
    return dy, grads[::-1]


# --------------------------------------------------------------------------
# a5: DNNDecoder with num_layers=0 (models/ed_decoders/dnn_decoder.py:53-57)
# --------------------------------------------------------------------------

def linear_fwd(x, p, dtype=np.float64):
    W = np.asarray(p['weights'], dtype)
    b = np.asarray(p['biases'], dtype)
    return np.asarray(x, dtype) @ W + b


def linear_bwd(x, p, dout):
    dtype = dout.dtype
    W = np.asarray(p['weights'], dtype)
    x2 = np.asarray(x, dtype).reshape(-1, x.shape[-1])
    d2 = dout.reshape(-1, dout.shape[-1])
    return (dout @ W.T), {'weights': x2.T @ d2, 'biases': d2.sum(0)}


# --------------------------------------------------------------------------
# a9: CTC (trainers/loss_functions.py:180-214 -> tf.nn.ctc_loss, appendix B7)
# --------------------------------------------------------------------------

def log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


def _ctc_one(logp, labels, blank):
    """alpha/beta in log space, TF convention: alpha includes the emission at
    t, beta excludes it, so alpha*beta is the posterior mass through (t, s)."""
    dtype = logp.dtype.type
    T, V = logp.shape
    L = len(labels)
    S = 2 * L + 1
    lp = np.full(S, blank, np.int64)
    lp[1::2] = labels
    repeats = int(np.sum(np.asarray(labels[1:]) == np.asarray(labels[:-1]))) \
        if L > 1 else 0
    if T < L + repeats:
        raise ValueError('Not enough time for target transition sequence '
                         '(required: %d, available: %d)' % (L + repeats, T))
    ninf = dtype(-np.inf)
    can_skip = np.zeros(S, bool)
    can_skip[2:] = (lp[2:] != blank) & (lp[2:] != lp[:-2])
    alpha = np.full((T, S), ninf, dtype)
    alpha[0, 0] = logp[0, blank]
    if S > 1:
        alpha[0, 1] = logp[0, lp[1]]
    with np.errstate(invalid='ignore', divide='ignore'):
        for t in range(1, T):
            a = alpha[t - 1]
            a1 = np.concatenate([[ninf], a])[:S]
            a2 = np.where(can_skip, np.concatenate([[ninf, ninf], a])[:S],
                          ninf)
            alpha[t] = np.logaddexp(np.logaddexp(a, a1), a2) + logp[t, lp]
        beta = np.full((T, S), ninf, dtype)
        beta[T - 1, S - 1] = 0
        if S > 1:
            beta[T - 1, S - 2] = 0
        skip_from = np.zeros(S, bool)           # s -> s+2 allowed
        skip_from[:-2] = can_skip[2:]
        for t in range(T - 2, -1, -1):
            b = beta[t + 1] + logp[t + 1, lp]
            b1 = np.concatenate([b, [ninf]])[1:S + 1]
            b2 = np.where(skip_from, np.concatenate([b, [ninf, ninf]])[2:S + 2],
                          ninf)
            beta[t] = np.logaddexp(np.logaddexp(b, b1), b2)
        log_p = np.logaddexp(alpha[T - 1, S - 1],
                             alpha[T - 1, S - 2] if S > 1 else ninf)
        ab = alpha + beta
        grad = np.exp(logp)
        for k in np.unique(lp):
            sel = ab[:, lp == k]
            m = sel.max(1, keepdims=True)
            m = np.where(np.isfinite(m), m, 0)
            lse = (m + np.log(np.exp(sel - m).sum(1, keepdims=True)))[:, 0]
            grad[:, k] -= np.exp(lse - log_p)
    return -log_p, grad


def ctc_loss_and_grad(logits, logit_lens, labels, label_lens, blank=None,
                      dtype=np.float64):
    """Per-utterance NLL [B] and d(NLL_b)/d(logits) [B,T,V] of tf.nn.ctc_loss
    with its defaults (preprocess_collapse_repeated=False, ctc_merge_repeated
    =True), blank = V-1, softmax applied internally; frames t >= len get zero
    gradient."""
    logits = np.asarray(logits, dtype)
    B, T, V = logits.shape
    blank = V - 1 if blank is None else blank
    loss = np.zeros(B, dtype)
    grad = np.zeros_like(logits)
    for b in range(B):
        Tb = int(logit_lens[b])
        lab = np.asarray(labels[b][:int(label_lens[b])], np.int64)
        lp = log_softmax(logits[b, :Tb])
        loss[b], grad[b, :Tb] = _ctc_one(lp, lab, blank)
    return loss, grad


def ctc_loss_mean(logits, logit_lens, labels, label_lens, dtype=np.float64):
    """loss_functions.CTC: reduce_mean over the batch (single output)."""
    loss, grad = ctc_loss_and_grad(logits, logit_lens, labels, label_lens,
                                   dtype=dtype)
    B = loss.shape[0]
    return loss.mean(), grad / dtype(B)


def ctc_brute_force(logits, labels, blank=None):
    """Exact -log p(l|x) by enumerating all V**T alignments (tiny cases)."""
    import itertools
    logits = np.asarray(logits, np.float64)
    T, V = logits.shape
    blank = V - 1 if blank is None else blank
    p = np.exp(log_softmax(logits))
    total = 0.0
    tgt = list(labels)
    for path in itertools.product(range(V), repeat=T):
        col = []
        prev = None
        for s in path:
            if s != prev and s != blank:
                col.append(s)
            prev = s
        if col == tgt:
            total += np.prod(p[np.arange(T), list(path)])
    return -np.log(total)


# --------------------------------------------------------------------------
# a10: average_cross_entropy (trainers/loss_functions.py:78-109,155-165)
# --------------------------------------------------------------------------

def average_cross_entropy(logits, targets, logit_lens, target_lens,
                          dtype=np.float64):
    """Masked sparse softmax CE summed over time, divided by the target
    length, batch mean.  Returns (loss, dlogits)."""
    logits = np.asarray(logits, dtype)
    B, U, V = logits.shape
    lp = log_softmax(logits)
    mask = (np.arange(U)[None, :] < np.asarray(logit_lens)[:, None])
    tl = np.asarray(target_lens, dtype)
    nll = -np.take_along_axis(lp, np.asarray(targets, np.int64)[:, :U, None],
                              2)[:, :, 0]
    per_utt = (nll * mask).sum(1) / tl
    loss = per_utt.mean()
    d = np.exp(lp)
    np.put_along_axis(d, np.asarray(targets, np.int64)[:, :U, None],
                      np.take_along_axis(
                          d, np.asarray(targets, np.int64)[:, :U, None], 2) - 1,
                      2)
    d *= (mask / tl[:, None] / dtype(B))[:, :, None]
    return loss, d


# --------------------------------------------------------------------------
# a6-a8: Speller (models/ed_decoders/{rnn_decoder.py:40-82,speller.py:29-69},
# components/rnn_cell.py:145-155, components/attention.py:142-240; TF
# AttentionWrapper / LSTMCell / BahdanauAttention, appendix B3-B6)
# --------------------------------------------------------------------------

def init_speller_params(rng, V, E, H, num_layers=2, attention='location_aware',
                        numfilt=10, filtersize=201, dtype=np.float32):
    """Variables of Speller.create_cell.  A = attention num_units =
    rnn_cell.output_size = H (speller.py:51)."""
    A = H
    p = {}
    for l in range(num_layers):
        din = (V + E) if l == 0 else H
        p['cell_%d_kernel' % l] = glorot_uniform(rng, (din + H, 4 * H), dtype)
        p['cell_%d_bias' % l] = np.zeros((4 * H,), dtype)
    p['memory_kernel'] = glorot_uniform(rng, (E, A), dtype)
    p['query_kernel'] = glorot_uniform(rng, (H, A), dtype)
    p['attention_v'] = glorot_uniform(rng, (A,), dtype)
    if attention == 'location_aware':
        p['conv_kernel'] = glorot_uniform(rng, (filtersize, 1, numfilt), dtype)
        p['conv_dense_kernel'] = glorot_uniform(rng, (numfilt, A), dtype)
    p['out_kernel'] = glorot_uniform(rng, (H + E, V), dtype)
    p['out_bias'] = np.zeros((V,), dtype)
    return p


def attention_keys(memory, mem_lens, p, dtype=np.float64):
    """values = memory * sequence_mask(len); keys = memory_layer(values)
    (no bias) -- computed once per utterance (appendix B4)."""
    memory = np.asarray(memory, dtype)
    B, Tm, E = memory.shape
    mask = (np.arange(Tm)[None, :] < np.asarray(mem_lens)[:, None])
    values = memory * mask[:, :, None]
    keys = values @ np.asarray(p['memory_kernel'], dtype)
    return values, keys, mask


def _conv_same(prev_align, Wc):
    """tf.layers.conv1d(alpha[...,None], F, k, padding='same', use_bias=False)
    (attention.py:163-169): cf[b,t,f] = sum_k alpha[b,t+k-padl]*Wc[k,0,f]."""
    B, Tm = prev_align.shape
    ksz, _, F = Wc.shape
    padl = (ksz - 1) // 2
    padr = ksz - 1 - padl
    ap = np.concatenate([np.zeros((B, padl), prev_align.dtype), prev_align,
                         np.zeros((B, padr), prev_align.dtype)], 1)
    # windows[b,t,k] = ap[b,t+k]
    idx = np.arange(Tm)[:, None] + np.arange(ksz)[None, :]
    win = ap[:, idx]                                  # [B,Tm,k]
    return win @ Wc[:, 0, :], win


def _lstm_cell(xin, h, c, K, b):
    """tf.contrib.rnn.LSTMCell step (appendix B3); gate order i,j,f,o,
    forget_bias 1.0."""
    H = h.shape[1]
    z = np.concatenate([xin, h], 1) @ K + b
    i = sigmoid(z[:, :H])
    g = np.tanh(z[:, H:2 * H])
    f = sigmoid(z[:, 2 * H:3 * H] + 1.0)
    o = sigmoid(z[:, 3 * H:])
    c_new = c * f + i * g
    h_new = np.tanh(c_new) * o
    return h_new, c_new, (i, g, f, o)


def speller_zero_state(B, Tm, E, H, num_layers, dtype=np.float64,
                       attention='vanilla'):
    al = np.zeros((B, Tm), dtype)
    if attention == 'windowed':           # WindowedAttention.initial_alignments
        al[:, 0] = 1                      # (attention.py:344-351): all mass on frame 0
    return {'h': [np.zeros((B, H), dtype) for _ in range(num_layers)],
            'c': [np.zeros((B, H), dtype) for _ in range(num_layers)],
            'attention': np.zeros((B, E), dtype),
            'alignments': al}


def rng_u32(seed, a, b, c):
    """The counter-based generator of the CUDA decoder kernels (csrc/speller_kernels.cuh: dec_rng_u32), restated: three
    rounds of the lowbias32 integer finaliser over (seed, a, b, c).  Arrays broadcast; all arithmetic mod 2^32."""
    M = np.uint64(0xFFFFFFFF)

    def mix(x):
        x = x & M
        x ^= x >> np.uint64(16)
        x = (x * np.uint64(0x7FEB352D)) & M
        x ^= x >> np.uint64(15)
        x = (x * np.uint64(0x846CA68B)) & M
        x ^= x >> np.uint64(16)
        return x
    seed, a, b, c = (np.asarray(v, np.uint64) for v in (seed, a, b, c))
    x = mix((seed ^ np.uint64(0x9E3779B9)) + a)
    x = mix(x + ((b * np.uint64(0x85EBCA6B)) & M))
    x = mix(x + ((c * np.uint64(0xC2B2AE35)) & M))
    return x.astype(np.uint32)


def rng_uniform(seed, a, b, c):
    """uniform in [0, 1) with 24 bits, as the kernels draw it"""
    return (rng_u32(seed, a, b, c) >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def speller_dropout_mask(seed, layer, u, B, H, keep):
    """DropoutWrapper(output_keep_prob=keep) mask of LSTM layer `layer` at decoder step u: [B, H] of 0/1"""
    r, j = np.meshgrid(np.arange(B), np.arange(H), indexing='ij')
    return (rng_uniform(seed, layer * 65536 + u, r, j) < keep).astype(np.float64)


def speller_sample_ids(seed, u, logits, sample_prob):
    """ScheduledEmbeddingTrainingHelper.sample (appendix B6) with the kernels' uniforms: -1 where the teacher's token
    is kept, else a draw from Categorical(logits) by inverse CDF over the softmax."""
    B, V = logits.shape
    r = np.arange(B)
    take = rng_uniform(seed, 0x40000000 + u, r, 0) < sample_prob
    ub = rng_uniform(seed, 0x40000000 + u, r, 1)
    p = np.exp(logits - logits.max(1, keepdims=True))
    cdf = np.cumsum(p, 1)
    ids = (cdf < (ub * cdf[:, -1])[:, None]).sum(1)
    return np.where(take, np.minimum(ids, V - 1), -1)


def attention_window(prev_align, left, right):
    """WindowedAttention's score window (attention.py:372-383): True where the
    score is kept.  half_step = cumsum(prev) > 0.5; the window is the xor of
    half_step shifted left by left+1 (True shifted in) and shifted right by
    `right` (False shifted in).  right must be >= 1 (the reference slices
    half_step[:, :-right])."""
    half = np.cumsum(prev_align, 1) > 0.5
    B, Tm = half.shape
    sl = np.ones((B, Tm), bool)
    if left + 1 < Tm:
        sl[:, :Tm - left - 1] = half[:, left + 1:]
    sr = np.zeros((B, Tm), bool)
    if right < Tm:
        sr[:, right:] = half[:, :Tm - right]
    return np.logical_xor(sl, sr)


def attention_step(query, prev_align, values, keys, mask, p, attention, dtype,
                   probability_fn='softmax', window=None):
    """One call of the attention mechanism + the AttentionWrapper's context
    (components/attention.py:142-184 LocationAwareAttention.__call__,
    :186-240 _bahdanau_location_score, :24-30 vanilla, :294-396 windowed;
    appendix B4/B5): alignments [B,Tm], context [B,E] and the intermediates
    the backward needs."""
    q = query @ np.asarray(p['query_kernel'], dtype)
    pre = q[:, None, :] + keys
    cf = win = None
    if attention == 'location_aware':
        Wc = np.asarray(p['conv_kernel'], dtype)
        Wd = np.asarray(p['conv_dense_kernel'], dtype)
        cf, win = _conv_same(prev_align, Wc)
        pre = pre + cf @ Wd
    sact = np.tanh(pre)
    e = sact @ np.asarray(p['attention_v'], dtype)
    if attention == 'windowed':
        e = np.where(attention_window(prev_align, *window), e, -np.inf)
    e = np.where(mask, e, -np.inf)
    ssum = None
    if probability_fn == 'softmax':
        m = e.max(1, keepdims=True)
        ex = np.exp(e - m)
        alpha = ex / ex.sum(1, keepdims=True)
    else:
        # components/attention.py:6-55: tf.sigmoid or normalized_sigmoid on the -inf masked scores
        with np.errstate(over='ignore'):
            sig = np.where(mask, 1.0 / (1.0 + np.exp(-np.where(mask, e, 0))), 0).astype(dtype)
        if probability_fn == 'sigmoid':
            alpha = sig
        elif probability_fn == 'normalized_sigmoid':
            ssum = sig.sum(1, keepdims=True)
            alpha = sig / ssum
        else:
            raise ValueError('unknown probability_fn %r' % probability_fn)
    ctx = np.einsum('bt,bte->be', alpha, values)
    return alpha, ctx, (sact, cf, win, ssum)


def speller_step(ids, state, values, keys, mask, p, attention, dtype,
                 want_cache=False, probability_fn='softmax', window=None,
                 drop=None):
    """One AttentionProjectionWrapper(AttentionWrapper(MultiRNNCell)) step
    (rnn_cell.py:145-155 + appendix B5) on one-hot inputs `ids` [B]."""
    num_layers = len(state['h'])
    V = p['out_bias'].shape[0]
    B = ids.shape[0]
    onehot = np.zeros((B, V), dtype)
    onehot[np.arange(B), ids] = 1
    xin = np.concatenate([onehot, state['attention']], 1)
    hs, cs, gs, xins = [], [], [], []
    inp = xin
    for l in range(num_layers):
        K = np.asarray(p['cell_%d_kernel' % l], dtype)
        b = np.asarray(p['cell_%d_bias' % l], dtype)
        xins.append(inp)
        h_new, c_new, g = _lstm_cell(inp, state['h'][l], state['c'][l], K, b)
        hs.append(h_new)
        cs.append(c_new)
        gs.append(g)
        # DropoutWrapper(output_keep_prob): the OUTPUT is dropped, the state is not
        inp = h_new if drop is None else h_new * drop[1][l] / drop[0]
    query = inp
    alpha, ctx, (sact, cf, win, ssum) = attention_step(
        query, state['alignments'], values, keys, mask, p, attention, dtype,
        probability_fn, window)
    out_in = np.concatenate([query, ctx], 1)
    logits = out_in @ np.asarray(p['out_kernel'], dtype) \
        + np.asarray(p['out_bias'], dtype)
    new_state = {'h': hs, 'c': cs, 'attention': ctx, 'alignments': alpha}
    cache = None
    if want_cache:
        cache = dict(ids=ids, xins=xins, gates=gs, cs=cs, c_prev=state['c'],
                     h_prev=state['h'], query=query, sact=sact, cf=cf,
                     win=win, alpha=alpha, ctx=ctx, out_in=out_in,
                     prev_align=state['alignments'], ssum=ssum, drop=drop)
    return logits, new_state, cache


def speller_fwd(memory, mem_lens, targets, target_lens, p, attention='vanilla',
                num_layers=2, dtype=np.float64, probability_fn='softmax',
                window=None, dropout_keep=1.0, sample_prob=0.0, seed=0):
    """RNNDecoder._decode with sample_prob=0, dropout=1 (rnn_decoder.py:40-82):
    prepend SOS=V-1, teacher-forced dynamic_decode(impute_finished=True).
    Returns logits [B, max(target_lens), V] (zeros past each target length)."""
    memory = np.asarray(memory, dtype)
    B, Tm, E = memory.shape
    V = p['out_bias'].shape[0]
    H = p['query_kernel'].shape[0]
    target_lens = np.asarray(target_lens, np.int64)
    U = int(target_lens.max())
    values, keys, mask = attention_keys(memory, mem_lens, p, dtype)
    ids_in = np.concatenate([np.full((B, 1), V - 1, np.int64),
                             np.asarray(targets, np.int64)[:, :U]], 1)
    state = speller_zero_state(B, Tm, E, H, num_layers, dtype, attention)
    logits = np.zeros((B, U, V), dtype)
    caches = []
    ids_in = ids_in.copy()
    for u in range(U):
        active = (u < target_lens)
        drop = None
        if dropout_keep < 1:
            drop = (dropout_keep, [speller_dropout_mask(seed, l, u, B, H, dropout_keep).astype(dtype)
                                   for l in range(num_layers)])
        lg, ns, cache = speller_step(ids_in[:, u], state, values, keys, mask,
                                     p, attention, dtype, want_cache=True,
                                     probability_fn=probability_fn,
                                     window=window, drop=drop)
        if sample_prob > 0 and u + 1 < U:
            smp = speller_sample_ids(seed, u, lg, sample_prob)
            ids_in[:, u + 1] = np.where(smp >= 0, smp, ids_in[:, u + 1])
        am = active[:, None]
        logits[:, u] = np.where(am, lg, 0)
        state = {
            'h': [np.where(am, a, b) for a, b in zip(ns['h'], state['h'])],
            'c': [np.where(am, a, b) for a, b in zip(ns['c'], state['c'])],
            'attention': np.where(am, ns['attention'], state['attention']),
            'alignments': np.where(am, ns['alignments'], state['alignments']),
        }
        cache['active'] = active
        caches.append(cache)
    ctx = dict(caches=caches, values=values, keys=keys, mask=mask,
               memory=memory, p=p, attention=attention, num_layers=num_layers,
               dtype=dtype, probability_fn=probability_fn, ids_in=ids_in)
    return logits, ctx


def speller_bwd(ctx, dlogits):
    """Manual BPTT of speller_fwd.  Returns (dmemory, grads dict)."""
    caches, values, keys, mask = (ctx['caches'], ctx['values'], ctx['keys'],
                                  ctx['mask'])
    p, attention, NL, dtype = (ctx['p'], ctx['attention'], ctx['num_layers'],
                               ctx['dtype'])
    P = {k: np.asarray(v, dtype) for k, v in p.items()}
    B, Tm, E = values.shape
    H = P['query_kernel'].shape[0]
    V = P['out_bias'].shape[0]
    g = {k: np.zeros_like(v) for k, v in P.items()}
    dvalues = np.zeros_like(values)
    dkeys = np.zeros_like(keys)
    dh = [np.zeros((B, H), dtype) for _ in range(NL)]
    dc = [np.zeros((B, H), dtype) for _ in range(NL)]
    dattn = np.zeros((B, E), dtype)       # grad wrt state.attention
    dalign = np.zeros((B, Tm), dtype)     # grad wrt state.alignments
    for u in range(len(caches) - 1, -1, -1):
        c = caches[u]
        am = c['active'][:, None]
        dl = np.where(am, dlogits[:, u], 0)
        # rows that are inactive pass their (zero) carried grads through
        g['out_kernel'] += c['out_in'].T @ dl
        g['out_bias'] += dl.sum(0)
        dout_in = dl @ P['out_kernel'].T
        dquery = dout_in[:, :H]                 # wrt the top layer's OUTPUT (the carried state gradient joins below)
        dctx = dout_in[:, H:] + np.where(am, dattn, 0)
        dalpha = np.einsum('be,bte->bt', dctx, values) + np.where(am, dalign, 0)
        dvalues += c['alpha'][:, :, None] * dctx[:, None, :]
        pf = ctx.get('probability_fn', 'softmax')
        if pf == 'softmax':
            de = c['alpha'] * (dalpha - (c['alpha'] * dalpha).sum(1, keepdims=True))
        elif pf == 'sigmoid':
            de = dalpha * c['alpha'] * (1 - c['alpha'])
        else:                                   # alpha = s / S,  s = sigmoid(e)
            sg = c['alpha'] * c['ssum']
            de = (dalpha - (c['alpha'] * dalpha).sum(1, keepdims=True)) / c['ssum'] * sg * (1 - sg)
        dpre = de[:, :, None] * P['attention_v'][None, None, :] \
            * (1 - c['sact'] ** 2)
        g['attention_v'] += np.einsum('bt,bta->a', de, c['sact'])
        dq = dpre.sum(1)
        dkeys += dpre
        new_dalign = np.zeros((B, Tm), dtype)
        if attention == 'location_aware':
            Wc, Wd = P['conv_kernel'], P['conv_dense_kernel']
            g['conv_dense_kernel'] += np.einsum('btf,bta->fa', c['cf'], dpre)
            dcf = dpre @ Wd.T                               # [B,Tm,F]
            g['conv_kernel'][:, 0, :] += np.einsum('btk,btf->kf', c['win'],
                                                   dcf)
            dwin = dcf @ Wc[:, 0, :].T                      # [B,Tm,k]
            ksz = Wc.shape[0]
            padl = (ksz - 1) // 2
            dap = np.zeros((B, Tm + ksz - 1), dtype)
            for k in range(ksz):
                dap[:, k:k + Tm] += dwin[:, :, k]
            new_dalign = dap[:, padl:padl + Tm]
        g['query_kernel'] += c['query'].T @ dq
        dquery = dquery + dq @ P['query_kernel'].T
        # LSTM layers, top to bottom
        dinp = dquery
        new_dh, new_dc = [None] * NL, [None] * NL
        for l in range(NL - 1, -1, -1):
            K = P['cell_%d_kernel' % l]
            i, gg, f, o = c['gates'][l]
            if c.get('drop') is not None:         # gradient wrt the layer's OUTPUT -> wrt h
                dinp = dinp * c['drop'][1][l] / c['drop'][0]
            dh_tot = dinp + np.where(am, dh[l], 0)
            tc = np.tanh(c['cs'][l])
            do = dh_tot * tc
            dcc = np.where(am, dc[l], 0) + dh_tot * o * (1 - tc * tc)
            dz = np.concatenate([dcc * gg * i * (1 - i), dcc * i * (1 - gg * gg),
                                 dcc * c['c_prev'][l] * f * (1 - f),
                                 do * o * (1 - o)], 1)
            dz = np.where(am, dz, 0)
            xh = np.concatenate([c['xins'][l], c['h_prev'][l]], 1)
            g['cell_%d_kernel' % l] += xh.T @ dz
            g['cell_%d_bias' % l] += dz.sum(0)
            dxh = dz @ K.T
            din = xh.shape[1] - H
            new_dh[l] = np.where(am, dxh[:, din:], dh[l])
            new_dc[l] = np.where(am, dcc * f, dc[l])
            dinp = dxh[:, :din]
        dattn = np.where(am, dinp[:, V:], dattn)
        dalign = np.where(am, new_dalign, dalign)
        dh, dc = new_dh, new_dc
    g['memory_kernel'] += values.reshape(B * Tm, E).T @ dkeys.reshape(B * Tm, -1)
    dvalues += dkeys @ P['memory_kernel'].T
    dmemory = dvalues * mask[:, :, None]
    return dmemory, g


# --------------------------------------------------------------------------
# a11: Trainer._update (trainers/trainer.py:512-580,161-166; appendix B10)
# --------------------------------------------------------------------------

def exponential_decay(lr0, step, decay_steps, decay_rate, fact=1.0):
    """tf.train.exponential_decay (no staircase) * learning_rate_fact."""
    return lr0 * decay_rate ** (float(step) / float(decay_steps)) * fact


def tf_adam_clip(theta, grad, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8,
                 clip=1.0, dtype=np.float32):
    """clip_by_value(g,-1,1) then tf.train.AdamOptimizer step number t>=1:
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)."""
    theta = np.asarray(theta, dtype)
    g = np.clip(np.asarray(grad, dtype), -clip, clip)
    m = np.asarray(m, dtype)
    v = np.asarray(v, dtype)
    # tensorflow/core/kernels/training_ops.cc ApplyAdam: m += (g-m)*(1-b1); v += (g*g-v)*(1-b2)
    m = m + (g - m) * (dtype(1) - dtype(beta1))
    v = v + (g * g - v) * (dtype(1) - dtype(beta2))
    lr_t = dtype(lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t))
    theta = theta - lr_t * m / (np.sqrt(v) + dtype(eps))
    return theta, m, v


# --------------------------------------------------------------------------
# a12: CTCDecoder (neuralnetworks/decoders/ctc_decoder.py:29-61 ->
# tf.nn.ctc_beam_search_decoder defaults beam_width=100, top_paths=1,
# merge_repeated=True; appendix B8 -- tensorflow/core/util/ctc/
# ctc_beam_search.h restated)
# --------------------------------------------------------------------------

_NINF32 = np.float32(-np.inf)


def _lse32(a, b):
    """TF's LogSumExp on float32 (ctc_loss_util.h)."""
    if a == _NINF32:
        return b
    if b == _NINF32:
        return a
    if a > b:
        return np.float32(np.float32(np.log1p(np.exp(np.float32(b - a)))) + a)
    return np.float32(np.float32(np.log1p(np.exp(np.float32(a - b)))) + b)


class _BeamEntry(object):
    __slots__ = ('parent', 'label', 'children', 'old', 'new')

    def __init__(self, parent, label):
        self.parent = parent
        self.label = label
        self.children = None
        # (total, blank, label) log-probabilities
        self.old = [_NINF32, _NINF32, _NINF32]
        self.new = [_NINF32, _NINF32, _NINF32]

    def active(self):
        return self.new[0] != _NINF32


def ctc_beam_search(logits, seq_len, beam_width=100, merge_repeated=True,
                    blank=None):
    """One utterance.  logits [T,V] raw (the TF op does not normalise; Step
    only subtracts the per-frame max).  Returns the best path's label list
    (int32), with TF's merge_repeated collapsing of adjacent equal labels."""
    logits = np.asarray(logits, np.float32)
    T, V = logits.shape
    blank = V - 1 if blank is None else blank
    root = _BeamEntry(None, -1)
    root.new = [np.float32(0), np.float32(0), _NINF32]
    leaves = [root]
    with np.errstate(over='ignore', invalid='ignore'):
        for t in range(int(seq_len)):
            x = logits[t] - logits[t].max()
            # Extract(): descending newp.total (stable wrt insertion order)
            branches = sorted(leaves, key=lambda e: -e.new[0])
            leaves = []
            for b in branches:
                b.old = list(b.new)
            for b in branches:
                if b.parent is not None:
                    if b.parent.active():
                        prev = b.parent.old[1] if b.label == b.parent.label \
                            else b.parent.old[0]
                        b.new[2] = _lse32(b.new[2], prev)
                    b.new[2] = np.float32(b.new[2] + x[b.label])
                b.new[1] = np.float32(b.old[0] + x[blank])
                b.new[0] = _lse32(b.new[1], b.new[2])
                leaves.append(b)

            def bottom():
                return min(leaves, key=lambda e: e.new[0])

            def is_candidate(prob):
                return prob[0] > _NINF32 and (
                    len(leaves) < beam_width or prob[0] > bottom().new[0])

            for b in branches:
                if not is_candidate(b.old):
                    continue
                if b.children is None:
                    b.children = [_BeamEntry(b, c) for c in range(V)
                                  if c != blank]
                for c in b.children:
                    if c.active():
                        continue
                    prev = b.old[1] if c.label == b.label else b.old[0]
                    c.new[1] = _NINF32
                    c.new[2] = np.float32(x[c.label] + prev)
                    c.new[0] = c.new[2]
                    if is_candidate(c.new):
                        if len(leaves) == beam_width:
                            bt = bottom()
                            bt.new = [_NINF32, _NINF32, _NINF32]
                            leaves.remove(bt)
                        leaves.append(c)
                    else:
                        c.old = [_NINF32, _NINF32, _NINF32]
                        c.new = [_NINF32, _NINF32, _NINF32]
    best = max(leaves, key=lambda e: e.new[0])
    labels = []
    prev_label = -1
    e = best
    while e.parent is not None:
        if not merge_repeated or e.label != prev_label:
            labels.append(e.label)
        prev_label = e.label
        e = e.parent
    return np.asarray(labels[::-1], np.int32), np.float32(-best.new[0])


# --------------------------------------------------------------------------
# a13: LAS beam search (neuralnetworks/decoders/beam_search_decoder.py:30-112,
# components/beam_search_decoder.py:136-485; appendix B9)
# --------------------------------------------------------------------------

def las_beam_search(memory, mem_lens, p, beam_width, max_steps,
                    attention='vanilla', num_layers=2, length_penalty=1.0,
                    temperature=1.0, dtype=np.float32, probability_fn='softmax',
                    window=None):
    """Returns sequences[B,W,L] int32, lengths[B,W] int32, scores[B,W] f32,
    alignments[B,W,L,Tm] f32 exactly as BeamSearchDecoder.__call__ does."""
    memory = np.asarray(memory, dtype)
    B, Tm, E = memory.shape
    W = beam_width
    V = p['out_bias'].shape[0]
    H = p['query_kernel'].shape[0]
    eos = V - 1
    f32 = np.float32
    fmax = np.finfo(np.float32).max
    # tile_batch: row b*W+w
    mem_t = np.repeat(memory, W, axis=0)
    len_t = np.repeat(np.asarray(mem_lens), W, axis=0)
    values, keys, mask = attention_keys(mem_t, len_t, p, dtype)
    state = speller_zero_state(B * W, Tm, E, H, num_layers, dtype, attention)
    ids = np.full((B, W), eos, np.int64)                 # start tokens
    logprobs = np.concatenate([np.zeros((B, 1), f32),
                               np.full((B, W - 1), -np.inf, f32)], 1)
    lengths = np.zeros((B, W), np.int32)
    finished = np.zeros((B, W), bool)
    loop_finished = np.zeros((B, W), bool)
    pred_hist, parent_hist, align_hist = [], [], []

    def score(lp, ln):
        if length_penalty == 0:
            return lp
        pen = ((f32(5.) + ln.astype(f32)) ** f32(length_penalty)) \
            / (f32(6.) ** f32(length_penalty))
        return (lp / pen).astype(f32)

    def flat_state(st):
        return [*st['h'], *st['c'], st['attention'], st['alignments']]

    def unflat_state(fl):
        return {'h': fl[:num_layers], 'c': fl[num_layers:2 * num_layers],
                'attention': fl[-2], 'alignments': fl[-1]}

    t = 0
    with np.errstate(over='ignore', invalid='ignore', divide='ignore'):
        while not loop_finished.all():
            logits, new_state, _ = speller_step(ids.reshape(-1), state, values,
                                                keys, mask, p, attention,
                                                dtype, probability_fn=probability_fn,
                                                window=window)
            out = (logits.astype(f32) / f32(temperature)).reshape(B, W, V)
            new_lp = log_softmax(out).astype(f32)
            new_lp = np.where(finished[:, :, None], -fmax, new_lp)
            cand_lp = (logprobs[:, :, None] + new_lp).reshape(B, W * V)
            cand_ids = np.tile(np.arange(V), (B, W))
            cand_len = np.repeat(lengths, V, axis=1)
            cand_len = np.where(cand_ids == eos, cand_len, cand_len + 1)
            stay_lp = np.where(finished, logprobs, -fmax).astype(f32)
            cand_ids = np.concatenate([cand_ids, np.full((B, W), eos)], 1)
            cand_lp = np.concatenate([cand_lp, stay_lp], 1)
            cand_len = np.concatenate([cand_len, lengths], 1)
            sc = score(cand_lp, cand_len)
            # tf.nn.top_k: descending value, ties -> lowest index first
            order = np.argsort(-sc, axis=1, kind='stable')[:, :W]
            parent = order // V
            parent = np.where(parent == W, order % V, parent)
            is_stay = order >= W * V
            bi = np.arange(B)[:, None]
            lengths = cand_len[bi, order].astype(np.int32)
            ids = cand_ids[bi, order]
            logprobs = cand_lp[bi, order]
            # expansions take the new cell state of their parent, stay
            # hypotheses keep the old state of their own slot
            src = (bi * W + parent).reshape(-1)
            stay = is_stay.reshape(-1)[:, None]
            old_fl, new_fl = flat_state(state), flat_state(new_state)
            state = unflat_state([np.where(stay, o[src], n[src])
                                  for o, n in zip(old_fl, new_fl)])
            finished = (ids == eos)
            pred_hist.append(ids.astype(np.int32))
            parent_hist.append(parent.astype(np.int32))
            align_hist.append(state['alignments'].reshape(B, W, Tm)
                              .astype(f32))
            loop_finished = loop_finished | finished | (t + 1 >= max_steps)
            t += 1
    L = len(pred_hist)
    seqs = np.zeros((B, W, L), np.int32)
    aligns = np.zeros((B, W, L, Tm), f32)
    beams = np.tile(np.arange(W), (B, 1))
    bi = np.arange(B)[:, None]
    for tt in range(L - 1, -1, -1):
        seqs[:, :, tt] = pred_hist[tt][bi, beams]
        aligns[:, :, tt] = align_hist[tt][bi, beams]
        beams = parent_hist[tt][bi, beams]
    scores = score(logprobs, lengths)
    return seqs, lengths, scores, aligns


def edit_distance(a, b):
    """tf.edit_distance(normalize=False) for one pair."""
    a, b = list(a), list(b)
    d = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        prev, d[0] = d[0], i
        for j in range(1, len(b) + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (a[i - 1] != b[j - 1]))
            prev = cur
    return d[len(b)]
