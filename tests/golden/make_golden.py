"""Freezes the outputs of the oracle's INDEPENDENT witnesses as small fixtures (tests/golden/*.npz).

The reference (Python-2 / TF-1.8) cannot be imported here and ships no golden vectors, so these
fixtures do not come from the reference: they come from torch CPU ops (F.ctc_loss, nn.LSTM on packed
sequences, an autograd twin of the attention decoder) and brute-force enumeration.  The oracle and the
CUDA kernels are both checked against them.     python tests/golden/make_golden.py
"""
import itertools
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.test_oracle import _tf_to_torch_lstm, _torch_speller  # noqa: E402


def glorot(rng, shape):
    fi, fo = (shape[0], shape[0]) if len(shape) == 1 else (shape[-2] * int(np.prod(shape[:-2])),
                                                           shape[-1] * int(np.prod(shape[:-2])))
    lim = np.sqrt(6.0 / (fi + fo))
    return rng.uniform(-lim, lim, size=shape)


def golden_ctc():
    rng = np.random.default_rng(100)
    B, T, V, L = 6, 50, 29, 6
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32).astype(np.float64)
    lens = np.array([50, 41, 33, 50, 27, 12], np.int32)
    labels = rng.integers(0, V - 1, size=(B, L)).astype(np.int32)
    labels[0, 1] = labels[0, 0]
    ll = np.array([6, 5, 0, 3, 6, 1], np.int32)
    lt = torch.tensor(logits, requires_grad=True)
    loss = F.ctc_loss(F.log_softmax(lt, -1).transpose(0, 1), torch.tensor(labels.astype(np.int64)),
                      torch.tensor(lens.astype(np.int64)), torch.tensor(ll.astype(np.int64)), blank=V - 1,
                      reduction='none')
    loss.sum().backward()
    grad = lt.grad.numpy()
    for b in range(B):
        grad[b, lens[b]:] = 0
    np.savez(os.path.join(HERE, 'ctc.npz'), logits=logits.astype(np.float32), lens=lens, labels=labels, ll=ll,
             loss=loss.detach().numpy(), grad=grad)


def golden_blstm():
    rng = np.random.default_rng(101)
    B, T, D, H = 5, 11, 8, 64
    x = rng.standard_normal((B, T, D)).astype(np.float32).astype(np.float64)
    lens = np.array([11, 8, 11, 3, 6], np.int32)
    p = {d + '_kernel': glorot(rng, (D + H, 4 * H)) for d in ('fw', 'bw')}
    p.update({d + '_bias': glorot(rng, (4 * H,)) for d in ('fw', 'bw')})
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
    lstm = torch.nn.LSTM(D, H, batch_first=True, bidirectional=True).double()
    with torch.no_grad():
        for sfx, d in (('', 'fw'), ('_reverse', 'bw')):
            wi, wh, bb = _tf_to_torch_lstm(p[d + '_kernel'], p[d + '_bias'], D)
            getattr(lstm, 'weight_ih_l0' + sfx).copy_(torch.tensor(wi))
            getattr(lstm, 'weight_hh_l0' + sfx).copy_(torch.tensor(wh))
            getattr(lstm, 'bias_ih_l0' + sfx).copy_(torch.tensor(bb))
            getattr(lstm, 'bias_hh_l0' + sfx).zero_()
    xt = torch.tensor(x, requires_grad=True)
    pk = torch.nn.utils.rnn.pack_padded_sequence(xt, torch.tensor(lens.astype(np.int64)), batch_first=True,
                                                 enforce_sorted=False)
    out, _ = lstm(pk)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=T)
    dy = rng.standard_normal(out.shape).astype(np.float32).astype(np.float64)
    (out * torch.tensor(dy)).sum().backward()

    def back(gih, ghh, gb):
        i, f, g, o = np.split(np.concatenate([gih, ghh], 1), 4, 0)
        bi, bf, bg, bo = np.split(gb, 4)
        return np.concatenate([i, g, f, o], 0).T, np.concatenate([bi, bg, bf, bo])
    dkf, dbf = back(lstm.weight_ih_l0.grad.numpy(), lstm.weight_hh_l0.grad.numpy(), lstm.bias_ih_l0.grad.numpy())
    dkb, dbb = back(lstm.weight_ih_l0_reverse.grad.numpy(), lstm.weight_hh_l0_reverse.grad.numpy(),
                    lstm.bias_ih_l0_reverse.grad.numpy())
    np.savez(os.path.join(HERE, 'blstm.npz'), x=x.astype(np.float32), lens=lens, dy=dy.astype(np.float32),
             y=out.detach().numpy().astype(np.float32), dx=xt.grad.numpy().astype(np.float32),
             dkf=dkf.astype(np.float32), dbf=dbf.astype(np.float32), dkb=dkb.astype(np.float32), dbb=dbb.astype(np.float32),
             **{k: v.astype(np.float32) for k, v in p.items()})


def golden_speller():
    rng = np.random.default_rng(102)
    B, Tm, E, V, H, NL, U = 4, 10, 16, 7, 8, 2, 5
    p = {}
    for l in range(NL):
        p['cell_%d_kernel' % l] = glorot(rng, ((V + E if l == 0 else H) + H, 4 * H))
        p['cell_%d_bias' % l] = rng.standard_normal(4 * H) * 0.1
    p['memory_kernel'] = glorot(rng, (E, H))
    p['query_kernel'] = glorot(rng, (H, H))
    p['attention_v'] = glorot(rng, (H,))
    p['conv_kernel'] = glorot(rng, (5, 1, 3))
    p['conv_dense_kernel'] = glorot(rng, (3, H))
    p['out_kernel'] = glorot(rng, (H + E, V))
    p['out_bias'] = rng.standard_normal(V) * 0.1
    p = {k: v.astype(np.float32).astype(np.float64) for k, v in p.items()}
    memory = rng.standard_normal((B, Tm, E)).astype(np.float32).astype(np.float64)
    mem_lens = np.array([10, 7, 10, 4], np.int32)
    tl = np.array([5, 3, 1, 4], np.int32)
    targets = rng.integers(0, V, size=(B, U)).astype(np.int32)
    dlog = rng.standard_normal((B, U, V)).astype(np.float32).astype(np.float64)
    for b in range(B):
        dlog[b, tl[b]:] = 0
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    tm = torch.tensor(memory, requires_grad=True)
    logits = _torch_speller(tm, mem_lens, targets, tl, tp, 'location_aware', NL)
    (logits * torch.tensor(dlog)).sum().backward()
    out = {'memory': memory.astype(np.float32), 'mem_lens': mem_lens, 'tl': tl, 'targets': targets,
           'dlog': dlog.astype(np.float32), 'logits': logits.detach().numpy(), 'dmemory': tm.grad.numpy()}
    for k, v in p.items():
        out['p_' + k] = v.astype(np.float32)
        out['g_' + k] = tp[k].grad.numpy()
    np.savez(os.path.join(HERE, 'speller.npz'), **out)


def golden_ctc_beam():
    """Exhaustive best labelling (beam wider than the hypothesis space), merge_repeated=False."""
    rng = np.random.default_rng(103)
    T, V, N = 6, 4, 8
    logits = (rng.standard_normal((N, T, V)) * 2).astype(np.float32)
    best = np.zeros((N, T), np.int32)
    best_len = np.zeros(N, np.int32)
    for n in range(N):
        x = logits[n] - logits[n].max(1, keepdims=True)
        mass = {}
        for al in itertools.product(range(V), repeat=T):
            lab, prev = [], None
            for s in al:
                if s != prev and s != V - 1:
                    lab.append(s)
                prev = s
            mass[tuple(lab)] = np.logaddexp(mass.get(tuple(lab), -np.inf), x[np.arange(T), list(al)].sum())
        lab = max(mass, key=mass.get)
        best[n, :len(lab)] = lab
        best_len[n] = len(lab)
    np.savez(os.path.join(HERE, 'ctc_beam.npz'), logits=logits, best=best, best_len=best_len)


if __name__ == '__main__':
    torch.manual_seed(0)
    golden_ctc()
    golden_blstm()
    golden_speller()
    golden_ctc_beam()
    print(sorted(os.listdir(HERE)))
