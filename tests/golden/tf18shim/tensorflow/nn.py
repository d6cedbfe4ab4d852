"""tf.nn (r1.8): the ops the reference calls"""
import numpy as np
import torch

import tensorflow as tf


def softmax(logits, axis=-1, name=None, dim=None):
    return tf.Tensor(torch.softmax(tf._t(logits), int(dim if dim is not None else axis)))


def log_softmax(logits, axis=-1, name=None, dim=None):
    return tf.Tensor(torch.log_softmax(tf._t(logits), int(dim if dim is not None else axis)))


sigmoid, tanh = tf.sigmoid, tf.tanh


def relu(x, name=None):
    return tf.Tensor(torch.relu(tf._t(x)))


def dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
    raise NotImplementedError('tf18shim: the goldens are made with dropout = 1')


def sparse_softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
    """nn_ops.sparse_softmax_cross_entropy_with_logits: -log softmax(logits)[label], any leading shape"""
    lg, lab = tf._t(logits), tf._t(labels).long()
    assert lg.shape[:-1] == lab.shape, 'logits %r vs labels %r' % (tuple(lg.shape), tuple(lab.shape))
    return tf.Tensor(-torch.gather(torch.log_softmax(lg, -1), -1, lab.unsqueeze(-1)).squeeze(-1))


def top_k(input, k=1, sorted=True, name=None):        # noqa: A002
    """nn_ops.top_k: descending; of equal values the LOWER index comes first (core/kernels/topk_op.cc)"""
    t = tf._t(input)
    vals, idx = torch.sort(t, dim=-1, descending=True, stable=True)
    return tf.Tensor(vals[..., :int(k)]), tf.Tensor(idx[..., :int(k)].to(torch.int32))


def ctc_loss(labels, inputs, sequence_length, preprocess_collapse_repeated=False, ctc_merge_repeated=True,
             ignore_longer_outputs_than_inputs=False, time_major=True):
    """ctc_ops.ctc_loss (core/util/ctc/ctc_loss_calculator.h): -log p(labels | softmax(inputs)) per batch entry, blank =
    num_classes - 1; here through torch's CTC loss on log_softmax (differentiable, the working precision)"""
    assert not preprocess_collapse_repeated and ctc_merge_repeated
    x = tf._t(inputs)
    if not time_major:
        x = x.transpose(0, 1)
    T, B, V = x.shape
    idx, vals = labels.indices.t.long(), labels.values.t.long()
    targets = [vals[idx[:, 0] == b] for b in range(B)]
    lens = torch.tensor([len(t) for t in targets])
    loss = torch.nn.functional.ctc_loss(torch.log_softmax(x, -1), torch.cat(targets), tf._t(sequence_length).long(),
                                        lens, blank=V - 1, reduction='none', zero_infinity=False)
    return tf.Tensor(loss)


def ctc_beam_search_decoder(inputs, sequence_length, beam_width=100, top_paths=1, merge_repeated=True):
    """ctc_ops.ctc_beam_search_decoder - NOT restated here: delegated to oracle.ctc_beam_search (TF's C++ beam search,
    core/util/ctc/ctc_beam_search.h, is pinned by nothing in this repository); only the reference's wrapper around the
    call (time-major transpose, SparseTensor output, cast) is exercised by running it."""
    import oracle as O
    assert top_paths == 1 and merge_repeated
    x = tf._t(inputs).detach().numpy().astype(np.float32)                   # [T, B, V]
    n = tf._t(sequence_length).numpy()
    rows, vals, width = [], [], 0
    logp = []
    for b in range(x.shape[1]):
        ids, score = O.ctc_beam_search(x[:, b], int(n[b]), beam_width)
        rows += [[b, i] for i in range(len(ids))]
        vals += list(ids)
        width = max(width, len(ids))
        logp.append([score])
    sparse = tf.SparseTensor(np.array(rows, np.int64).reshape(-1, 2), np.array(vals, np.int64),
                             np.array([x.shape[1], width], np.int64))
    return [sparse], tf.Tensor(torch.tensor(logp))
