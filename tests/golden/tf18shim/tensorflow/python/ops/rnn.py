"""tensorflow/python/ops/rnn.py (r1.8): dynamic_rnn with sequence_length, bidirectional_dynamic_rnn"""
import torch

import tensorflow as tf
from tensorflow import nest_impl as nest


def dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, parallel_iterations=None,
                swap_memory=False, time_major=False, scope=None):
    """rnn.dynamic_rnn / _dynamic_rnn_loop / _rnn_step: past an entry's length the output is zero and the state is
    copied through"""
    assert not time_major
    with tf.variable_scope(scope or 'rnn'):
        x = tf._t(inputs)
        B, T = x.shape[0], x.shape[1]
        n = tf._t(sequence_length).long() if sequence_length is not None else torch.full([B], T)
        state = initial_state if initial_state is not None else cell.zero_state(B, dtype)
        outputs = []
        with tf._traced_once() as again:
            for t in range(T):
                again()
                out, new_state = cell(tf.Tensor(x[:, t]), state)
                live = (t < n).unsqueeze(1)
                outputs.append(torch.where(live, tf._t(out), torch.zeros_like(tf._t(out))))
                state = nest.map_structure(lambda new, old: tf.Tensor(torch.where(live, tf._t(new), tf._t(old))),
                                           new_state, state)
        return tf.Tensor(torch.stack(outputs, 1)), state


def bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, sequence_length=None, initial_state_fw=None,
                              initial_state_bw=None, dtype=None, parallel_iterations=None, swap_memory=False,
                              time_major=False, scope=None):
    """rnn.bidirectional_dynamic_rnn: scopes bidirectional_rnn/{fw,bw}; the backward direction sees
    reverse_sequence(inputs, sequence_length) and its outputs are reversed back the same way"""
    with tf.variable_scope(scope or 'bidirectional_rnn'):
        with tf.variable_scope('fw') as fw_scope:
            out_fw, state_fw = dynamic_rnn(cell_fw, inputs, sequence_length, initial_state_fw, dtype, scope=fw_scope)
        with tf.variable_scope('bw') as bw_scope:
            rev = tf.reverse_sequence(inputs, sequence_length, seq_axis=1, batch_axis=0)
            tmp, state_bw = dynamic_rnn(cell_bw, rev, sequence_length, initial_state_bw, dtype, scope=bw_scope)
        out_bw = tf.reverse_sequence(tmp, sequence_length, seq_axis=1, batch_axis=0)
    return (out_fw, out_bw), (state_fw, state_bw)
