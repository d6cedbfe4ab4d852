"""tf.contrib.rnn (python/ops/rnn_cell_impl.py and contrib/rnn/python/ops/rnn_cell.py, r1.8)"""
import collections

import torch

import tensorflow as tf
from tensorflow.layers import Layer
from tensorflow import nest_impl as nest

LSTMStateTuple = collections.namedtuple('LSTMStateTuple', ('c', 'h'))


def _zero_state(size, batch_size, dtype):
    """rnn_cell_impl._zero_state_tensors"""
    return nest.map_structure(lambda s: tf.zeros([int(batch_size)] + tf._ints(s), dtype), size) \
        if nest.is_sequence(size) else tf.zeros([int(batch_size)] + tf._ints(size), dtype)


class RNNCell(Layer):
    """rnn_cell_impl.RNNCell: __call__(inputs, state, scope=None) runs Layer.__call__ under the given scope, or under
    the layer's own (default-named or, with reuse, plainly named) scope"""

    def __init__(self, trainable=True, name=None, dtype=None, _reuse=None, **kwargs):
        Layer.__init__(self, name=name, _reuse=_reuse)

    def __call__(self, inputs, state, scope=None):
        return Layer.__call__(self, inputs, state, scope=scope)

    def zero_state(self, batch_size, dtype):
        return _zero_state(self.state_size, batch_size, dtype)


class _LSTMBase(RNNCell):
    _bias_initializer = 'zeros'

    def __init__(self, num_units, forget_bias=1.0, reuse=None, name=None, **kwargs):
        RNNCell.__init__(self, name=name, _reuse=reuse)
        self._num_units, self._forget_bias = int(num_units), float(forget_bias)

    @property
    def state_size(self):
        return LSTMStateTuple(self._num_units, self._num_units)

    @property
    def output_size(self):
        return self._num_units

    def build(self, _):
        self._kernel = self._bias = None

    def call(self, inputs, state):
        """gates in the order i, j, f, o: c' = c * sigmoid(f + forget_bias) + sigmoid(i) * tanh(j), h' = tanh(c') *
        sigmoid(o)  (LSTMCell.call without peepholes / projection; LayerNormBasicLSTMCell.call with layer_norm=False)"""
        c, h = state
        x = torch.cat([tf._t(inputs), tf._t(h)], 1)
        if self._kernel is None:
            self._kernel = tf.get_variable('kernel', [x.shape[1], 4 * self._num_units])
            self._bias = tf.get_variable('bias', [4 * self._num_units], initializer=self._bias_initializer)
        z = torch.matmul(x, self._kernel.t) + self._bias.t
        i, j, f, o = torch.chunk(z, 4, 1)
        new_c = tf._t(c) * torch.sigmoid(f + self._forget_bias) + torch.sigmoid(i) * torch.tanh(j)
        new_h = torch.tanh(new_c) * torch.sigmoid(o)
        return tf.Tensor(new_h), LSTMStateTuple(tf.Tensor(new_c), tf.Tensor(new_h))


class LSTMCell(_LSTMBase):
    """rnn_cell_impl.LSTMCell(num_units, use_peepholes=False, num_proj=None, forget_bias=1.0, state_is_tuple=True):
    variables `kernel` (the scope's default initializer) and `bias` (zeros)"""
    _bias_initializer = 'zeros'


class LayerNormBasicLSTMCell(_LSTMBase):
    """contrib/rnn/python/ops/rnn_cell.py: LayerNormBasicLSTMCell with layer_norm=False, dropout_keep_prob=1: _linear
    creates `kernel` AND `bias` with vs.get_variable and no initializer, i.e. the bias is glorot-uniform too"""
    _bias_initializer = None

    def __init__(self, num_units, forget_bias=1.0, layer_norm=True, dropout_keep_prob=1.0, reuse=None, **kwargs):
        assert not layer_norm and dropout_keep_prob == 1.0
        _LSTMBase.__init__(self, num_units, forget_bias, reuse)


class MultiRNNCell(RNNCell):
    """rnn_cell_impl.MultiRNNCell: cell i runs under variable_scope('cell_%d' % i), state = tuple of the cells' states"""

    def __init__(self, cells, state_is_tuple=True):
        RNNCell.__init__(self)
        self._cells = list(cells)

    @property
    def state_size(self):
        return tuple(c.state_size for c in self._cells)

    @property
    def output_size(self):
        return self._cells[-1].output_size

    def zero_state(self, batch_size, dtype):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    def call(self, inputs, state):
        new_states = []
        for i, cell in enumerate(self._cells):
            with tf.variable_scope('cell_%d' % i):
                inputs, s = cell(inputs, state[i])
                new_states.append(s)
        return inputs, tuple(new_states)


class DropoutWrapper(RNNCell):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('tf18shim: the goldens are made with dropout = 1')
