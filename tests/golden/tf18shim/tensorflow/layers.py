"""tf.layers (tensorflow/python/layers/base.py, core.py, convolutional.py, r1.8): the scope handling of Layer and the
two layers the reference calls."""
import re

import torch

import tensorflow as tf


def _snake(name):
    """base.py: _to_snake_case"""
    s = re.sub('(.)([A-Z][a-z0-9]+)', r'\1_\2', name)
    s = re.sub('([a-z])([A-Z])', r'\1_\2', s).lower()
    return 'private' + s if s[0] == '_' else s


class Layer(object):
    """base.py: Layer._set_scope / __call__ - the layer's variable scope is captured at the FIRST call (by name when the
    layer was given `_reuse`, through default-name uniquification otherwise) and re-entered afterwards."""

    def __init__(self, trainable=True, name=None, dtype=None, _scope=None, _reuse=None, **kwargs):
        self._base_name = name or _snake(type(self).__name__)
        self._given_scope, self._reuse, self._scope, self.built = _scope, _reuse, None, False

    def _enter_scope(self, scope=None):
        if self._scope is not None:
            return tf.variable_scope(self._scope, reuse=(tf.AUTO_REUSE if self.built else None))
        scope = scope if scope is not None else self._given_scope
        if self._reuse:
            return tf.variable_scope(scope if scope is not None else self._base_name)
        return tf.variable_scope(scope, default_name=self._base_name)

    def __call__(self, inputs, *args, **kwargs):
        scope = kwargs.pop('scope', None)
        with self._enter_scope(scope) as captured:
            if self._scope is None:
                self._scope = captured
            if not self.built:
                self.build(getattr(inputs, 'shape', None))
                self.built = True
            return self.call(inputs, *args, **kwargs)

    def build(self, input_shape):
        pass

    @property
    def scope_name(self):
        return self._scope.name


class Dense(Layer):
    """core.py: Dense - outputs = activation(inputs . kernel + bias) on the last axis"""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
                 name=None, names=('kernel', 'bias'), **kwargs):
        Layer.__init__(self, name=name, **kwargs)
        self.units, self.activation, self.use_bias = int(units), activation, use_bias
        self.kernel_initializer, self.bias_initializer, self._names = kernel_initializer, bias_initializer, names

    def build(self, input_shape):
        depth = int(input_shape[-1])
        self.kernel = tf.get_variable(self._names[0], [depth, self.units], initializer=self.kernel_initializer)
        self.bias = tf.get_variable(self._names[1], [self.units],
                                    initializer=self.bias_initializer or tf.zeros_initializer()) \
            if self.use_bias else None

    def call(self, inputs):
        out = torch.matmul(tf._t(inputs), self.kernel.t)
        if self.bias is not None:
            out = out + self.bias.t
        out = tf.Tensor(out)
        return self.activation(out) if self.activation is not None else out


def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name=None,
          reuse=None, **kwargs):
    """core.py: dense - a Dense built with _scope=name, _reuse=reuse and called once"""
    layer = Dense(units, activation, use_bias, kernel_initializer, bias_initializer, name=name, _scope=name,
                  _reuse=reuse)
    return layer(inputs)


class Conv1D(Layer):
    """convolutional.py: Conv1D, channels_last, stride 1, no dilation; kernel [width, in, filters]; padding 'same' puts
    (width - 1) // 2 zeros in front and the rest behind (nn_ops convolution, SAME); cross-correlation"""

    def __init__(self, filters, kernel_size, padding='valid', use_bias=True, name=None, **kwargs):
        Layer.__init__(self, name=name, **kwargs)
        self.filters, self.width, self.padding, self.use_bias = int(filters), int(kernel_size), padding, use_bias

    def build(self, input_shape):
        self.kernel = tf.get_variable('kernel', [self.width, int(input_shape[-1]), self.filters])
        self.bias = tf.get_variable('bias', [self.filters], initializer=tf.zeros_initializer()) \
            if self.use_bias else None

    def call(self, inputs):
        x = tf._t(inputs).transpose(1, 2)                                   # [B, in, T]
        if self.padding.lower() == 'same':
            total = self.width - 1
            x = torch.nn.functional.pad(x, (total // 2, total - total // 2))
        out = torch.nn.functional.conv1d(x, self.kernel.t.permute(2, 1, 0))  # [filters, in, width]
        out = out.transpose(1, 2)
        if self.bias is not None:
            out = out + self.bias.t
        return tf.Tensor(out)


def conv1d(inputs, filters, kernel_size, strides=1, padding='valid', use_bias=True, name=None, reuse=None, **kwargs):
    assert strides == 1
    return Conv1D(filters, kernel_size, padding, use_bias, name=name, _scope=name, _reuse=reuse)(inputs)
