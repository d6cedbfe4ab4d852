"""tf.contrib.layers (contrib/layers/python/layers/layers.py, r1.8)"""
import tensorflow as tf
from tensorflow.layers import Dense


def fully_connected(inputs, num_outputs, activation_fn=tf.nn.relu, weights_initializer=None, biases_initializer='zeros',
                    reuse=None, scope=None, **kwargs):
    """layers.fully_connected: variable_scope(scope, 'fully_connected'), a core Dense whose variables are renamed
    kernel -> weights, bias -> biases; xavier (= glorot uniform) weights, zero biases"""
    with tf.variable_scope(scope, default_name='fully_connected', reuse=reuse) as sc:
        layer = Dense(num_outputs, None, True, weights_initializer, biases_initializer, names=('weights', 'biases'),
                      _scope=sc, _reuse=True)
        layer._scope = sc
        out = layer(inputs)
    return activation_fn(out) if activation_fn is not None else out


def linear(inputs, num_outputs, **kwargs):
    return fully_connected(inputs, num_outputs, activation_fn=None, **kwargs)


def layer_norm(*args, **kwargs):
    raise NotImplementedError('tf18shim: layer_norm (the pinned recipes do not use it)')
