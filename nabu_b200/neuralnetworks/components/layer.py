"""Network layers (reference: nabu/neuralnetworks/components/layer.py)."""
from ... import engine
from . import ops

_CELL = 'layer_norm_basic_lstm_cell'   # TF scope of LayerNormBasicLSTMCell (layer.py:35-42)


def _blstm_vars(store, scope, D, H):
    """Declare the four variables of one BLSTM under `scope` (TF names).
    Kernel AND bias are glorot-uniform: get_variable's default, the cell passes no initializer."""
    out = []
    for d in ('fw', 'bw'):
        base = '%s/bidirectional_rnn/%s/%s' % (scope, d, _CELL)
        out.append(store.get(base + '/kernel', (D + H, 4 * H), 'glorot'))
        out.append(store.get(base + '/bias', (4 * H,), 'glorot'))
    return out


def declare_blstm(store, scope, input_dim, num_units):
    _blstm_vars(store, scope, input_dim, num_units)
    return 2 * num_units


def blstm(store, inputs, sequence_length, num_units, scope, pad_to=1, planes=None, want_planes=False):
    """layer.py:8-51: forward and backward LSTM over time, outputs concatenated (fw | bw).
    pad_to > 1 rounds the output time axis up to a multiple (zero frames) for pyramid_stack.
    `planes` / `want_planes`: the tensor-core operand planes of `inputs` from the layer that produced it, and whether to
    return (outputs, planes of outputs) -- see engine.blstm."""
    B, T, D = inputs.shape
    kf, bf, kb, bb = _blstm_vars(store, scope, D, num_units)
    yT = (T + pad_to - 1) // pad_to * pad_to
    return engine.blstm(inputs, sequence_length, kf, bf, kb, bb, num_units, yT, x_planes=planes, want_planes=want_planes)


def declare_pblstm(store, scope, input_dim, num_units, num_steps=2):
    return declare_blstm(store, scope + '/BLSTM', input_dim, num_units) * num_steps


def pblstm(store, inputs, sequence_length, num_units, num_steps=2, scope='PBLSTM', planes=None, want_planes=False):
    """layer.py:53-94: BLSTM followed by pyramid_stack (time /num_steps, features *num_steps).  The stacking is a
    reshape, of the outputs and of their operand planes alike."""
    out = blstm(store, inputs, sequence_length, num_units, scope + '/BLSTM', pad_to=num_steps, planes=planes,
                want_planes=want_planes)
    outputs, out_planes = out if want_planes else (out, None)
    stacked, lengths = ops.pyramid_stack(outputs, sequence_length, num_steps)
    return (stacked, lengths, out_planes) if want_planes else (stacked, lengths)
