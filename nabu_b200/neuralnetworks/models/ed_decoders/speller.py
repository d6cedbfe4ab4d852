"""LAS speller (reference: .../ed_decoders/speller.py:10-69)."""
from . import rnn_decoder
from ...components import attention, rnn_cell
from .... import engine


class Speller(rnn_decoder.RNNDecoder):
    """MultiRNNCell(LSTMCell x num_layers) + Bahdanau / location-aware attention over the listener
    output + linear projection of [cell output, context] to the output classes."""

    def _vars(self, encoded_dim):
        V = list(self.output_dims.values())[0]
        H, NL = int(self.conf['num_units']), int(self.conf['num_layers'])
        att, numfilt, filtersize = attention.factory(self.conf)
        E, A = encoded_dim, H
        s, g = self.scope, self.store.get
        cell = s + '/decoder/attention_wrapper/multi_rnn_cell/cell_%d/lstm_cell/%s'
        kernels = [g(cell % (l, 'kernel'), ((V + E if l == 0 else H) + H, 4 * H), 'glorot') for l in range(NL)]
        biases = [g(cell % (l, 'bias'), (4 * H,), 'zeros') for l in range(NL)]
        mem = g(s + '/memory_layer/kernel', (E, A), 'glorot')
        att_scope = s + '/decoder/attention_wrapper/' + ('location_aware_attention' if att.startswith('location_aware')
                                                         else 'windowed_attention' if att.startswith('windowed')
                                                         else 'bahdanau_attention')
        qk = g(att_scope + '/query_layer/kernel', (H, A), 'glorot')
        v = g(att_scope + '/attention_v', (A,), 'glorot')
        ck = dk = None
        if att.startswith('location_aware'):
            ck = g(att_scope + '/conv1d/kernel', (filtersize, 1, numfilt), 'glorot')
            dk = g(att_scope + '/process_conv_features/kernel', (numfilt, A), 'glorot')
        ok = g(s + '/decoder/dense/kernel', (H + E, V), 'glorot')
        ob = g(s + '/decoder/dense/bias', (V,), 'zeros')
        return engine.SpellerVars(kernels, biases, mem, qk, v, ck, dk, ok, ob), (V, H, NL, att, numfilt, filtersize)

    def declare(self, encoded_dims):
        if len(encoded_dims) != 1:
            raise Exception('speller: exactly one encoded input is on the B200 hot path')
        self._vars(list(encoded_dims.values())[0])

    def create_cell(self, encoded, encoded_seq_length, is_training):
        name = list(encoded.keys())[0]
        memory = encoded[name]
        svars, (V, H, NL, att, numfilt, filtersize) = self._vars(memory.shape[-1])
        # DropoutWrapper(output_keep_prob = dropout) on every LSTMCell when training (speller.py:37-41); a fresh seed
        # per call for the counter generator of the kernels
        keep = float(self.conf['dropout']) if is_training else 1.0
        self._calls = getattr(self, '_calls', 0) + 1
        return rnn_cell.AttentionProjectionCell(svars, memory, encoded_seq_length[name], V, H, NL, att, numfilt,
                                                filtersize, dropout_keep=keep, seed=0x5EED0000 + self._calls)
