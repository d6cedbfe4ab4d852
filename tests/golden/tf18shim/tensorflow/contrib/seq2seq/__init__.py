"""tf.contrib.seq2seq (contrib/seq2seq/python/ops/{attention_wrapper,basic_decoder,helper,decoder,beam_search_decoder}.py,
r1.8): what nabu's Speller, its attention mechanisms and its beam search stand on."""
import collections

import torch

import tensorflow as tf
from tensorflow import nest_impl as nest
from tensorflow.contrib.rnn import RNNCell
from tensorflow.layers import Dense


# ------------------------------------------------------------------------------------------------ attention mechanisms
def _prepare_memory(memory, memory_sequence_length):
    """attention_wrapper._prepare_memory: memory rows past an entry's length are zeroed"""
    m = tf._t(memory)
    if memory_sequence_length is None:
        return tf.Tensor(m)
    mask = tf._t(tf.sequence_mask(memory_sequence_length, m.shape[1], tf.float32))
    return tf.Tensor(m * mask.reshape(list(mask.shape) + [1] * (m.dim() - 2)))


def _maybe_mask_score(score, memory_sequence_length, score_mask_value):
    """attention_wrapper._maybe_mask_score: where(sequence_mask, score, score_mask_value)"""
    if memory_sequence_length is None:
        return score
    s = tf._t(score)
    mask = tf._t(tf.sequence_mask(memory_sequence_length, s.shape[1]))
    return tf.Tensor(torch.where(mask, s, torch.full_like(s, score_mask_value)))


class AttentionMechanism(object):
    pass


class _BaseAttentionMechanism(AttentionMechanism):
    """attention_wrapper._BaseAttentionMechanism.__init__: values = masked memory, keys = memory_layer(values) (built
    HERE, under the scope current at construction), probability_fn wrapped with the score mask"""

    def __init__(self, query_layer, memory, probability_fn, memory_sequence_length=None, memory_layer=None,
                 check_inner_dims_defined=True, score_mask_value=None, name=None):
        self._query_layer, self._memory_layer = query_layer, memory_layer
        if score_mask_value is None:
            score_mask_value = float('-inf')
        self._probability_fn = lambda score, prev: probability_fn(
            _maybe_mask_score(score, memory_sequence_length, score_mask_value), prev)
        self._values = _prepare_memory(memory, memory_sequence_length)
        self._keys = self.memory_layer(self._values) if self.memory_layer else self._values
        self._batch_size = self._keys.shape[0].value
        self._alignments_size = self._keys.shape[1].value

    memory_layer = property(lambda self: self._memory_layer)
    query_layer = property(lambda self: self._query_layer)
    values = property(lambda self: self._values)
    keys = property(lambda self: self._keys)
    batch_size = property(lambda self: self._batch_size)
    alignments_size = property(lambda self: self._alignments_size)
    state_size = property(lambda self: self._alignments_size)

    def initial_alignments(self, batch_size, dtype):
        return tf.zeros([int(batch_size), self._alignments_size], dtype)

    def initial_state(self, batch_size, dtype):
        return self.initial_alignments(batch_size, dtype)


def _bahdanau_score(processed_query, keys, normalize):
    """attention_wrapper._bahdanau_score: sum_k v_k tanh(keys + query)"""
    assert not normalize
    num_units = keys.shape[2].value
    v = tf.get_variable('attention_v', [num_units])
    return tf.reduce_sum(v * tf.tanh(keys + tf.expand_dims(processed_query, 1)), [2])


class BahdanauAttention(_BaseAttentionMechanism):
    """attention_wrapper.BahdanauAttention: query_layer / memory_layer = Dense(num_units, use_bias=False) named
    'query_layer' / 'memory_layer'; probability_fn defaults to softmax and is called with the score alone"""

    def __init__(self, num_units, memory, memory_sequence_length=None, normalize=False, probability_fn=None,
                 score_mask_value=None, dtype=None, name='BahdanauAttention'):
        if probability_fn is None:
            probability_fn = tf.nn.softmax
        _BaseAttentionMechanism.__init__(
            self, query_layer=Dense(num_units, name='query_layer', use_bias=False),
            memory_layer=Dense(num_units, name='memory_layer', use_bias=False), memory=memory,
            probability_fn=lambda score, _: probability_fn(score), memory_sequence_length=memory_sequence_length,
            score_mask_value=score_mask_value, name=name)
        self._num_units, self._normalize, self._name = num_units, normalize, name

    def __call__(self, query, state):
        with tf.variable_scope(None, 'bahdanau_attention', [query]):
            processed_query = self.query_layer(query) if self.query_layer else query
            score = _bahdanau_score(processed_query, self._keys, self._normalize)
        alignments = self._probability_fn(score, state)
        return alignments, alignments


def hardmax(logits, name=None):
    raise NotImplementedError('tf18shim: hardmax (named in the reference\'s doc strings only)')


class AttentionWrapperState(collections.namedtuple(
        'AttentionWrapperState',
        ('cell_state', 'attention', 'time', 'alignments', 'alignment_history', 'attention_state'))):
    def clone(self, **kwargs):
        return self._replace(**kwargs)


def _compute_attention(mechanism, cell_output, attention_state, attention_layer):
    """attention_wrapper._compute_attention: context = alignments . values"""
    alignments, next_attention_state = mechanism(cell_output, state=attention_state)
    context = torch.matmul(tf._t(alignments).unsqueeze(1), tf._t(mechanism.values)).squeeze(1)
    context = tf.Tensor(context)
    if attention_layer is not None:
        attention = attention_layer(tf.concat([cell_output, context], 1))
    else:
        attention = context
    return attention, alignments, next_attention_state


class AttentionWrapper(RNNCell):
    """attention_wrapper.AttentionWrapper.call, without an attention layer and without alignment history:
    cell_inputs = concat([inputs, previous attention]); the cell; one attention step per mechanism on the cell output;
    attention = concat of the contexts; the OUTPUT is the cell output when output_attention is False"""

    def __init__(self, cell, attention_mechanism, attention_layer_size=None, alignment_history=False,
                 cell_input_fn=None, output_attention=True, initial_cell_state=None, name=None):
        RNNCell.__init__(self, name=name)
        assert attention_layer_size is None and not alignment_history and cell_input_fn is None
        self._is_multi = isinstance(attention_mechanism, (list, tuple))
        self._mechanisms = list(attention_mechanism) if self._is_multi else [attention_mechanism]
        self._cell, self._output_attention = cell, output_attention
        self._attention_layer_size = sum(m.values.shape[-1].value for m in self._mechanisms)

    def _item_or_tuple(self, seq):
        return tuple(seq) if self._is_multi else list(seq)[0]

    @property
    def output_size(self):
        return self._attention_layer_size if self._output_attention else self._cell.output_size

    @property
    def state_size(self):
        return AttentionWrapperState(
            cell_state=self._cell.state_size, time=[], attention=self._attention_layer_size,
            alignments=self._item_or_tuple(m.alignments_size for m in self._mechanisms),
            attention_state=self._item_or_tuple(m.state_size for m in self._mechanisms),
            alignment_history=self._item_or_tuple(() for _ in self._mechanisms))

    def zero_state(self, batch_size, dtype):
        return AttentionWrapperState(
            cell_state=self._cell.zero_state(batch_size, dtype), time=tf.zeros([], tf.int32),
            attention=tf.zeros([int(batch_size), self._attention_layer_size], dtype),
            alignments=self._item_or_tuple(m.initial_alignments(batch_size, dtype) for m in self._mechanisms),
            attention_state=self._item_or_tuple(m.initial_state(batch_size, dtype) for m in self._mechanisms),
            alignment_history=self._item_or_tuple(() for _ in self._mechanisms))

    def call(self, inputs, state):
        cell_inputs = tf.concat([inputs, state.attention], -1)
        cell_output, next_cell_state = self._cell(cell_inputs, state.cell_state)
        previous = state.attention_state if self._is_multi else [state.attention_state]
        all_alignments, all_attentions, all_states = [], [], []
        for i, mechanism in enumerate(self._mechanisms):
            attention, alignments, next_attention_state = _compute_attention(mechanism, cell_output, previous[i], None)
            all_alignments.append(alignments)
            all_attentions.append(attention)
            all_states.append(next_attention_state)
        attention = tf.concat(all_attentions, 1)
        next_state = AttentionWrapperState(
            time=state.time + 1, cell_state=next_cell_state, attention=attention,
            attention_state=self._item_or_tuple(all_states), alignments=self._item_or_tuple(all_alignments),
            alignment_history=self._item_or_tuple(() for _ in self._mechanisms))
        return (attention if self._output_attention else cell_output), next_state


# -------------------------------------------------------------------------------------------------- decoders, helpers
class Decoder(object):
    """decoder.Decoder: the interface dynamic_decode drives"""

    def finalize(self, outputs, final_state, sequence_lengths):
        raise NotImplementedError


class BasicDecoderOutput(collections.namedtuple('BasicDecoderOutput', ('rnn_output', 'sample_id'))):
    pass


class TrainingHelper(object):
    """helper.TrainingHelper: feeds inputs[:, t]; finished when t + 1 >= sequence_length; zeros once all are finished"""

    def __init__(self, inputs, sequence_length, time_major=False, name=None):
        assert not time_major
        self._inputs, self._sequence_length = tf._t(inputs), tf._t(sequence_length).long()
        self._batch_size = self._inputs.shape[0]

    batch_size = property(lambda self: self._batch_size)

    def _read(self, time, finished):
        if bool(finished.all()):
            return tf.Tensor(torch.zeros_like(self._inputs[:, 0]))
        return tf.Tensor(self._inputs[:, time])

    def initialize(self, name=None):
        finished = self._sequence_length == 0
        return tf.Tensor(finished), self._read(0, finished)

    def sample(self, time, outputs, state, name=None):
        return tf.Tensor(torch.argmax(tf._t(outputs), -1).to(torch.int32))

    def next_inputs(self, time, outputs, state, sample_ids, name=None):
        next_time = int(time) + 1
        finished = next_time >= self._sequence_length
        return tf.Tensor(finished), self._read(next_time, finished), state


class ScheduledEmbeddingTrainingHelper(TrainingHelper):
    """helper.ScheduledEmbeddingTrainingHelper with sampling_probability = 0: bernoulli(0) never selects a sample, the
    sample ids are all -1 and next_inputs is TrainingHelper's"""

    def __init__(self, inputs, sequence_length, embedding, sampling_probability, time_major=False, seed=None,
                 scheduling_seed=None, name=None):
        assert float(sampling_probability) == 0.0, 'tf18shim: the goldens are made with sample_prob = 0'
        TrainingHelper.__init__(self, inputs, sequence_length, time_major)

    def sample(self, time, outputs, state, name=None):
        return tf.Tensor(torch.full([self._batch_size], -1, dtype=torch.int32))


class BasicDecoder(Decoder):
    """basic_decoder.BasicDecoder"""

    def __init__(self, cell, helper, initial_state, output_layer=None):
        assert output_layer is None
        self._cell, self._helper, self._initial_state = cell, helper, initial_state

    batch_size = property(lambda self: self._helper.batch_size)

    def initialize(self, name=None):
        return self._helper.initialize() + (self._initial_state,)

    def step(self, time, inputs, state, name=None):
        cell_outputs, cell_state = self._cell(inputs, state)
        sample_ids = self._helper.sample(time=time, outputs=cell_outputs, state=cell_state)
        finished, next_inputs, next_state = self._helper.next_inputs(
            time=time, outputs=cell_outputs, state=cell_state, sample_ids=sample_ids)
        return BasicDecoderOutput(cell_outputs, sample_ids), next_state, next_inputs, finished


def _transpose_batch_time(x):
    t = tf._t(x)
    return x if t.dim() < 2 else tf.Tensor(t.transpose(0, 1))


def dynamic_decode(decoder, output_time_major=False, impute_finished=False, maximum_iterations=None,
                   parallel_iterations=32, swap_memory=False, scope=None):
    """decoder.dynamic_decode (r1.8): variable_scope(scope, 'decoder'); step until every entry is finished;
    `finished` is sticky (logical_or with the previous value); sequence_lengths record the step at which an entry
    finished; with impute_finished the outputs of finished entries are zero and their state is copied through;
    finalize() is tried on the stacked (time-major) outputs, then every output is transposed to batch-major"""
    with tf.variable_scope(scope, 'decoder'):
        finished, inputs, state = decoder.initialize()
        finished = tf._t(finished)
        if maximum_iterations is not None:
            finished = finished | (0 >= int(maximum_iterations))
        lengths = torch.zeros_like(finished, dtype=torch.int32)
        time, collected = 0, []
        with tf._traced_once() as again:
            while not bool(finished.all()):
                again()
                outputs, decoder_state, next_inputs, decoder_finished = decoder.step(time, inputs, state)
                next_finished = tf._t(decoder_finished) | finished
                if maximum_iterations is not None:
                    next_finished = next_finished | (time + 1 >= int(maximum_iterations))
                lengths = torch.where(~finished & next_finished, torch.full_like(lengths, time + 1), lengths)
                if impute_finished:
                    def zero_out(out, fin=finished):
                        o = tf._t(out)
                        return tf.Tensor(torch.where(fin.reshape([-1] + [1] * (o.dim() - 1)), torch.zeros_like(o), o))

                    def keep(new, old, fin=finished):
                        if isinstance(new, tf.TensorArray) or tf._t(new).dim() == 0:
                            return new
                        n = tf._t(new)
                        return tf.Tensor(torch.where(fin.reshape([-1] + [1] * (n.dim() - 1)), tf._t(old), n))
                    outputs = nest.map_structure(zero_out, outputs)
                    decoder_state = nest.map_structure(keep, decoder_state, state)
                collected.append(outputs)
                time, inputs, state, finished = time + 1, next_inputs, decoder_state, next_finished
        final_outputs = nest.map_structure(lambda *steps: tf.stack(list(steps), 0), *collected)
        final_lengths = tf.Tensor(lengths)
        try:
            final_outputs, state = decoder.finalize(final_outputs, state, final_lengths)
        except NotImplementedError:
            pass
        if not output_time_major:
            final_outputs = nest.map_structure(_transpose_batch_time, final_outputs)
    return final_outputs, state, final_lengths


def tile_batch(t, multiplier, name=None):
    """beam_search_decoder.tile_batch: every batch entry repeated `multiplier` times, consecutively"""
    return nest.map_structure(lambda x: tf.Tensor(torch.repeat_interleave(tf._t(x), int(multiplier), 0)), t)
