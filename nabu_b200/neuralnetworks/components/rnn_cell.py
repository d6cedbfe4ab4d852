"""The Speller's cell (reference: components/rnn_cell.py:109-155 AttentionProjectionWrapper around
tf.contrib.seq2seq.AttentionWrapper(MultiRNNCell(LSTMCell...)), speller.py:29-69).

`AttentionProjectionCell` is a description, not a step function: the fused kernels run the whole
target sequence (training) or the whole beam search (decoding) in one C-ABI call."""
from ... import engine


class AttentionProjectionCell(object):

    def __init__(self, svars, memory, memory_seq_length, output_dim, num_units, num_layers, attention, numfilt,
                 filtersize, dropout_keep=1.0, sample_prob=0.0, seed=0):
        self.dropout_keep, self.sample_prob, self.seed = dropout_keep, sample_prob, seed
        self.svars, self.memory, self.memory_seq_length = svars, memory, memory_seq_length
        self.output_dim, self.num_units, self.num_layers = output_dim, num_units, num_layers
        self.attention, self.numfilt, self.filtersize = attention, numfilt, filtersize

    def teacher_forced(self, targets, target_seq_length):
        """logits [B, U, V] of rnn_decoder.py:40-82 (output dropout and scheduled sampling when the cell carries them)."""
        return engine.speller(self.memory, self.memory_seq_length, targets, target_seq_length, self.svars,
                              self.output_dim, self.num_units, self.num_layers, self.attention, self.numfilt,
                              self.filtersize, self.dropout_keep, self.sample_prob, self.seed)

    def beam_search(self, beam_width, max_steps, length_penalty, temperature):
        return engine.las_beam_search(self.memory, self.memory_seq_length, self.svars, self.output_dim,
                                      self.num_units, self.num_layers, self.attention, self.numfilt, self.filtersize,
                                      beam_width, max_steps, length_penalty, temperature)
