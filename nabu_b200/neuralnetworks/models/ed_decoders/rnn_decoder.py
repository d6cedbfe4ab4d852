"""Recurrent decoder base (reference: .../ed_decoders/rnn_decoder.py:13-122).

The reference builds BasicDecoder(ScheduledEmbeddingTrainingHelper) around `create_cell` and runs
dynamic_decode(impute_finished=True).  Here `create_cell` returns a cell *description* (the variables
plus the attention memory) and the whole teacher-forced sequence is one fused call."""
from abc import ABCMeta, abstractmethod

from . import ed_decoder


class RNNDecoder(ed_decoder.EDDecoder, metaclass=ABCMeta):

    def _decode(self, encoded, encoded_seq_length, targets, target_seq_length, is_training):
        output_name = list(self.output_dims.keys())[0]
        cell = self.create_cell(encoded, encoded_seq_length, is_training)
        if is_training:                       # ScheduledEmbeddingTrainingHelper(sampling_probability) :59-64
            cell.sample_prob = float(self.conf['sample_prob'])
        tgt = list(targets.values())[0]
        tgt_len = list(target_seq_length.values())[0]
        logits = cell.teacher_forced(tgt, tgt_len)
        return {output_name: logits}, {output_name: tgt_len}, None

    @abstractmethod
    def create_cell(self, encoded, encoded_seq_length, is_training):
        """the decoder cell bound to the attention memory"""

    def zero_state(self, encoded_dim, batch_size):
        return None
