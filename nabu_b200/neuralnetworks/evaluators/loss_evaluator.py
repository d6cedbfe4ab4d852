"""LossEvaluator (reference: nabu/neuralnetworks/evaluators/loss_evaluator.py:8-64): the validation loss
is the utterance-weighted running mean of the training loss function on `is_training=False` logits."""
import torch

from . import evaluator
from ..trainers import loss_functions


class LossEvaluator(evaluator.Evaluator):
    def update_loss(self, loss, inputs, input_seq_length, targets, target_seq_length):
        with torch.no_grad():
            logits, logit_seq_length = self.model(inputs, input_seq_length, targets, target_seq_length, False)
            batch_loss = float(loss_functions.factory(self.conf['loss'])(targets, logits, logit_seq_length,
                                                                        target_seq_length))
        batch_utt = float(list(logits.values())[0].shape[0])
        new_num = loss['count'] + batch_utt
        # loss.assign((loss*num_utt + batch_loss*batch_utt)/new_num_utt)  (loss_evaluator.py:54-57)
        loss['loss'] = (loss['loss'] * loss['count'] + batch_loss * batch_utt) / new_num
        loss['count'] = new_num
        return loss['loss']
