"""Evaluator base class (reference: nabu/neuralnetworks/evaluators/evaluator.py:9-123).

The reference builds a validation input pipeline from the database sections named in the evaluator
cfg and returns `(loss variable, update_loss op, numbatches)`; the trainer re-initialises the loss and
runs `update_loss` numbatches times (trainers/trainer.py:657-665).  Here the data come from a batch
source (an iterable of `(inputs, input_seq_length, targets, target_seq_length)` dict tuples -- the
TFRecord pipeline is row f1) and `evaluate()` runs that loop: same running mean, same batch count
(the reference drops the tail: numbatches = len(data) // batch_size, evaluator.py:83-86; a batch
source is expected to hold whole batches already).
"""
import os
from abc import ABCMeta, abstractmethod

from ...tools.default_conf import apply_defaults


class Evaluator(object, metaclass=ABCMeta):
    def __init__(self, conf, dataconf, model, batch_source=None):
        self.conf = dict(conf.items('evaluator'))
        apply_defaults(self.conf, os.path.join(os.path.dirname(os.path.realpath(__file__)), 'defaults',
                                               type(self).__name__.lower() + '.cfg'))
        self.model = model
        self.dataconf = dataconf
        self.batch_source = batch_source
        targets = self.conf['targets'].split(' ')
        self.target_names = [] if targets == [''] else targets

    def init_loss(self):
        """the `loss` (and companion counter) variables of evaluator.py:68-76 at their initial value"""
        return {'loss': 0.0, 'count': 0.0}

    def source(self):
        """the batch source: the one given, else the database sections the evaluator cfg names (evaluator.py:78-104:
        one bucket, whole batches only)"""
        if self.batch_source is None:
            if self.dataconf is None:
                raise Exception('Evaluator.evaluate needs a batch_source or a database configuration')
            from ...processing import input_pipeline
            self.batch_source = input_pipeline.source_from_conf(
                self.conf, self.dataconf, self.model.input_names, self.target_names,
                device=getattr(self.model, 'device', 'cuda'))
        return self.batch_source

    def evaluate(self):
        """Returns (validation loss, number of batches): init_validation + numbatches x update_loss."""
        state = self.init_loss()
        numbatches = 0
        for batch in self.source():
            self.update_loss(state, *batch)
            numbatches += 1
        return state['loss'], numbatches

    @abstractmethod
    def update_loss(self, loss, inputs, input_seq_length, targets, target_seq_length):
        """fold one batch into the running validation loss (`loss` is the dict from init_loss)"""
