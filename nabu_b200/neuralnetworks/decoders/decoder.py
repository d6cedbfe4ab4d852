"""Decoder base class (reference: nabu/neuralnetworks/decoders/decoder.py:8-69)."""
import os
from abc import ABCMeta, abstractmethod

from ...tools.default_conf import apply_defaults


class Decoder(object, metaclass=ABCMeta):
    """Decoder(conf, model): inference-time wrapper around a trained Model."""

    def __init__(self, conf, model):
        self.conf = dict(conf.items('decoder'))
        apply_defaults(self.conf, os.path.join(os.path.dirname(os.path.realpath(__file__)), 'defaults',
                                               type(self).__name__.lower() + '.cfg'))
        self.model = model

    @abstractmethod
    def __call__(self, inputs, input_seq_length):
        """decode a batch: {name: [B,T,dim]}, {name: [B]} -> outputs dict"""

    @abstractmethod
    def write(self, outputs, directory, names):
        """write the decoded batch to `directory`"""

    @abstractmethod
    def update_evaluation_loss(self, loss, outputs, references, reference_seq_length):
        """fold this batch into the running evaluation loss"""


def edit_distance(a, b):
    """tf.edit_distance(normalize=False) for one hypothesis / reference pair (host side)."""
    a, b = list(a), list(b)
    d = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        prev, d[0] = d[0], i
        for j in range(1, len(b) + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (a[i - 1] != b[j - 1]))
            prev = cur
    return d[len(b)]


class RunningLoss(object):
    """The reference's `loss` + `num_targets` variable pair (ctc_decoder.py:94-135): a running mean of
    errors per reference label."""

    def __init__(self):
        self.loss, self.num_targets = 0.0, 0.0

    def update(self, errors, batch_targets):
        new = self.num_targets + float(batch_targets)
        self.loss = (self.loss * self.num_targets + float(errors)) / new
        self.num_targets = new
        return self.loss
