"""EDEncoder base class (reference: nabu/neuralnetworks/models/ed_encoders/ed_encoder.py:9-96)."""
from abc import ABCMeta, abstractmethod

from ....tools.default_conf import apply_defaults, defaults_path


class EDEncoder(object, metaclass=ABCMeta):
    """Turns {name: [B,T,dim]} features into {name: [B,T',dim']} high level representations."""

    def __init__(self, conf, constraint, name=None):
        self.conf = dict(conf.items('encoder'))
        apply_defaults(self.conf, defaults_path(__file__, type(self)))
        self.constraint = constraint
        self.scope = name or type(self).__name__
        self.store = None          # engine.ParamStore, attached by Model

    def __call__(self, inputs, input_seq_length, is_training):
        if self.store is None or not self.store.materialised:
            raise RuntimeError('%s: build the Model (Model.build) before calling the encoder' % self.scope)
        return self.encode(inputs, input_seq_length, is_training)

    @abstractmethod
    def declare(self, input_dims):
        """Declare the variables for {name: feature dim}; returns {name: output dim}."""

    @abstractmethod
    def encode(self, inputs, input_seq_length, is_training):
        """Returns (encoded dict, encoded sequence-length dict)."""

    @property
    def variables(self):
        return [v for v in self.store.order if v.name.startswith(self.scope + '/')]
