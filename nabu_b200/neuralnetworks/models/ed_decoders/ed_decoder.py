"""EDDecoder base class (reference: nabu/neuralnetworks/models/ed_decoders/ed_decoder.py:9-136)."""
from abc import ABCMeta, abstractmethod

from ....tools.default_conf import apply_defaults, defaults_path


class EDDecoder(object, metaclass=ABCMeta):
    """Turns the encoder's representations into output logits."""

    def __init__(self, conf, output_dims, constraint, name=None):
        self.conf = dict(conf.items('decoder'))
        apply_defaults(self.conf, defaults_path(__file__, type(self)))
        self.outputs = list(output_dims.keys())
        self.output_dims = output_dims
        self.constraint = constraint
        self.scope = name or type(self).__name__
        self.store = None

    def __call__(self, encoded, encoded_seq_length, targets, target_seq_length, is_training):
        if self.store is None or not self.store.materialised:
            raise RuntimeError('%s: build the Model (Model.build) before calling the decoder' % self.scope)
        return self._decode(encoded, encoded_seq_length, targets, target_seq_length, is_training)

    @abstractmethod
    def declare(self, encoded_dims):
        """Declare the variables for {name: encoded feature dim}."""

    @abstractmethod
    def _decode(self, encoded, encoded_seq_length, targets, target_seq_length, is_training):
        """Returns (logits dict, logit sequence-length dict, final state)."""

    @abstractmethod
    def zero_state(self, encoded_dim, batch_size):
        """The decoder's zero state."""

    @property
    def variables(self):
        return [v for v in self.store.order if v.name.startswith(self.scope + '/')]
