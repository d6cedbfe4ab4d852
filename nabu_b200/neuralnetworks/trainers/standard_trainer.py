"""StandardTrainer (reference: nabu/neuralnetworks/trainers/standard_trainer.py:6-41)."""
from . import trainer


class StandardTrainer(trainer.Trainer):
    """the plain trainer: no extra loss, no hooks"""

    def aditional_loss(self):
        return None

    def chief_only_hooks(self, outputs):
        return []

    def hooks(self, outputs):
        return []
