"""cfg string -> Trainer class (reference: nabu/neuralnetworks/trainers/trainer_factory.py:4-17)."""


def factory(trainer):
    if trainer == 'standard':
        from . import standard_trainer
        return standard_trainer.StandardTrainer
    raise Exception('Undefined trainer type: %s' % trainer)
