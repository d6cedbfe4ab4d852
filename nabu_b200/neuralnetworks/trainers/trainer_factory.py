"""cfg string -> Trainer class (reference: nabu/neuralnetworks/trainers/trainer_factory.py:4-17)."""
import importlib

# cfg string -> (module, class).  Modules are imported on first use.
_CLASSES = {
    'standard': ('standard_trainer', 'StandardTrainer'),
}
_OUT_OF_SCOPE = ()


def factory(trainer):
    entry = _CLASSES.get(trainer)
    if entry is None:
        if trainer in _OUT_OF_SCOPE:
            raise Exception('trainer type %s is outside the B200 hot path (SURVEY.md section 8)' % trainer)
        raise Exception('Undefined trainer type: %s' % trainer)
    module = importlib.import_module('.' + entry[0], __package__)
    return getattr(module, entry[1])
