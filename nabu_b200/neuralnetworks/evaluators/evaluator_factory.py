"""cfg string -> Evaluator class (reference: nabu/neuralnetworks/evaluators/evaluator_factory.py)."""
import importlib

# cfg string -> (module, class).  Modules are imported on first use.
_CLASSES = {
    'decoder_evaluator': ('decoder_evaluator', 'DecoderEvaluator'),
    'loss_evaluator': ('loss_evaluator', 'LossEvaluator'),
}
_OUT_OF_SCOPE = ()


def factory(evaluator):
    entry = _CLASSES.get(evaluator)
    if entry is None:
        if evaluator in _OUT_OF_SCOPE:
            raise Exception('evaluator type %s is outside the B200 hot path (SURVEY.md section 8)' % evaluator)
        raise Exception('Undefined evaluator type: %s' % evaluator)
    module = importlib.import_module('.' + entry[0], __package__)
    return getattr(module, entry[1])
