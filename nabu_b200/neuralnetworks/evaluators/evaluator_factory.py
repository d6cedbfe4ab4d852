"""cfg string -> Evaluator class (reference: nabu/neuralnetworks/evaluators/evaluator_factory.py)."""


def factory(evaluator):
    if evaluator == 'decoder_evaluator':
        from . import decoder_evaluator
        return decoder_evaluator.DecoderEvaluator
    if evaluator == 'loss_evaluator':
        from . import loss_evaluator
        return loss_evaluator.LossEvaluator
    raise Exception('Undefined evaluator type: %s' % evaluator)
