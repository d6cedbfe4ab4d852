"""DecoderEvaluator (reference: nabu/neuralnetworks/evaluators/decoder_evaluator.py:8-53): decode the
validation batch with the decoder named in the cfg's [decoder] section and let the decoder fold the
result into the running evaluation loss (e.g. label error rate for ctc_decoder / beam_search_decoder)."""
from . import evaluator
from ..decoders import decoder as decoder_mod
from ..decoders import decoder_factory


class DecoderEvaluator(evaluator.Evaluator):
    def __init__(self, conf, dataconf, model, batch_source=None):
        super(DecoderEvaluator, self).__init__(conf, dataconf, model, batch_source)
        self.decoder = decoder_factory.factory(conf.get('decoder', 'decoder'))(conf, model)

    def init_loss(self):
        running = decoder_mod.RunningLoss()
        return {'loss': 0.0, 'count': 0.0, 'running': running}

    def update_loss(self, loss, inputs, input_seq_length, targets, target_seq_length):
        outputs = self.decoder(inputs, input_seq_length)
        loss['loss'] = self.decoder.update_evaluation_loss(loss['running'], outputs, targets, target_seq_length)
        loss['count'] = loss['running'].num_targets
        return loss['loss']
