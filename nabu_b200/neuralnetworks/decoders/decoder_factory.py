"""cfg string -> Decoder class (reference: nabu/neuralnetworks/decoders/decoder_factory.py:4-37)."""
import importlib

# cfg string -> (module, class).  Modules are imported on first use.
_CLASSES = {
    'ctc_decoder': ('ctc_decoder', 'CTCDecoder'),
    'beam_search_decoder': ('beam_search_decoder', 'BeamSearchDecoder'),
}
_OUT_OF_SCOPE = ('max_decoder', 'threshold_decoder', 'feature_decoder', 'alignment_decoder', 'random_decoder')


def factory(decoder):
    entry = _CLASSES.get(decoder)
    if entry is None:
        if decoder in _OUT_OF_SCOPE:
            raise Exception('decoder type %s is outside the B200 hot path (SURVEY.md section 8)' % decoder)
        raise Exception('Undefined decoder type: %s' % decoder)
    module = importlib.import_module('.' + entry[0], __package__)
    return getattr(module, entry[1])
