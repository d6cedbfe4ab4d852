"""cfg string -> EDDecoder class (reference: .../ed_decoders/ed_decoder_factory.py:4-24)."""
import importlib

# cfg string -> (module, class).  Modules are imported on first use.
_CLASSES = {
    'speller': ('speller', 'Speller'),
    'dnn_decoder': ('dnn_decoder', 'DNNDecoder'),
}
_OUT_OF_SCOPE = ('hotstart_decoder',)


def factory(decoder):
    entry = _CLASSES.get(decoder)
    if entry is None:
        if decoder in _OUT_OF_SCOPE:
            raise Exception('decoder type %s is outside the B200 hot path (SURVEY.md section 8)' % decoder)
        raise Exception('undefined decoder type: %s' % decoder)
    module = importlib.import_module('.' + entry[0], __package__)
    return getattr(module, entry[1])
