"""cfg string -> EDDecoder class (reference: .../ed_decoders/ed_decoder_factory.py:4-24)."""


def factory(decoder):
    if decoder == 'speller':
        from . import speller
        return speller.Speller
    if decoder == 'dnn_decoder':
        from . import dnn_decoder
        return dnn_decoder.DNNDecoder
    if decoder == 'hotstart_decoder':
        raise Exception('decoder type %s is outside the B200 hot path (SURVEY.md section 8)' % decoder)
    raise Exception('undefined decoder type: %s' % decoder)
