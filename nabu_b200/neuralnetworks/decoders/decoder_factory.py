"""cfg string -> Decoder class (reference: nabu/neuralnetworks/decoders/decoder_factory.py:4-37)."""


def factory(decoder):
    if decoder == 'ctc_decoder':
        from . import ctc_decoder
        return ctc_decoder.CTCDecoder
    if decoder == 'beam_search_decoder':
        from . import beam_search_decoder
        return beam_search_decoder.BeamSearchDecoder
    if decoder in ('max_decoder', 'threshold_decoder', 'feature_decoder', 'alignment_decoder', 'random_decoder'):
        raise Exception('decoder type %s is outside the B200 hot path (SURVEY.md section 8)' % decoder)
    raise Exception('Undefined decoder type: %s' % decoder)
