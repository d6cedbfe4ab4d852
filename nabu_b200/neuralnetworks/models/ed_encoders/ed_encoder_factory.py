"""cfg string -> EDEncoder class (reference: .../ed_encoders/ed_encoder_factory.py:4-29)."""


def factory(encoder):
    if encoder == 'listener':
        from . import listener
        return listener.Listener
    if encoder == 'dblstm':
        from . import dblstm
        return dblstm.DBLSTM
    if encoder in ('dummy_encoder', 'dnn', 'hotstart_encoder'):
        raise Exception('encoder type %s is outside the B200 hot path (SURVEY.md section 8)' % encoder)
    raise Exception('undefined encoder type: %s' % encoder)
