"""cfg string -> EDEncoder class (reference: .../ed_encoders/ed_encoder_factory.py:4-29)."""
import importlib

# cfg string -> (module, class).  Modules are imported on first use.
_CLASSES = {
    'listener': ('listener', 'Listener'),
    'dblstm': ('dblstm', 'DBLSTM'),
}
_OUT_OF_SCOPE = ('dummy_encoder', 'dnn', 'hotstart_encoder')


def factory(encoder):
    entry = _CLASSES.get(encoder)
    if entry is None:
        if encoder in _OUT_OF_SCOPE:
            raise Exception('encoder type %s is outside the B200 hot path (SURVEY.md section 8)' % encoder)
        raise Exception('undefined encoder type: %s' % encoder)
    module = importlib.import_module('.' + entry[0], __package__)
    return getattr(module, entry[1])
