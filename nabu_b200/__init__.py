"""nabu_b200: B200-native engine for nabu's per-utterance training / decoding hot path.

The package mirrors the reference's plugin API (Model, EDEncoder, EDDecoder, Trainer, loss
functions, Decoder -- selected by the same cfg strings) on top of hand-written sm_100a kernels
reached through the C-ABI in include/nabu_b200.h.  There is no CPU fallback.
"""
__version__ = '0.1.0'
