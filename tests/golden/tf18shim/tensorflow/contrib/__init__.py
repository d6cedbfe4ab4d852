from . import framework, rnn, seq2seq, layers, sparsemax  # noqa: F401
