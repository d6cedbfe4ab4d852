def sparsemax(logits, name=None):
    raise NotImplementedError('tf18shim: sparsemax (named in the reference\'s doc strings only)')
