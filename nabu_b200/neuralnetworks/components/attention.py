"""Attention mechanisms (reference: components/attention.py:6-39 factory).

Only the description lives here; scoring, masking, softmax and the context are fused into
dec_attn_step (csrc/speller_kernels.cuh)."""


def factory(conf):
    """Returns (attention name, numfilt, filtersize) for the kernels.  A probability function other than softmax
    (attention.py:9-13: `normalized_sigmoid`, `sigmoid`) travels as a suffix of the name: 'location_aware+sigmoid'."""
    pf = conf['probability_fn']
    if pf not in ('softmax', 'normalized_sigmoid', 'sigmoid'):
        raise Exception('unknown probability_fn %s' % pf)
    suffix = '' if pf == 'softmax' else '+' + pf
    if conf['attention'] == 'location_aware':
        return 'location_aware' + suffix, int(conf['numfilt']), int(conf['filtersize'])
    if conf['attention'] == 'vanilla':
        return 'vanilla' + suffix, 0, 1
    if conf['attention'] == 'windowed':          # the two widths travel in the numfilt / filtersize slots
        return 'windowed' + suffix, int(conf['left_window_width']), int(conf['right_window_width'])
    raise Exception('unknown attention %s' % conf['attention'])
