"""Attention mechanisms (reference: components/attention.py:6-39 factory).

Only the description lives here; scoring, masking, softmax and the context are fused into
dec_attn_step (csrc/speller_kernels.cuh)."""


def factory(conf):
    """Returns (attention name, numfilt, filtersize) for the kernels."""
    if conf['probability_fn'] != 'softmax':
        raise Exception('probability_fn %s is outside the B200 hot path (SURVEY.md section 8 f4)'
                        % conf['probability_fn'])
    if conf['attention'] == 'location_aware':
        return 'location_aware', int(conf['numfilt']), int(conf['filtersize'])
    if conf['attention'] == 'vanilla':
        return 'vanilla', 0, 1
    if conf['attention'] == 'windowed':
        raise Exception('windowed attention is outside the B200 hot path (SURVEY.md section 8 f4)')
    raise Exception('unknown attention %s' % conf['attention'])
