"""Training losses (reference: nabu/neuralnetworks/trainers/loss_functions.py:7-214).

All take (targets, logits, logit_seq_length, target_seq_length) dicts and return a scalar tensor;
CTC and the softmax cross-entropies run fused loss+gradient kernels."""
from ... import engine


def factory(loss_function):
    if loss_function == 'average_cross_entropy':
        return average_cross_entropy
    if loss_function == 'CTC':
        return CTC
    if loss_function in ('sum_cross_entropy', 'average_sigmoid_cross_entropy', 'marigin'):
        raise Exception('loss function %s is outside the B200 hot path (SURVEY.md section 8)' % loss_function)
    raise Exception('unknown loss function %s' % loss_function)


def average_cross_entropy(targets, logits, logit_seq_length, target_seq_length):
    """loss_functions.py:155-165: masked CE summed over time / target length, batch mean, summed
    over outputs."""
    loss = None
    for t in targets:
        l = engine.masked_ce_mean(logits[t], targets[t], logit_seq_length[t], target_seq_length[t])
        loss = l if loss is None else loss + l
    return loss


def CTC(targets, logits, logit_seq_length, target_seq_length):
    """loss_functions.py:180-214: batch mean of tf.nn.ctc_loss (blank = last class), summed over
    outputs.  The dense->sparse label conversion of the reference is folded into the kernel
    (labels [B,L] + lengths)."""
    loss = None
    for t in targets:
        l = engine.ctc_mean(logits[t], logit_seq_length[t], targets[t], target_seq_length[t])
        loss = l if loss is None else loss + l
    return loss
