"""Sequence helpers (reference: nabu/neuralnetworks/components/ops.py)."""
import torch

from ... import engine


def pyramid_stack(inputs, sequence_lengths, numsteps):
    """ops.py:6-60.  `inputs` must be the [B, T_pad, C] output of layer.blstm with T_pad a multiple
    of `numsteps` (layer.pblstm asks the kernel for that zero padding), so concatenating `numsteps`
    consecutive frames on the feature axis is a reshape: no data moves.  Lengths: ceil(len/n)."""
    B, T, C = inputs.shape
    if T % numsteps:
        raise ValueError('pyramid_stack needs T %% numsteps == 0 (got T=%d); use layer.pblstm' % T)
    outputs = inputs.reshape(B, T // numsteps, C * numsteps)
    return outputs, engine.pyramid_lengths(sequence_lengths, numsteps)


def dense_sequence_to_sparse(sequences, sequence_lengths):
    """ops.py:121-145 as (indices [N,2], values [N], dense_shape) host tensors."""
    sequences = sequences.cpu()
    lens = sequence_lengths.cpu()
    mask = torch.arange(sequences.shape[1])[None, :] < lens[:, None]
    idx = mask.nonzero()
    return idx, sequences[mask], tuple(sequences.shape)
