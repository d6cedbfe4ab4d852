"""Deep bidirectional LSTM encoder (reference: .../ed_encoders/dblstm.py:11-59)."""
import torch

from . import ed_encoder
from ...components import layer


class DBLSTM(ed_encoder.EDEncoder):

    def declare(self, input_dims):
        out = {}
        for inp, dim in input_dims.items():
            for l in range(int(self.conf['num_layers'])):
                dim = layer.declare_blstm(self.store, '%s/%s/layer%d' % (self.scope, inp, l), dim,
                                          int(self.conf['num_units']))
            out[inp] = dim
        return out

    def encode(self, inputs, input_seq_length, is_training):
        encoded, encoded_seq_length = {}, {}
        noise = float(self.conf['input_noise'])
        keep = float(self.conf['dropout'])
        for inp in inputs:
            h = inputs[inp]
            if is_training and noise > 0:
                h = h + torch.randn_like(h) * noise
            planes = None       # tensor-core operand planes of h, valid while h is a BLSTM output nobody touched
            for l in range(int(self.conf['num_layers'])):
                h, planes = layer.blstm(self.store, h, input_seq_length[inp], int(self.conf['num_units']),
                                        '%s/%s/layer%d' % (self.scope, inp, l), planes=planes, want_planes=True)
                if is_training and keep < 1:
                    h = torch.nn.functional.dropout(h, 1 - keep, True)
                    planes = None
            encoded[inp] = h
            encoded_seq_length[inp] = input_seq_length[inp]
        return encoded, encoded_seq_length
