"""Pyramidal BLSTM listener (reference: .../ed_encoders/listener.py:14-74)."""
import torch

from . import ed_encoder
from ...components import layer


class Listener(ed_encoder.EDEncoder):

    def declare(self, input_dims):
        out = {}
        n, H, steps = (int(self.conf['num_layers']), int(self.conf['num_units']),
                       int(self.conf['pyramid_steps']))
        for inp, dim in input_dims.items():
            for l in range(n):
                dim = layer.declare_pblstm(self.store, '%s/%s/layer%d' % (self.scope, inp, l), dim, H, steps)
            out[inp] = layer.declare_blstm(self.store, '%s/%s/layer%d' % (self.scope, inp, n), dim, H)
        return out

    def encode(self, inputs, input_seq_length, is_training):
        encoded, encoded_seq_length = {}, {}
        noise = float(self.conf['input_noise'])
        keep = float(self.conf['dropout'])
        n, H, steps = (int(self.conf['num_layers']), int(self.conf['num_units']),
                       int(self.conf['pyramid_steps']))
        for inp in inputs:
            h = inputs[inp]
            lens = input_seq_length[inp]
            if is_training and noise > 0:
                h = h + torch.randn_like(h) * noise
            planes = None       # tensor-core operand planes of h, valid while h is a (stacked) BLSTM output nobody touched
            for l in range(n):
                h, lens, planes = layer.pblstm(self.store, h, lens, H, steps, '%s/%s/layer%d' % (self.scope, inp, l),
                                               planes=planes, want_planes=True)
                if is_training and keep < 1:
                    h = torch.nn.functional.dropout(h, 1 - keep, True)
                    planes = None
            h = layer.blstm(self.store, h, lens, H, '%s/%s/layer%d' % (self.scope, inp, n), planes=planes)
            if is_training and keep < 1:
                h = torch.nn.functional.dropout(h, 1 - keep, True)
            encoded[inp] = h
            encoded_seq_length[inp] = lens
        return encoded, encoded_seq_length
