"""CTC decoder (reference: nabu/neuralnetworks/decoders/ctc_decoder.py:10-135)."""
import collections
import os

import numpy as np
import torch

from . import decoder
from ... import engine

SparseTensorValue = collections.namedtuple('SparseTensorValue', ['indices', 'values', 'dense_shape'])


class CTCDecoder(decoder.Decoder):
    """Model forward (is_training=False) + tf.nn.ctc_beam_search_decoder defaults: beam 100, top path,
    merge_repeated=True, blank = last class."""

    def __init__(self, conf, model):
        super(CTCDecoder, self).__init__(conf, model)
        self.alphabets = {o: self.conf['%s_alphabet' % o].split(' ') for o in model.output_names}

    def __call__(self, inputs, input_seq_length):
        with torch.no_grad():
            logits, logits_seq_length = self.model(inputs, input_seq_length, targets=[], target_seq_length=[],
                                                   is_training=False)
            outputs = {}
            for o in logits:
                ids, lens, _ = engine.ctc_beam_search(logits[o], logits_seq_length[o], beam_width=100,
                                                      merge_repeated=True)
                ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
                # the int32 SparseTensor the reference returns (ctc_decoder.py:57-59)
                mask = np.arange(ids.shape[1])[None, :] < lens[:, None]
                indices = np.argwhere(mask).astype(np.int64)
                width = int(lens.max()) if lens.size else 0
                outputs[o] = SparseTensorValue(indices, ids[mask].astype(np.int32),
                                               np.array([ids.shape[0], width], np.int64))
        return outputs

    def write(self, outputs, directory, names):
        for o in outputs:
            batch_size = int(outputs[o].dense_shape[0])
            with open(os.path.join(directory, o), 'a') as fid:
                for i in range(batch_size):
                    sel = np.where(outputs[o].indices[:, 0] == i)[0]
                    text = ' '.join(self.alphabets[o][j] for j in outputs[o].values[sel])
                    fid.write('%s %s\n' % (names[i], text))

    def update_evaluation_loss(self, loss, outputs, references, reference_seq_length):
        """loss: decoder.RunningLoss.  Sum of edit distances / number of reference labels."""
        errors, batch_targets = 0, 0
        for o in outputs:
            refs = references[o].cpu().numpy() if torch.is_tensor(references[o]) else np.asarray(references[o])
            rl = reference_seq_length[o].cpu().numpy() if torch.is_tensor(reference_seq_length[o]) \
                else np.asarray(reference_seq_length[o])
            for i in range(int(outputs[o].dense_shape[0])):
                sel = np.where(outputs[o].indices[:, 0] == i)[0]
                errors += decoder.edit_distance(outputs[o].values[sel], refs[i, :rl[i]])
            batch_targets += int(rl.sum())
        return loss.update(errors, batch_targets)
