"""LAS beam-search decoder (reference: nabu/neuralnetworks/decoders/beam_search_decoder.py:11-196)."""
import os

import numpy as np
import torch

from . import decoder


class BeamSearchDecoder(decoder.Decoder):
    """Encoder forward, then the fused beam search over the Speller cell.  Outputs per output name the
    reference's tuple (sequences [B,W,L], lengths [B,W], scores [B,W], alignments [B,W,L,T'])."""

    def __init__(self, conf, model):
        super(BeamSearchDecoder, self).__init__(conf, model)
        self.alphabet = self.conf['alphabet'].split(' ')

    def __call__(self, inputs, input_seq_length):
        output_name = list(self.model.output_dims.keys())[0]
        with torch.no_grad():
            encoded, encoded_seq_length = self.model.encoder(inputs, input_seq_length, False)
            # the reference tiles memory and lengths beam_width times (tile_batch); the kernels index
            # the un-tiled memory instead
            cell = self.model.decoder.create_cell(encoded, encoded_seq_length, False)
            sequences, lengths, scores, alignments = cell.beam_search(
                int(self.conf['beam_width']), int(self.conf['max_steps']), float(self.conf['length_penalty']),
                float(self.conf['temperature']))
        return {output_name: (sequences, lengths, scores, alignments)}

    def write(self, outputs, directory, names):
        sequences, lengths, scores, alignments = [t.cpu().numpy() for t in list(outputs.values())[0]]
        for i, name in enumerate(names):
            with open(os.path.join(directory, name), 'w') as fid:
                for b in range(sequences.shape[1]):
                    text = ' '.join(self.alphabet[s] for s in sequences[i, b, :lengths[i, b]])
                    fid.write('%f %s\n' % (scores[i, b], text))
            if self.conf.get('visualize_alignments') == 'True':
                np.save(os.path.join(directory, name + '_alignments.npy'), alignments[i])

    def update_evaluation_loss(self, loss, outputs, references, reference_seq_length):
        """Edit distance of the best hypothesis against the reference without its EOS
        (beam_search_decoder.py:166-196: reference_seq_length - 1), per reference label incl. EOS."""
        sequences, lengths = [t.cpu().numpy() for t in list(outputs.values())[0][:2]]
        refs = list(references.values())[0]
        rl = list(reference_seq_length.values())[0]
        refs = refs.cpu().numpy() if torch.is_tensor(refs) else np.asarray(refs)
        rl = rl.cpu().numpy() if torch.is_tensor(rl) else np.asarray(rl)
        errors = 0
        for i in range(sequences.shape[0]):
            errors += decoder.edit_distance(sequences[i, 0, :lengths[i, 0]], refs[i, :rl[i] - 1])
        return loss.update(errors, int(rl.sum()))
