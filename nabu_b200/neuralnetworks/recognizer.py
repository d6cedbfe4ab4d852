"""Recognizer (reference: nabu/neuralnetworks/recognizer.py:18-143) -- SURVEY.md section 8 row f2.

Decodes a data set batch by batch with the decoder named in the recognizer cfg and writes the results
through `decoder.write` into `<expdir>/decoded`.  The reference builds a TF input pipeline over the
database sections named in the cfg and restores `model/network.ckpt`; here the batches come from a batch
source (an iterable of `(inputs, input_seq_length)` dict pairs plus the utterance names -- the TFRecord
pipeline is row f1) and the parameters from the TF checkpoint `<expdir>/model/network.ckpt` (row f3; what
`Trainer.train` here and the reference's SaveAtEnd hook both write), else `<expdir>/model/network.pt`.  The
reference's default file `defaults/recognizer.cfg` does not exist and `apply_defaults` tolerates that (tools/default_conf.py:19); `batch_size` must therefore be in the cfg.
"""
import os
import shutil

import torch

from ..tools.default_conf import apply_defaults
from .decoders import decoder_factory


class Recognizer(object):
    """Recognizer(model, conf, dataconf, expdir).recognize()"""

    def __init__(self, model, conf, dataconf, expdir, batch_source=None, names=None):
        self.conf = dict(conf.items('recognizer'))
        apply_defaults(self.conf, os.path.join(os.path.dirname(os.path.realpath(__file__)), 'defaults',
                                               type(self).__name__.lower() + '.cfg'))
        self.expdir = expdir
        self.model = model
        self.dataconf = dataconf
        self.decoder = decoder_factory.factory(conf.get('decoder', 'decoder'))(conf, self.model)
        self.batch_size = int(self.conf['batch_size'])
        self.batch_source = batch_source
        # the reference's names carry the index the pipeline appended ("<utt>-<i>"); it is cut off before writing
        self.names = list(names) if names is not None else None

    def _input_dims(self):
        src = getattr(self, '_src', None)
        if src is None:
            raise Exception('the model has no variables yet and the batch source does not tell the input dimensions')
        return src.input_dims

    def recognize(self):
        if self.batch_source is None:
            if self.dataconf is None:
                raise Exception('Recognizer.recognize needs a batch_source or a database configuration')
            from ..processing import input_pipeline             # recognizer.py:42-81: inputs only, smaller final batch
            src = input_pipeline.source_from_conf(self.conf, self.dataconf, self.model.input_names, [],
                                                  device=getattr(self.model, 'device', 'cuda'),
                                                  allow_smaller_final_batch=True)
            self.names = src.names
            self._src = src
            self.batch_source = ((b[0], b[1]) for b in src)
        # LoadAtBegin (recognizer.py:105-108): model/network.ckpt, a TF checkpoint -- one written by Trainer.train here
        # or by a nabu / TF-1.8 training run; network.pt (the store with its Adam moments) is the fallback
        tfckpt = os.path.join(self.expdir, 'model', 'network.ckpt') if self.expdir else None
        ckpt = os.path.join(self.expdir, 'model', 'network.pt') if self.expdir else None
        if tfckpt and os.path.isfile(tfckpt + '.index'):
            if not self.model.store.materialised:
                self.model.build(self._input_dims(), getattr(self.model, 'device', 'cuda'))
            self.model.store.load_tf_checkpoint(tfckpt)
        elif ckpt and os.path.isfile(ckpt) and self.model.store.materialised:
            self.model.store.load_state_dict(torch.load(ckpt))
        directory = os.path.join(self.expdir, 'decoded')
        if os.path.isdir(directory):
            shutil.rmtree(directory)
        os.makedirs(directory)
        nameid = 0
        for inputs, input_seq_length in self.batch_source:
            outputs = self.decoder(inputs, input_seq_length)
            n = int(list(input_seq_length.values())[0].shape[0])
            if self.names is not None:
                names = self.names[nameid:nameid + n]
                names = ['-'.join(name.split('-')[:-1]) if '-' in name else name for name in names]
            else:
                names = ['utt%d' % (nameid + i) for i in range(n)]
            self.decoder.write(outputs, directory, names)
            nameid += n
        return directory
