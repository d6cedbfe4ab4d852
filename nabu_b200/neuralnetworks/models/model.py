"""Encoder-decoder model (reference: nabu/neuralnetworks/models/model.py:7-86)."""
import torch

from ... import engine
from .ed_encoders import ed_encoder_factory
from .ed_decoders import ed_decoder_factory


class Model(object):
    """Model(conf, trainlabels, constraint): builds the encoder and decoder named in model.cfg and
    owns the flat parameter store both write their variables into."""

    def __init__(self, conf, trainlabels, constraint=None, seed=0):
        self.conf = conf
        self.input_names = [n for n in conf.get('io', 'inputs').split(' ') if n]
        self.output_names = [n for n in conf.get('io', 'outputs').split(' ') if n]
        self.output_dims = {}
        for i, d in enumerate(conf.get('io', 'output_dims').split(' ')):
            self.output_dims[self.output_names[i]] = int(d) + trainlabels
        self.store = engine.ParamStore(seed)
        self.encoder = ed_encoder_factory.factory(conf.get('encoder', 'encoder'))(conf, constraint)
        self.decoder = ed_decoder_factory.factory(conf.get('decoder', 'decoder'))(
            conf, self.output_dims, constraint)
        self.encoder.store = self.store
        self.decoder.store = self.store

    def build(self, input_dims, device='cuda'):
        """Declare every variable (the TF graph-construction step) and allocate the flat buffers."""
        self.device = device
        if not self.store.materialised:
            encoded_dims = self.encoder.declare(dict(input_dims))
            self.decoder.declare(encoded_dims)
            self.store.materialise(device)
        return self

    def __call__(self, inputs, input_seq_length, targets, target_seq_length, is_training):
        if not self.store.materialised:
            first = next(iter(inputs.values()))
            self.build({k: v.shape[-1] for k, v in inputs.items()}, first.device)
        encoded, encoded_seq_length = self.encoder(inputs, input_seq_length, is_training)
        logits, logit_seq_length, _ = self.decoder(encoded, encoded_seq_length, targets, target_seq_length,
                                                   is_training)
        return logits, logit_seq_length

    # ---- model/model.pkl (trainers/trainer.py:790-792, scripts/decode.py:46-48) --------------------------------
    def description(self):
        """What the reference's pickle of the Model object amounts to: the parsed model.cfg and the number of
        train labels -- enough to rebuild the same model; the variables travel separately as model/network.ckpt."""
        conf = {sec: dict(self.conf.items(sec)) for sec in self.conf.sections()}
        trainlabels = self.output_dims[self.output_names[0]] - int(self.conf.get('io', 'output_dims').split(' ')[0])
        return {'format': 'nabu_b200.model/1', 'conf': conf, 'trainlabels': trainlabels}

    def save(self, path):
        import pickle
        with open(path, 'wb') as fid:
            pickle.dump(self.description(), fid, protocol=2)

    @staticmethod
    def load(path, seed=0):
        import configparser
        import pickle
        with open(path, 'rb') as fid:
            desc = pickle.load(fid)
        if not isinstance(desc, dict) or desc.get('format') != 'nabu_b200.model/1':
            raise Exception('%s was not written by nabu_b200 (a pickled TF-graph builder of the reference cannot be '
                            'loaded; rebuild from model.cfg and restore model/network.ckpt)' % path)
        conf = configparser.ConfigParser()
        conf.read_dict(desc['conf'])
        return Model(conf, desc['trainlabels'], None, seed=seed)

    @property
    def variables(self):
        return self.encoder.variables + self.decoder.variables
