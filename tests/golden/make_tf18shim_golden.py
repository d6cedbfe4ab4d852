#!/usr/bin/env python
"""Golden vectors made by EXECUTING THE REFERENCE'S OWN PYTHON (from /root/reference) over tests/golden/tf18shim.

    python tests/golden/make_tf18shim_golden.py            # rewrites tests/golden/tf18shim_cases/*

Runs in the build container only (the reference tree is not on the GPU box); the directories it writes are committed
and are what tests/test_tf18_golden.py checks the oracle (CPU) and the CUDA path (GPU) against.  Same layout as
tools/tf18_dump.py (the script for a real TF-1.8 environment), so one harness serves both:

  model.cfg / trainer.cfg / recognizer.cfg, network.ckpt.* (float32, this repository's bundle writer),
  inputs.npz, outputs.npz (logits, logits_len, loss, grad/<variable>, decoded_*), PROVENANCE.txt

What executes here, unmodified apart from the in-memory Python-2 adaptations listed in tests/golden/py2ref.py:
  nabu/neuralnetworks/models/model.py                        Model.__init__, __call__, variables
  nabu/neuralnetworks/models/ed_encoders/{listener,dblstm,ed_encoder,ed_encoder_factory}.py
  nabu/neuralnetworks/models/ed_decoders/{speller,rnn_decoder,dnn_decoder,ed_decoder,ed_decoder_factory}.py
  nabu/neuralnetworks/components/{layer,ops,attention,rnn_cell,beam_search_decoder}.py
  nabu/neuralnetworks/trainers/loss_functions.py             CTC, average_cross_entropy
  nabu/neuralnetworks/decoders/{decoder_factory,ctc_decoder,beam_search_decoder,decoder}.py
  nabu/tools/default_conf.py and the defaults/*.cfg files next to the classes
TensorFlow's side of every call is the shim's restatement (see its module doc string for what that does and does not
pin).  Logits, loss and gradients are computed in fp64 and stored as float32; the beam search is run twice, in fp64 and
in fp32 (TF's arithmetic); a case whose ids differ between the two is rejected (a near-tie that rounding decides is
not a parity vector) and the seed is moved on.
"""
import configparser
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, 'tf18shim'))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import tensorflow as tf                                      # noqa: E402  (the shim)
import py2ref                                                # noqa: E402

OUT = os.environ.get('NABU_SHIM_OUT') or os.path.join(HERE, 'tf18shim_cases')

DBLSTM = ('[io]\ninputs = features\noutputs = text\noutput_dims = 7\n[encoder]\nencoder = dblstm\nnum_units = %(H)d\n'
          'num_layers = 2\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n')
LAS = ('[io]\ninputs = features\noutputs = text\noutput_dims = 6\n[encoder]\nencoder = listener\nnum_units = %(H)d\n'
       'num_layers = 2\npyramid_steps = %(steps)d\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = speller\n'
       'num_layers = 2\nnum_units = %(H)d\ndropout = 1\nsample_prob = 0\nattention = %(attention)s\n'
       'probability_fn = %(fn)s\n%(extra)s')
CTC_TRAINER = '[trainer]\ntrainer = standard\nloss = CTC\ntrainlabels = 1\ntargets = text\n'
CE_TRAINER = '[trainer]\ntrainer = standard\nloss = average_cross_entropy\ntrainlabels = 1\ntargets = text\n'
CTC_RECOGNIZER = '[recognizer]\nbatch_size = 4\n[decoder]\ndecoder = ctc_decoder\ntext_alphabet = a b c d e f g\n'
BEAM_RECOGNIZER = ('[recognizer]\nbatch_size = 4\n[decoder]\ndecoder = beam_search_decoder\nmax_steps = %(max_steps)d\n'
                   'beam_width = %(beam)d\nalphabet = a b c d e f <eos>\n')

CASES = {
    'dblstm_ctc': dict(kind='ctc', H=32, B=5, T=30, D=12),
    'las_vanilla': dict(kind='las', H=16, steps=2, attention='vanilla', fn='softmax', extra='', B=4, T=27, D=12,
                        beam=3, max_steps=8),
    'las_location_aware': dict(kind='las', H=32, steps=2, attention='location_aware', fn='softmax',
                               extra='numfilt = 3\nfiltersize = 5\n', B=5, T=37, D=12, beam=4, max_steps=10),
    'las_location_aware_normalized_sigmoid': dict(kind='las', H=16, steps=2, attention='location_aware',
                                                  fn='normalized_sigmoid', extra='numfilt = 4\nfiltersize = 7\n',
                                                  B=4, T=29, D=12, beam=3, max_steps=8),
    'las_vanilla_sigmoid': dict(kind='las', H=16, steps=2, attention='vanilla', fn='sigmoid', extra='', B=4, T=26,
                                D=12, beam=3, max_steps=8),
    'las_windowed_pyramid3': dict(kind='las', H=16, steps=3, attention='windowed', fn='softmax',
                                  extra='left_window_width = 2\nright_window_width = 3\n', B=4, T=50, D=12, beam=3,
                                  max_steps=8),
}


def _conf(text):
    conf = configparser.ConfigParser()
    conf.read_string(text)
    return conf


def _inputs(case, seed):
    rng = np.random.RandomState(seed)
    B, T, D = case['B'], case['T'], case['D']
    x = rng.randn(B, T, D).astype(np.float32)
    xl = rng.randint(int(0.6 * T), T + 1, size=B).astype(np.int32)
    xl[0] = T
    if case['kind'] == 'ctc':
        yl = np.maximum(xl // 10, 1).astype(np.int32)
        y = rng.randint(0, 7, size=(B, int(yl.max()))).astype(np.int32)
    else:
        yl = rng.randint(3, 8, size=B).astype(np.int32)
        y = rng.randint(0, 6, size=(B, int(yl.max()))).astype(np.int32)
        for b in range(B):
            y[b, yl[b] - 1] = 6                              # EOS = output_dims (string_reader_eos.py:88-111)
    for b in range(B):
        x[b, xl[b]:] = 0
        y[b, yl[b]:] = 0
    return dict(features=x, features_len=xl, targets=y, targets_len=yl)


def _livelier_values(variables, seed):
    """The initialisers give near-uniform output distributions (no beam ever ends, no gate saturates).  The values are
    this script's to choose - they travel in the checkpoint - so: every matrix times 2.5, biases redrawn from N(0, 0.3),
    and the speller's output bias pushed towards EOS so that hypotheses finish at different steps."""
    import torch
    rng = np.random.RandomState(seed + 7)
    with torch.no_grad():
        for v in variables:
            if v.t.dim() >= 2:
                v.t.mul_(2.5)
            else:
                v.t.copy_(torch.from_numpy(rng.normal(0, 0.3, size=tuple(v.t.shape))))
            if v.op.name == 'Speller/decoder/dense/bias':
                v.t[-1] += 0.4
            v.t.copy_(v.t.float().double())                   # the values ARE float32 numbers


def _decode(recognizer_cfg, model, inp, bits, i_name='features', o_name='text'):
    """a new graph (names start over, variables stay = restored) in the given precision, through the reference's
    decoder_factory and the decoder's own __call__"""
    from nabu.neuralnetworks.decoders import decoder_factory
    tf.reset_default_graph()
    tf.cast_variables_(bits)
    decoder = decoder_factory.factory(recognizer_cfg.get('decoder', 'decoder'))(recognizer_cfg, model)
    out = decoder({i_name: tf.constant(inp['features'])}, {i_name: tf.constant(inp['features_len'])})[o_name]
    tf.cast_variables_(64)
    return out


def make_case(name, case, seed):
    ctc = case['kind'] == 'ctc'
    model_cfg = _conf((DBLSTM if ctc else LAS) % case)
    trainer_cfg = _conf(CTC_TRAINER if ctc else CE_TRAINER)
    recognizer_cfg = _conf(CTC_RECOGNIZER if ctc else BEAM_RECOGNIZER % case)
    return run_case(os.path.join(OUT, name), model_cfg, trainer_cfg, recognizer_cfg, _inputs(case, seed + 1), seed)


def make_recipe_case(recipe_dir, path, seed=100, B=3, T=60, max_steps=12, beam_width=None):
    """the reference's own SHIPPED recipe (config/recipes/<...>/{model,trainer,recognizer}.cfg, unchanged except that the
    stochastic parts are off - input_noise 0, dropout 1, sample_prob 0 - and the beam search is cut to `max_steps`)"""
    def read(fname):
        conf = configparser.ConfigParser()
        conf.read(os.path.join(recipe_dir, fname))
        return conf
    model_cfg, trainer_cfg, recognizer_cfg = read('model.cfg'), read('trainer.cfg'), read('recognizer.cfg')
    model_cfg.set('encoder', 'input_noise', '0')
    model_cfg.set('encoder', 'dropout', '1')
    las = model_cfg.get('decoder', 'decoder') == 'speller'
    if las:
        model_cfg.set('decoder', 'dropout', '1')
        model_cfg.set('decoder', 'sample_prob', '0')
        recognizer_cfg.set('decoder', 'max_steps', str(max_steps))
        if beam_width:
            recognizer_cfg.set('decoder', 'beam_width', str(beam_width))
    V = int(model_cfg.get('io', 'output_dims').split(' ')[0])
    rng = np.random.RandomState(seed + 1)
    x = rng.randn(B, T, 40).astype(np.float32)
    xl = rng.randint(int(0.6 * T), T + 1, size=B).astype(np.int32)
    xl[0] = T
    if las:
        yl = rng.randint(3, 8, size=B).astype(np.int32)
        y = rng.randint(0, V, size=(B, int(yl.max()))).astype(np.int32)
        for b in range(B):
            y[b, yl[b] - 1] = V
    else:
        yl = np.maximum(xl // 10, 1).astype(np.int32)
        y = rng.randint(0, V, size=(B, int(yl.max()))).astype(np.int32)
    for b in range(B):
        x[b, xl[b]:] = 0
        y[b, yl[b]:] = 0
    return run_case(path, model_cfg, trainer_cfg, recognizer_cfg,
                    dict(features=x, features_len=xl, targets=y, targets_len=yl), seed)


def run_case(path, model_cfg, trainer_cfg, recognizer_cfg, inp, seed):
    from nabu.neuralnetworks.models.model import Model
    from nabu.neuralnetworks.trainers import loss_functions

    tf.reset_all(seed)
    tf.set_float_bits(64)
    i_name = model_cfg.get('io', 'inputs').split(' ')[0]
    o_name = model_cfg.get('io', 'outputs').split(' ')[0]
    ctc = trainer_cfg.get('trainer', 'loss') == 'CTC'

    # ---- the training graph: Model, loss function, gradients (reference trainer.py:297-371 builds exactly these)
    model = Model(conf=model_cfg, trainlabels=int(trainer_cfg.get('trainer', 'trainlabels')), constraint=None)
    feed = lambda key: {(i_name if key.startswith('features') else o_name): tf.constant(inp[key])}   # noqa: E731
    model(feed('features'), feed('features_len'), feed('targets'), feed('targets_len'), True)   # creates the variables
    _livelier_values(model.variables, seed)
    tf.reset_default_graph()
    logits, logit_len = model(feed('features'), feed('features_len'), feed('targets'), feed('targets_len'), True)
    loss = loss_functions.factory(trainer_cfg.get('trainer', 'loss'))(
        feed('targets'), logits, logit_len, feed('targets_len'))
    variables = model.variables
    assert len(set(v.op.name for v in variables)) == len(variables) == len(tf.global_variables())
    loss.t.backward()
    out = {'logits': logits[o_name].numpy().astype(np.float32), 'logits_len': logit_len[o_name].numpy().astype(np.int32),
           'loss': np.float32(loss.numpy())}
    for v in variables:
        assert v.t.grad is not None, v.op.name
        out['grad/' + v.op.name] = v.t.grad.numpy().astype(np.float32)
    params = {v.op.name: v.numpy().astype(np.float32) for v in variables}

    # ---- the decoding graph (reference recognizer.py:60-83)
    dec64 = _decode(recognizer_cfg, model, inp, 64, i_name, o_name)
    assert len(tf.global_variables()) == len(variables), 'the decoding graph created variables of its own'
    if ctc:
        out['decoded_indices'] = dec64.indices.numpy().astype(np.int64)
        out['decoded_values'] = dec64.values.numpy().astype(np.int32)
        out['decoded_shape'] = dec64.dense_shape.numpy().astype(np.int64)
    else:
        dec32 = _decode(recognizer_cfg, model, inp, 32, i_name, o_name)
        seq64, len64 = dec64[0].numpy(), dec64[1].numpy()
        seq32, len32 = dec32[0].numpy(), dec32[1].numpy()
        if not (np.array_equal(len64, len32) and np.array_equal(seq64, seq32)):
            return None                                       # rounding decides a near-tie: not a parity vector
        out['decoded_sequences'], out['decoded_lengths'] = seq64.astype(np.int32), len64.astype(np.int32)
        out['decoded_scores'] = dec64[2].numpy().astype(np.float32)
        out['decoded_alignments'] = dec64[3].numpy().astype(np.float32)

    # ---- write
    from nabu_b200.processing.tfcheckpoint import write_checkpoint
    if os.path.isdir(path):
        shutil.rmtree(path)
    os.makedirs(path)
    for fname, conf in (('model.cfg', model_cfg), ('trainer.cfg', trainer_cfg), ('recognizer.cfg', recognizer_cfg)):
        with open(os.path.join(path, fname), 'w') as fid:
            conf.write(fid)
    np.savez(os.path.join(path, 'inputs.npz'), **inp)
    np.savez(os.path.join(path, 'outputs.npz'), **out)
    write_checkpoint(os.path.join(path, 'network.ckpt'), params, state_file=False)
    with open(os.path.join(path, 'PROVENANCE.txt'), 'w') as fid:
        fid.write('made by tests/golden/make_tf18shim_golden.py (seed %d): the reference\'s own Python, vrenkens/nabu @ '
                  '39deb62, executed from /root/reference over tests/golden/tf18shim (an eager restatement of the '
                  'TensorFlow-1.8 calls on torch fp64).\nNOT an output of TensorFlow itself.\nvariables: %d, loss %.9g\n'
                  % (seed, len(variables), float(out['loss'])))
    return out


def main():
    py2ref.install(os.environ.get('NABU_REFERENCE', '/root/reference'))
    if len(sys.argv) > 3 and sys.argv[1] == '--recipe':          # --recipe <recipe dir> <output dir> [seed]
        for seed in range(int(sys.argv[4]) if len(sys.argv) > 4 else 100, 140):
            if make_recipe_case(sys.argv[2], sys.argv[3], seed) is not None:
                print('%s: seed %d' % (sys.argv[2], seed))
                return
        raise RuntimeError('no robust seed')
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        for seed in range(100, 120):
            out = make_case(name, case, seed)
            if out is not None:
                break
            print('%s: seed %d rejected (fp32 and fp64 beam searches differ)' % (name, seed))
        else:
            raise RuntimeError('%s: no robust seed' % name)
        extra = ''
        if 'decoded_lengths' in out:
            extra = ' beam lengths %s' % out['decoded_lengths'].tolist()
        print('%s: seed %d, loss %.6f, %d gradients%s' % (name, seed, out['loss'], sum(k.startswith('grad/') for k in out),
                                                          extra))


if __name__ == '__main__':
    main()
