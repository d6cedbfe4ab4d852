"""Shared helpers for the parity tests: seeded synthetic inputs and the
configs the reference recipes would provide."""
import configparser

import numpy as np


def make_conf(text):
    c = configparser.ConfigParser()
    c.read_string(text)
    return c


def synthetic_ctc_batch(B, T, D, V, ragged, seed=1234):
    """features N(0,1) [B,T,D]; lengths = T or U{ceil(.6T)..T}; labels U{0..V-2}, L = len//10."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    if ragged:
        lens = np.random.default_rng(4321).integers(int(np.ceil(0.6 * T)), T + 1, size=B).astype(np.int32)
        lens[0] = T
    else:
        lens = np.full(B, T, np.int32)
    lab_len = np.maximum(lens // 10, 1).astype(np.int32)
    Lmax = int(lab_len.max())
    labels = np.random.default_rng(99).integers(0, V - 1, size=(B, Lmax)).astype(np.int32)
    for b in range(B):
        x[b, lens[b]:] = 0
    return x, lens, labels, lab_len


def synthetic_las_batch(B, T, D, V, U, ragged, seed=1234):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    if ragged:
        lens = np.random.default_rng(4321).integers(int(np.ceil(0.6 * T)), T + 1, size=B).astype(np.int32)
        lens[0] = T
        tl = np.random.default_rng(77).integers(max(1, U // 2), U + 1, size=B).astype(np.int32)
        tl[0] = U
    else:
        lens = np.full(B, T, np.int32)
        tl = np.full(B, U, np.int32)
    targets = np.random.default_rng(99).integers(0, V - 1, size=(B, U)).astype(np.int32)
    for b in range(B):
        targets[b, tl[b] - 1] = V - 1          # EOS terminated (string_reader_eos.py:57-60)
        targets[b, tl[b]:] = 0
        x[b, lens[b]:] = 0
    return x, lens, targets, tl


def rel_err(a, b):
    """max |a - b| / max |b| over the whole tensor (the error relative to the tensor's own scale)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def row_rel_err(a, b, floor=1e-3):
    """The worst ROW-normalised error: max over the rows r (all leading axes) of max |a_r - b_r| / max |b_r|.  Rows whose
    own scale is below `floor` x the tensor's scale are normalised by that floor instead (a row of zeros has no relative
    error; fp32 sums over 1e5 terms do not resolve 1e-3 of the tensor scale to 1e-4 either)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.ndim < 2:
        return rel_err(a, b)
    a2, b2 = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    scale = np.maximum(np.abs(b2).max(axis=1), floor * max(np.abs(b2).max(), 1e-30))
    return float((np.abs(a2 - b2).max(axis=1) / scale).max())


class ParityLog(object):
    """Collects (tensor, global relative error, worst row-normalised error) of a parity test, asserts the bound on both
    and leaves the table under gpurun_out/parity/ when that directory's parent exists (the GPU box), so that the errors
    actually measured -- not just pass/fail -- end up in profiles/."""

    def __init__(self, name):
        self.name, self.rows = name, []

    def check(self, tensor, got, ref, tol, row_tol=None):
        g, r = rel_err(got, ref), row_rel_err(got, ref)
        self.rows.append((tensor, g, r, tol, row_tol))
        assert g < tol, '%s: %s relative error %.3e >= %.1e' % (self.name, tensor, g, tol)
        if row_tol is not None:
            assert r < row_tol, '%s: %s row-normalised error %.3e >= %.1e' % (self.name, tensor, r, row_tol)

    def dump(self):
        import json
        import os
        root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
        if not os.path.isdir(root):
            return
        os.makedirs(os.path.join(root, 'parity'), exist_ok=True)
        with open(os.path.join(root, 'parity', self.name + '.json'), 'w') as fid:
            json.dump([{'tensor': t, 'rel_err': g, 'row_rel_err': r, 'tol': tol, 'row_tol': rt}
                       for t, g, r, tol, rt in self.rows], fid, indent=1)


def write_experiment(root, num_epochs=2, variable_batch_size=True, n_train=16, model='dblstm'):
    """An experiment directory as `run train` prepares it (database.conf, model.cfg, trainer.cfg,
    validation_evaluator.cfg, test_evaluator.cfg, recognizer.cfg) over small data directories in nabu's on-disk
    format under `root`: DBLSTM 2x64 + CTC (model='dblstm', the DBLSTM/TIMIT recipe's shape) or Listener 2x64 +
    Speller + beam search (model='las', the LAS/TIMIT recipe's shape) on 40-dim features, 4 symbols.  Returns the
    experiment directory."""
    import os
    from nabu_b200.processing import tfwriters
    rng = np.random.default_rng(5)
    alphabet = ['a', 'b', 'c', 'd']
    D = 40
    sections = []
    for tag, n in (('train', n_train), ('dev', 4), ('test', 6)):
        lens = rng.integers(20, 50, size=n)
        fdir, tdir = os.path.join(root, tag + 'fbank'), os.path.join(root, tag + 'text')
        fw, tw = tfwriters.factory('audio_feature')(fdir), tfwriters.factory('string_eos')(tdir)
        for i, L in enumerate(lens):
            fw.write(rng.standard_normal((L, D)).astype(np.float32), '%s%d' % (tag, i))
            tw.write(' '.join(rng.choice(alphabet, size=max(1, L // 12))), '%s%d' % (tag, i))
        fw.write_metadata(D)
        tw.write_metadata(alphabet)
        sections.append('[%sfbank]\ndir = %s\ntype = audio_feature\n[%stext]\ndir = %s\ntype = string_eos\n'
                        % (tag, fdir, tag, tdir))
    expdir = os.path.join(root, 'exp')
    os.makedirs(expdir)
    V = len(alphabet) + 1                    # + EOS (string_eos appends it); CTC adds its blank through trainlabels
    files = {
        'database.conf': ''.join(sections),
        'model.cfg': '[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = dblstm\n'
                     'num_units = 64\nnum_layers = 2\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = dnn_decoder\n'
                     'num_layers = 0\n' % V,
        'trainer.cfg': '[trainer]\ntrainer = standard\nloss = CTC\ntrainlabels = 1\ntargets = text\nnum_epochs = %d\n'
                       'batch_size = 4\nnumbuckets = 2\nvariable_batch_size = %s\nvalid_frequency = 3\n'
                       'num_tries = None\nfeatures = trainfbank\ntext = traintext\n'
                       % (num_epochs, variable_batch_size),
        'validation_evaluator.cfg': '[evaluator]\nevaluator = loss_evaluator\nloss = CTC\ntargets = text\n'
                                    'batch_size = 2\nfeatures = devfbank\ntext = devtext\n',
        'test_evaluator.cfg': '[evaluator]\nevaluator = decoder_evaluator\ntargets = text\nbatch_size = 3\n'
                              'features = testfbank\ntext = testtext\n[decoder]\ndecoder = ctc_decoder\n'
                              'text_alphabet = %s\n' % ' '.join(alphabet + ['<eos>']),
        'recognizer.cfg': '[recognizer]\nbatch_size = 4\nfeatures = testfbank\n[decoder]\ndecoder = ctc_decoder\n'
                          'text_alphabet = %s\n' % ' '.join(alphabet + ['<eos>']),
    }
    if model == 'las':
        beam = '[decoder]\ndecoder = beam_search_decoder\nmax_steps = 12\nbeam_width = 4\nalphabet = %s\n' \
            % ' '.join(alphabet + ['<eos>'])
        files.update({
            'model.cfg': '[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = listener\n'
                         'input_noise = 0.6\nnum_layers = 2\nnum_units = 64\npyramid_steps = 2\ndropout = 0.5\n'
                         '[decoder]\ndecoder = speller\nnum_layers = 2\nnum_units = 64\ndropout = 0.5\n' % (V - 1),
            'trainer.cfg': files['trainer.cfg'].replace('loss = CTC', 'loss = average_cross_entropy'),
            'validation_evaluator.cfg': '[evaluator]\nevaluator = decoder_evaluator\ntargets = text\nbatch_size = 2\n'
                                        'features = devfbank\ntext = devtext\n' + beam,
            'test_evaluator.cfg': '[evaluator]\nevaluator = loss_evaluator\nloss = average_cross_entropy\n'
                                  'targets = text\nbatch_size = 3\nfeatures = testfbank\ntext = testtext\n',
            'recognizer.cfg': '[recognizer]\nbatch_size = 4\nfeatures = testfbank\n' + beam,
        })
    for name, text in files.items():
        with open(os.path.join(expdir, name), 'w') as fid:
            fid.write(text)
    return expdir
