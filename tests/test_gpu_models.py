"""GPU parity tests, model level: the Python mirror of nabu's plugin API (Model / Trainer / losses)
driving the CUDA kernels, against the NumPy oracle on shared weights and seeded inputs."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from tests.util import make_conf, rel_err, synthetic_ctc_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4      # BASELINE.json north_star: 1e-4 relative fp32


def _dblstm_layers(params, NL, inp='features'):
    layers = []
    for l in range(NL):
        base = 'DBLSTM/%s/layer%d/bidirectional_rnn/%%s/layer_norm_basic_lstm_cell/%%s' % (inp, l)
        layers.append({'fw_kernel': params[base % ('fw', 'kernel')], 'fw_bias': params[base % ('fw', 'bias')],
                       'bw_kernel': params[base % ('bw', 'kernel')], 'bw_bias': params[base % ('bw', 'bias')]})
    return layers


def _ctc_trainer(H, NL, V, dev, seed=3):
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = dblstm\n'
                      'num_units = %d\nnum_layers = %d\ninput_noise = 0\ndropout = 1\n[decoder]\n'
                      'decoder = dnn_decoder\nnum_layers = 0\n' % (V - 1, H, NL))
    tconf = make_conf('[trainer]\ntrainer = standard\nloss = CTC\ntrainlabels = 1\ntargets = text\n'
                      'initial_learning_rate = 1e-3\nlearning_rate_decay = 0.1\n')
    tr = trainer_factory.factory('standard')(tconf, None, mconf, None, None, None, 0, device=dev, seed=seed)
    tr.num_steps = 100
    return tr


@pytest.mark.parametrize('B,T,H,NL,ragged', [(6, 40, 64, 2, True), (32, 200, 256, 2, True),
                                              (6, 40, 100, 2, True)])     # 100: not a multiple of the 64-unit tile
def test_dblstm_ctc_train_step_matches_oracle(B, T, H, NL, ragged):
    """cfg-1 (DBLSTM 2x256 + CTC, 29 labels, 32x200x40): loss, every gradient and the Adam update."""
    dev = torch.device('cuda', 0)
    D, V = 40, 29
    tr = _ctc_trainer(H, NL, V, dev)
    tr.model.build({'features': D}, dev)
    store = tr.model.store
    x, lens, labels, ll = synthetic_ctc_batch(B, T, D, V, ragged)
    params = store.to_numpy()
    batch = ({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)},
             {'text': torch.from_numpy(labels).to(dev)}, {'text': torch.from_numpy(ll).to(dev)})
    loss, lr = tr.update(*batch)
    # oracle (float64) on the same weights
    layers = _dblstm_layers(params, NL)
    lin = {'weights': params['DNNDecoder/text/outlayer/weights'], 'biases': params['DNNDecoder/text/outlayer/biases']}
    enc, _, caches = O.dblstm_fwd(x, lens, layers)
    logits = O.linear_fwd(enc, lin)
    ref_loss, dlogits = O.ctc_loss_mean(logits, lens, labels, ll)
    denc, glin = O.linear_bwd(enc, lin, dlogits)
    _, glayers = O.dblstm_bwd(caches, denc)
    assert abs(float(loss) - ref_loss) / abs(ref_loss) < TOL
    grads = store.grads_numpy()
    assert rel_err(grads['DNNDecoder/text/outlayer/weights'], glin['weights']) < TOL
    assert rel_err(grads['DNNDecoder/text/outlayer/biases'], glin['biases']) < TOL
    for l in range(NL):
        base = 'DBLSTM/features/layer%d/bidirectional_rnn/%%s/layer_norm_basic_lstm_cell/%%s' % l
        for d in ('fw', 'bw'):
            for k in ('kernel', 'bias'):
                assert rel_err(grads[base % (d, k)], glayers[l]['%s_%s' % (d, k)]) < 5 * TOL, (l, d, k)
    # update: clip(-1,1) + TF-Adam step 1 at lr0 (global_step 0)
    assert abs(lr - 1e-3) < 1e-12
    new = store.to_numpy()
    name = 'DBLSTM/features/layer0/bidirectional_rnn/fw/layer_norm_basic_lstm_cell/kernel'
    th, _, _ = O.tf_adam_clip(params[name], glayers[0]['fw_kernel'], 0 * params[name], 0 * params[name], 1e-3, 1,
                              dtype=np.float64)
    # Adam's first step is lr*sign(g) wherever |g| >> eps; compare where the oracle gradient is not ~0
    big = np.abs(glayers[0]['fw_kernel']) > 1e-4
    assert np.abs(new[name] - th)[big].max() < 2e-6
    assert tr.global_step == 1


def test_logits_match_oracle_full_length():
    dev = torch.device('cuda', 0)
    B, T, D, H, NL, V = 16, 120, 40, 128, 3, 29
    tr = _ctc_trainer(H, NL, V, dev, seed=11)
    tr.model.build({'features': D}, dev)
    x, lens, labels, ll = synthetic_ctc_batch(B, T, D, V, ragged=False)
    params = tr.model.store.to_numpy()
    with torch.no_grad():
        logits, out_lens = tr.model({'features': torch.from_numpy(x).to(dev)},
                                    {'features': torch.from_numpy(lens).to(dev)}, None, None, False)
    layers = _dblstm_layers(params, NL)
    lin = {'weights': params['DNNDecoder/text/outlayer/weights'], 'biases': params['DNNDecoder/text/outlayer/biases']}
    enc, _, _ = O.dblstm_fwd(x, lens, layers)
    ref = O.linear_fwd(enc, lin)
    assert rel_err(logits['text'].cpu().numpy(), ref) < TOL
    assert np.array_equal(out_lens['text'].cpu().numpy(), lens)


def test_training_reduces_loss():
    dev = torch.device('cuda', 0)
    B, T, D, H, NL, V = 8, 60, 40, 64, 2, 29
    tr = _ctc_trainer(H, NL, V, dev, seed=5)
    tr.model.build({'features': D}, dev)
    x, lens, labels, ll = synthetic_ctc_batch(B, T, D, V, ragged=True)
    batch = ({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)},
             {'text': torch.from_numpy(labels).to(dev)}, {'text': torch.from_numpy(ll).to(dev)})
    losses = [float(tr.update(*batch)[0]) for _ in range(30)]
    assert losses[-1] < 0.8 * losses[0], losses


def test_deferred_weight_gradients_change_nothing():
    """nabu_set_overlap (weight-gradient GEMMs on a side stream under the next layer's backward recurrence) is a pure
    scheduling change: two updates from the same initial state give bit-identical parameters with it on and off."""
    from nabu_b200 import engine
    dev = torch.device('cuda', 0)
    B, T, D, H, NL, V = 24, 36, 40, 512, 3, 29
    x, lens, labels, ll = synthetic_ctc_batch(B, T, D, V, ragged=True)
    batch = ({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)},
             {'text': torch.from_numpy(labels).to(dev)}, {'text': torch.from_numpy(ll).to(dev)})
    results = []
    try:
        for on in (True, False):
            tr = _ctc_trainer(H, NL, V, dev, seed=3)
            tr.model.build({'features': D}, dev)
            engine.set_overlap(on)
            losses = [float(tr.update(*batch)[0]) for _ in range(2)]
            torch.cuda.synchronize()
            results.append((losses, tr.model.store.to_numpy()))
    finally:
        engine.set_overlap(True)
    assert results[0][0] == results[1][0]
    for k in results[0][1]:
        assert np.array_equal(results[0][1][k], results[1][1][k]), k


def test_loss_evaluator_and_validation_in_train_loop(capsys):
    """Row f2: LossEvaluator's utterance-weighted validation loss against the oracle, and Trainer.train running the
    validation controller (valid_frequency, early stop after num_tries worse results)."""
    from nabu_b200.neuralnetworks.evaluators import evaluator_factory
    dev = torch.device('cuda', 0)
    D, H, NL, V = 40, 64, 2, 29
    tr = _ctc_trainer(H, NL, V, dev, seed=2)
    tr.model.build({'features': D}, dev)

    def batch(B, T, seed):
        x, lens, labels, ll = synthetic_ctc_batch(B, T, D, V, ragged=True, seed=seed)
        return (({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)},
                 {'text': torch.from_numpy(labels).to(dev)}, {'text': torch.from_numpy(ll).to(dev)}), (x, lens, labels, ll))

    val = [batch(4, 30, 1), batch(7, 24, 2), batch(3, 36, 3)]
    econf = make_conf('[evaluator]\nevaluator = loss_evaluator\nloss = CTC\ntargets = text\nbatch_size = 4\nfeatures = devfbank\n'
                      'text = devtext\n')
    ev = evaluator_factory.factory('loss_evaluator')(econf, None, tr.model, batch_source=[b[0] for b in val])
    got, n = ev.evaluate()
    params = tr.model.store.to_numpy()
    layers = _dblstm_layers(params, NL)
    lin = {'weights': params['DNNDecoder/text/outlayer/weights'], 'biases': params['DNNDecoder/text/outlayer/biases']}
    tot, cnt = 0.0, 0
    for _, (x, lens, labels, ll) in val:
        enc, _, _ = O.dblstm_fwd(x, lens, layers)
        l, _ = O.ctc_loss_mean(O.linear_fwd(enc, lin), lens, labels, ll)
        tot += l * x.shape[0]
        cnt += x.shape[0]
    assert n == 3 and abs(got - tot / cnt) / (tot / cnt) < TOL

    # the train loop: validate every 2 steps; a validation set that never improves stops training after num_tries
    class Src(list):
        input_dims = {'features': D}
    tconf = make_conf('[trainer]\ntrainer = standard\nloss = CTC\ntrainlabels = 1\ntargets = text\nnum_epochs = 50\n'
                      'valid_frequency = 2\nnum_tries = 1\ninitial_learning_rate = 0\n')
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = dblstm\nnum_units = %d\n'
                      'num_layers = %d\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n'
                      % (V - 1, H, NL))
    tr2 = trainer_factory.factory('standard')(tconf, None, mconf, econf, None, None, 0, device=dev, seed=2,
                                              batch_source=Src([val[0][0], val[1][0]]), val_source=[b[0] for b in val])
    tr2.train()
    out = capsys.readouterr().out
    # lr = 0: the validation loss repeats -> step 0 better (saved), step 2 worse (try 1), step 4 worse -> terminate
    assert out.count('validating model') == 3 and 'terminating training' in out
    assert tr2.global_step == 0                        # the terminate path restores the validated model (saved at step 0)


def test_train_validate_recognize_from_nabu_data_directories(tmp_path, capsys):
    """Rows f1 + f2 end to end: data prepared in nabu's on-disk format (TFRecord files, pointers.scp, metadata) ->
    Trainer.train with the sections named in trainer.cfg / database.conf (bucketed, variable batch size), validation by
    a loss evaluator reading its own sections, then Recognizer.recognize writing the decoded set."""
    import configparser
    from tests.test_processing import _write_stream
    from nabu_b200.neuralnetworks.recognizer import Recognizer
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(3)
    alphabet = ['a', 'b', 'c', 'd', 'e']
    D = 40

    def make_set(tag, n):
        lens = rng.integers(20, 60, size=n)
        feats = [('%s%d' % (tag, i), rng.standard_normal((L, D)).astype(np.float32)) for i, L in enumerate(lens)]
        texts = [('%s%d' % (tag, i), ' '.join(rng.choice(alphabet, size=max(1, L // 12)))) for i, L in enumerate(lens)]
        fdir, tdir = str(tmp_path / (tag + 'fbank')), str(tmp_path / (tag + 'text'))
        _write_stream(fdir, 'audio', feats, dim=D)
        _write_stream(tdir, 'text', texts, alphabet=alphabet)
        return fdir, tdir

    trf, trt = make_set('train', 24)
    dvf, dvt = make_set('dev', 8)
    dataconf = configparser.ConfigParser()
    dataconf.read_string('[trainfbank]\ndir = %s\ntype = audio_feature\n[traintext]\ndir = %s\ntype = string_eos\n'
                         '[devfbank]\ndir = %s\ntype = audio_feature\n[devtext]\ndir = %s\ntype = string_eos\n'
                         % (trf, trt, dvf, dvt))
    V = len(alphabet) + 1                    # + EOS (string_eos appends it); CTC adds its blank through trainlabels
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = dblstm\n'
                      'num_units = 64\nnum_layers = 2\ninput_noise = 0\ndropout = 1\n[decoder]\ndecoder = dnn_decoder\n'
                      'num_layers = 0\n' % V)
    tconf = make_conf('[trainer]\ntrainer = standard\nloss = CTC\ntrainlabels = 1\ntargets = text\nnum_epochs = 2\n'
                      'batch_size = 4\nnumbuckets = 3\nvariable_batch_size = True\nvalid_frequency = 3\nnum_tries = None\n'
                      'features = trainfbank\ntext = traintext\n')
    econf = make_conf('[evaluator]\nevaluator = loss_evaluator\nloss = CTC\ntargets = text\nbatch_size = 4\n'
                      'features = devfbank\ntext = devtext\n')
    expdir = str(tmp_path / 'exp')
    tr = trainer_factory.factory('standard')(tconf, dataconf, mconf, econf, expdir, None, 0, device=dev, seed=2)
    tr.train()                               # the evaluator reads the sections validation_evaluator.cfg names
    out = capsys.readouterr().out
    assert tr.global_step == tr.num_steps and tr.num_steps > 0
    assert out.count('validating model') >= 2 and 'validation loss' in out
    assert os.path.isfile(os.path.join(expdir, 'model', 'network.pt'))
    rconf = make_conf('[recognizer]\nbatch_size = 3\nfeatures = devfbank\n[decoder]\ndecoder = ctc_decoder\n'
                      'text_alphabet = %s\n' % ' '.join(alphabet + ['<eos>']))
    directory = Recognizer(tr.model, rconf, dataconf, expdir).recognize()
    text = open(os.path.join(directory, 'text')).read()
    lines = text.strip().split('\n')
    assert len(lines) == 8 and sorted(l.split(' ')[0] for l in lines) == sorted('dev%d' % i for i in range(8))
    # row f3: the TF checkpoint SaveAtEnd wrote carries the reference's variable names; a model that has no variables
    # yet (another seed) restores from it inside Recognizer.recognize and decodes the same text
    from nabu_b200.neuralnetworks.models.model import Model
    from nabu_b200.processing import tfcheckpoint
    saved = dict((n, s) for n, s, _ in tfcheckpoint.list_variables(os.path.join(expdir, 'model', 'network.ckpt')))
    assert saved['DNNDecoder/text/outlayer/weights'] == (128, V + 1)
    assert saved['DBLSTM/features/layer1/bidirectional_rnn/bw/layer_norm_basic_lstm_cell/kernel'] == (128 + 64, 256)
    os.remove(os.path.join(expdir, 'model', 'network.pt'))
    fresh = Model(mconf, 1, seed=99)
    fresh.device = dev
    directory = Recognizer(fresh, rconf, dataconf, expdir).recognize()
    assert open(os.path.join(directory, 'text')).read() == text
    assert torch.equal(fresh.store.theta, tr.model.store.theta)


def test_scripts_train_test_decode_on_an_experiment_directory(tmp_path, capsys):
    """The three per-experiment entry points (`run train|test|decode` end in nabu/scripts/{train,test,decode}.py):
    everything is read from the cfg files of the experiment directory; the model travels from train to test / decode
    as the TF checkpoint model/network.ckpt, as in the reference."""
    from nabu_b200.scripts import decode, test, train
    from tests.util import write_experiment
    expdir = write_experiment(str(tmp_path))
    tr = train.train(expdir, device=torch.device('cuda', 0))
    out = capsys.readouterr().out
    assert tr.global_step == tr.num_steps > 0 and 'validation loss' in out
    assert os.path.isfile(os.path.join(expdir, 'model', 'network.ckpt.index'))
    loss = test.test(expdir, device=torch.device('cuda', 0))
    assert 0.0 <= loss and float(open(os.path.join(expdir, 'result')).read()) == loss      # label error rate
    rec = decode.decode(expdir, device=torch.device('cuda', 0))
    assert torch.equal(rec.model.store.theta, tr.model.store.theta)
    lines = open(os.path.join(expdir, 'decoded', 'text')).read().strip().split('\n')
    assert sorted(l.split(' ')[0] for l in lines) == sorted('test%d' % i for i in range(6))


def test_scripts_on_a_las_experiment_directory(tmp_path, capsys):
    """The LAS/TIMIT recipe's shape end to end (listener + speller with the recipe's input noise / dropout, trained
    with average_cross_entropy on EOS-terminated targets, validated by beam-search label error rate, tested by loss,
    decoded with the beam search into n-best files) through the three entry points."""
    from nabu_b200.scripts import decode, test, train
    from tests.util import write_experiment
    expdir = write_experiment(str(tmp_path), model='las')
    tr = train.train(expdir, device=torch.device('cuda', 0))
    out = capsys.readouterr().out
    assert tr.global_step == tr.num_steps > 0 and 'validation loss' in out
    loss = test.test(expdir, device=torch.device('cuda', 0))
    assert np.isfinite(loss) and loss > 0
    rec = decode.decode(expdir, device=torch.device('cuda', 0))
    assert torch.equal(rec.model.store.theta, tr.model.store.theta)
    nbest = open(os.path.join(expdir, 'decoded', 'test0')).read().strip().split('\n')
    assert len(nbest) == 4 and all(float(l.split(' ')[0]) <= 0 for l in nbest)          # score, then the symbols
    assert np.load(os.path.join(expdir, 'decoded', 'test0_alignments.npy')).shape[0] == 4
