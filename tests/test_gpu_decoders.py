"""GPU parity tests for the decoders (rows a12, a13): ids must be bit-exact against the oracle."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from tests.util import make_conf, synthetic_ctc_batch, synthetic_las_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,T,V,W,merge,scale', [
    (3, 12, 5, 4, True, 2.0),
    (4, 40, 8, 10, False, 1.0),
    (8, 200, 29, 100, True, 1.0),       # the reference's defaults: beam 100, merge_repeated
    (2, 60, 29, 100, True, 4.0),        # peaky posteriors (trained-model-like)
    (2, 30, 3, 100, True, 1.0),         # beam wider than the number of prefixes that exist early on
])
def test_ctc_beam_search_ids_bit_exact(B, T, V, W, merge, scale):
    from nabu_b200 import engine
    rng = np.random.default_rng(B * 7 + T)
    logits = (rng.standard_normal((B, T, V)) * scale).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    lens[0] = T
    ids, out_len, nlp = engine.ctc_beam_search(torch.tensor(logits).cuda(), torch.tensor(lens).cuda(), W, merge)
    ids, out_len, nlp = ids.cpu().numpy(), out_len.cpu().numpy(), nlp.cpu().numpy()
    for b in range(B):
        ref, ref_nlp = O.ctc_beam_search(logits[b], lens[b], W, merge)
        assert out_len[b] == len(ref), (b, out_len[b], len(ref))
        assert np.array_equal(ids[b, :out_len[b]], ref), b              # label ids: bit-exact
        assert abs(nlp[b] - ref_nlp) < 1e-3 * max(1.0, abs(ref_nlp))


def test_ctc_decoder_plugin(tmp_path):
    """CTCDecoder(conf, model): __call__ -> sparse int32 ids, write(), update_evaluation_loss()."""
    from nabu_b200.neuralnetworks.decoders import decoder_factory
    from nabu_b200.neuralnetworks.decoders.decoder import RunningLoss
    from nabu_b200.neuralnetworks.models.model import Model
    dev = torch.device('cuda', 0)
    B, T, D, H, NL, V = 5, 50, 40, 64, 2, 9
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = dblstm\n'
                      'num_units = %d\nnum_layers = %d\n[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n'
                      % (V - 1, H, NL))
    model = Model(mconf, 1, None, seed=4).build({'features': D}, dev)
    alphabet = ' '.join('s%d' % i for i in range(V - 1))
    dconf = make_conf('[decoder]\ndecoder = ctc_decoder\ntext_alphabet = %s\n' % alphabet)
    dec = decoder_factory.factory('ctc_decoder')(dconf, model)
    x, lens, labels, ll = synthetic_ctc_batch(B, T, D, V, ragged=True)
    out = dec({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)})
    # oracle: same weights -> logits -> TF prefix beam search
    p = model.store.to_numpy()
    layers = []
    for l in range(NL):
        base = 'DBLSTM/features/layer%d/bidirectional_rnn/%%s/layer_norm_basic_lstm_cell/%%s' % l
        layers.append({'fw_kernel': p[base % ('fw', 'kernel')], 'fw_bias': p[base % ('fw', 'bias')],
                       'bw_kernel': p[base % ('bw', 'kernel')], 'bw_bias': p[base % ('bw', 'bias')]})
    enc, _, _ = O.dblstm_fwd(x, lens, layers, np.float32)
    logits = O.linear_fwd(enc, {'weights': p['DNNDecoder/text/outlayer/weights'],
                                'biases': p['DNNDecoder/text/outlayer/biases']}, np.float32)
    sp = out['text']
    assert sp.values.dtype == np.int32 and sp.dense_shape[0] == B
    errors = 0
    for b in range(B):
        ref, _ = O.ctc_beam_search(logits[b], lens[b], 100, True)
        got = sp.values[sp.indices[:, 0] == b]
        assert np.array_equal(got, ref), b
        errors += O.edit_distance(ref, labels[b, :ll[b]])
    dec.write(out, str(tmp_path), ['utt%d' % i for i in range(B)])
    lines = open(os.path.join(str(tmp_path), 'text')).read().strip().split('\n')
    assert len(lines) == B and lines[0].startswith('utt0')
    loss = RunningLoss()
    v = dec.update_evaluation_loss(loss, out, {'text': torch.from_numpy(labels)}, {'text': torch.from_numpy(ll)})
    assert abs(v - errors / float(ll.sum())) < 1e-9


def test_beam_search_decoder_plugin(tmp_path):
    from nabu_b200.neuralnetworks.decoders import decoder_factory
    from nabu_b200.neuralnetworks.decoders.decoder import RunningLoss
    from nabu_b200.neuralnetworks.models.model import Model
    from tests.test_gpu_speller import las_oracle_params
    dev = torch.device('cuda', 0)
    B, T, D, H, NL, V, U = 3, 41, 40, 64, 2, 8, 6
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = listener\n'
                      'num_units = %d\nnum_layers = %d\npyramid_steps = 2\n[decoder]\ndecoder = speller\n'
                      'num_layers = 2\nnum_units = 32\nattention = location_aware\nnumfilt = 3\nfiltersize = 5\n'
                      % (V - 1, H, NL))
    model = Model(mconf, 1, None, seed=8).build({'features': D}, dev)
    alphabet = ' '.join('s%d' % i for i in range(V))
    dconf = make_conf('[decoder]\ndecoder = beam_search_decoder\nmax_steps = 12\nbeam_width = 4\nalphabet = %s\n'
                      % alphabet)
    dec = decoder_factory.factory('beam_search_decoder')(dconf, model)
    x, lens, targets, tl = synthetic_las_batch(B, T, D, V, U, ragged=True)
    out = dec({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)})
    seqs, lengths, scores, aligns = [t.cpu().numpy() for t in out['text']]
    layers, sp = las_oracle_params(model.store.to_numpy(), NL)
    enc, elens, _ = O.listener_fwd(x, lens, layers, 2, np.float32)
    ref = O.las_beam_search(enc, elens, sp, 4, 12, 'location_aware', 2, 1.0, 1.0, np.float32)
    assert seqs.shape == ref[0].shape and np.array_equal(seqs, ref[0]) and np.array_equal(lengths, ref[1])
    assert aligns.shape == ref[3].shape
    dec.write(out, str(tmp_path), ['u%d' % i for i in range(B)])
    assert len(open(os.path.join(str(tmp_path), 'u0')).read().strip().split('\n')) == 4
    assert os.path.exists(os.path.join(str(tmp_path), 'u0_alignments.npy'))
    loss = RunningLoss()
    v = dec.update_evaluation_loss(loss, out, {'text': torch.from_numpy(targets)}, {'text': torch.from_numpy(tl)})
    err = sum(O.edit_distance(ref[0][i, 0, :ref[1][i, 0]], targets[i, :tl[i] - 1]) for i in range(B))
    assert abs(v - err / float(tl.sum())) < 1e-9


def test_recognizer_and_decoder_evaluator(tmp_path):
    """Row f2: Recognizer.recognize writes one line per utterance through the decoder (names with the pipeline's index
    suffix cut off); DecoderEvaluator's running label error rate equals the decoder's own evaluation loss."""
    from nabu_b200.neuralnetworks.evaluators import evaluator_factory
    from nabu_b200.neuralnetworks.models.model import Model
    from nabu_b200.neuralnetworks.recognizer import Recognizer
    dev = torch.device('cuda', 0)
    D, H, NL, V = 40, 64, 2, 9
    mconf = make_conf('[io]\ninputs = features\noutputs = text\noutput_dims = %d\n[encoder]\nencoder = dblstm\n'
                      'num_units = %d\nnum_layers = %d\n[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n'
                      % (V - 1, H, NL))
    model = Model(mconf, 1, None, seed=4).build({'features': D}, dev)
    alphabet = ' '.join('s%d' % i for i in range(V - 1))
    batches, names = [], []
    for i, B in enumerate((4, 3)):
        x, lens, labels, ll = synthetic_ctc_batch(B, 40, D, V, ragged=True, seed=10 + i)
        batches.append(({'features': torch.from_numpy(x).to(dev)}, {'features': torch.from_numpy(lens).to(dev)},
                        {'text': torch.from_numpy(labels)}, {'text': torch.from_numpy(ll)}))
        names += ['spk%d-utt%d-%d' % (i, j, len(names) + j) for j in range(B)]
    rconf = make_conf('[recognizer]\nbatch_size = 4\nfeatures = testfbank\n[decoder]\ndecoder = ctc_decoder\n'
                      'text_alphabet = %s\n' % alphabet)
    rec = Recognizer(model, rconf, None, str(tmp_path), batch_source=[b[:2] for b in batches], names=names)
    directory = rec.recognize()
    lines = open(os.path.join(directory, 'text')).read().strip().split('\n')
    assert len(lines) == 7 and lines[0].split(' ')[0] == 'spk0-utt0' and lines[4].split(' ')[0] == 'spk1-utt0'
    # decoder evaluator == the decoder's own update_evaluation_loss over the same batches
    econf = make_conf('[evaluator]\nevaluator = decoder_evaluator\ntargets = text\nbatch_size = 4\n[decoder]\n'
                      'decoder = ctc_decoder\ntext_alphabet = %s\n' % alphabet)
    ev = evaluator_factory.factory('decoder_evaluator')(econf, None, model, batch_source=batches)
    val, n = ev.evaluate()
    from nabu_b200.neuralnetworks.decoders.decoder import RunningLoss
    loss = RunningLoss()
    for b in batches:
        ref = rec.decoder.update_evaluation_loss(loss, rec.decoder(b[0], b[1]), b[2], b[3])
    assert n == 2 and abs(val - ref) < 1e-12 and 0.0 < val
