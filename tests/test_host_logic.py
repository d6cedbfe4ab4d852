"""CPU tests of the host-side mirror of nabu's plugin API: config defaults, factories, variable
declaration (names / shapes / initialisers), learning-rate schedule and the data-parallel
arithmetic (world_size 2 over gloo)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.util import make_conf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_apply_defaults(tmp_path):
    from nabu_b200.tools.default_conf import apply_defaults
    f = tmp_path / 'x.cfg'
    f.write_text('[default]\na = 1\nb = 2\nrequired =\n')
    with pytest.raises(Exception):
        apply_defaults({'a': '5'}, str(f))
    conf = apply_defaults({'a': '5', 'required': 'yes'}, str(f))
    assert conf == {'a': '5', 'b': '2', 'required': 'yes'}
    assert apply_defaults({'q': '1'}, str(tmp_path / 'missing.cfg')) == {'q': '1'}


def test_factories_select_by_cfg_string():
    from nabu_b200.neuralnetworks.models.ed_encoders import ed_encoder_factory as ef
    from nabu_b200.neuralnetworks.models.ed_decoders import ed_decoder_factory as df
    from nabu_b200.neuralnetworks.trainers import trainer_factory as tf_, loss_functions as lf
    from nabu_b200.neuralnetworks.decoders import decoder_factory as dcf
    assert ef.factory('listener').__name__ == 'Listener' and ef.factory('dblstm').__name__ == 'DBLSTM'
    assert df.factory('speller').__name__ == 'Speller' and df.factory('dnn_decoder').__name__ == 'DNNDecoder'
    assert tf_.factory('standard').__name__ == 'StandardTrainer'
    assert lf.factory('CTC') is lf.CTC and lf.factory('average_cross_entropy') is lf.average_cross_entropy
    assert dcf.factory('ctc_decoder').__name__ == 'CTCDecoder'
    assert dcf.factory('beam_search_decoder').__name__ == 'BeamSearchDecoder'
    for fn, bad in ((ef.factory, 'nope'), (df.factory, 'nope'), (tf_.factory, 'nope'), (lf.factory, 'nope'),
                    (dcf.factory, 'nope'), (ef.factory, 'dnn'), (dcf.factory, 'max_decoder')):
        with pytest.raises(Exception):
            fn(bad)


LAS_CONF = ('[io]\ninputs = features\noutputs = text\noutput_dims = 39\n[encoder]\nencoder = listener\n'
            'num_units = 128\nnum_layers = 2\n[decoder]\ndecoder = speller\nnum_units = 128\n'
            'attention = location_aware\nnumfilt = 10\nfiltersize = 201\n')


def test_model_declares_tf_named_variables():
    from nabu_b200.neuralnetworks.models.model import Model
    m = Model(make_conf(LAS_CONF), trainlabels=1)
    assert m.output_dims == {'text': 40}
    assert m.encoder.conf['pyramid_steps'] == '2' and m.decoder.conf['sample_prob'] == '0.1'   # defaults merged
    m.build({'features': 40}, 'cpu')
    names = {v.name: v.shape for v in m.store.order}
    k = 'Listener/features/layer0/BLSTM/bidirectional_rnn/fw/layer_norm_basic_lstm_cell/kernel'
    assert names[k] == (40 + 128, 512)
    assert names['Listener/features/layer1/BLSTM/bidirectional_rnn/bw/layer_norm_basic_lstm_cell/kernel'] == (
        512 + 128, 512)
    assert names['Listener/features/layer2/bidirectional_rnn/fw/layer_norm_basic_lstm_cell/bias'] == (512,)
    assert names['Speller/decoder/attention_wrapper/multi_rnn_cell/cell_0/lstm_cell/kernel'] == (40 + 256 + 128, 512)
    assert names['Speller/decoder/attention_wrapper/location_aware_attention/conv1d/kernel'] == (201, 1, 10)
    assert names['Speller/decoder/dense/kernel'] == (128 + 256, 40)
    # initialisers: LayerNormBasicLSTMCell bias is glorot (non-zero), LSTMCell / dense biases are zero
    p = m.store.to_numpy()
    assert np.abs(p['Listener/features/layer2/bidirectional_rnn/fw/layer_norm_basic_lstm_cell/bias']).max() > 0
    assert np.abs(p['Speller/decoder/dense/bias']).max() == 0
    lim = np.sqrt(6.0 / (168 + 512))
    assert np.abs(p[k]).max() <= lim and np.abs(p[k]).max() > 0.9 * lim
    assert len(m.variables) == len(m.store.order)
    # every variable 256-byte aligned inside the flat buffer, grads are views of ONE flat buffer
    assert all(v.offset % 64 == 0 for v in m.store.order)
    assert m.store.order[3].grad.data_ptr() == m.store.grad.data_ptr() + 4 * m.store.order[3].offset


def test_learning_rate_schedule_matches_exponential_decay():
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    tconf = make_conf('[trainer]\ntrainer = standard\nloss = CTC\ntargets = t\ninitial_learning_rate = 0.01\n'
                      'learning_rate_decay = 0.1\n')
    mconf = make_conf('[io]\ninputs = f\noutputs = t\noutput_dims = 3\n[encoder]\nencoder = dblstm\n'
                      '[decoder]\ndecoder = dnn_decoder\nnum_layers = 0\n')
    tr = trainer_factory.factory('standard')(tconf, None, mconf, None, None, None, 0, device='cpu')
    tr.num_steps = 200
    tr.global_step = 50
    assert abs(tr.learning_rate() - 0.01 * 0.1 ** 0.25) < 1e-12
    tr.learning_rate_fact = 0.5
    assert abs(tr.learning_rate() - 0.005 * 0.1 ** 0.25) < 1e-12
    with pytest.raises(Exception):
        tr.train()            # no batch source: the TFRecord pipeline (row f1) is not built


def test_data_parallel_gradients_world_size_2_gloo():
    """Rank r takes utterances r::2, scales by 1/world after a SUM all-reduce: the result must equal
    the global-batch gradient (trainer.update's arithmetic, checked with the oracle on CPU)."""
    script = os.path.join(ROOT, 'tests', 'dp_worker.py')
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29533', script],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'DP_OK' in r.stdout
