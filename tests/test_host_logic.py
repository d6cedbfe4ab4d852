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


# ---- row f2: the validation / early-stopping controller (reference trainers/trainer.py:189-265, 646-733) ----------
def _controller(**kw):
    from nabu_b200.neuralnetworks.trainers.trainer import ValidationController
    conf = {'valid_frequency': '10', 'valid_adapt': 'False', 'go_back': 'False', 'num_tries': '2', 'reset_tries': 'True'}
    conf.update({k: str(v) for k, v in kw.items()})
    log = []
    state = {'lr_fact': 1.0, 'saved': None}
    ctl = ValidationController(conf, save=lambda: log.append('save'), restore=lambda: log.append('restore'),
                               half_lr=lambda: (log.append('half'), state.__setitem__('lr_fact', state['lr_fact'] / 2)))
    return ctl, log, state


def test_validation_controller_early_stopping_follows_the_reference():
    ctl, log, _ = _controller()
    assert ctl.should_validate(0)                      # validated_step starts at -valid_frequency
    assert ctl.update(1.0, 0) == 'continue' and log == ['save'] and ctl.best_validation == 1.0
    assert not ctl.should_validate(9) and ctl.should_validate(10)
    assert ctl.update(0.9, 10) == 'continue' and ctl.best_validation == 0.9 and ctl.num_tries == 0
    assert ctl.update(0.95, 20) == 'continue' and ctl.num_tries == 1 and ctl.validated_step == 20
    assert ctl.update(0.9, 30) == 'continue' and ctl.num_tries == 2        # equal counts as worse (>=)
    assert log == ['save', 'save']                     # nothing saved for the two worse results
    assert ctl.update(0.99, 40) == 'terminate' and log[-1] == 'restore'   # num_tries == conf -> restore and stop
    # a better result resets the tries (reset_tries = True)
    ctl, log, _ = _controller()
    ctl.update(1.0, 0); ctl.update(1.1, 10)
    assert ctl.num_tries == 1
    ctl.update(0.5, 20)
    assert ctl.num_tries == 0 and ctl.best_validation == 0.5
    # num_tries = None disables early stopping
    ctl, log, _ = _controller(num_tries='None')
    ctl.update(1.0, 0)
    for i in range(1, 6):
        assert ctl.update(2.0, 10 * i) == 'continue'
    assert ctl.num_tries == 5


def test_validation_controller_go_back_and_lr_halving():
    ctl, log, state = _controller(go_back=True, valid_adapt=True)
    ctl.update(1.0, 0)
    assert log == ['save']
    # worse: restore the validated model, halve the learning rate, save the halved state (trainer.py:700-725)
    assert ctl.update(1.2, 10) == 'continue'
    assert log == ['save', 'restore', 'half', 'save'] and state['lr_fact'] == 0.5
    assert ctl.validated_step == 0                     # go_back does not advance validated_step (the restore winds it back)
    # without go_back the step advances and the learning rate still halves
    ctl, log, state = _controller(valid_adapt=True)
    ctl.update(1.0, 0); ctl.update(1.2, 10)
    assert log == ['save', 'half', 'save'] and ctl.validated_step == 10 and state['lr_fact'] == 0.5


def test_evaluator_factory_and_running_mean():
    from nabu_b200.neuralnetworks.evaluators import evaluator_factory, evaluator
    assert evaluator_factory.factory('loss_evaluator').__name__ == 'LossEvaluator'
    assert evaluator_factory.factory('decoder_evaluator').__name__ == 'DecoderEvaluator'
    with pytest.raises(Exception):
        evaluator_factory.factory('nope')

    class Fixed(evaluator.Evaluator):                  # utterance-weighted running mean of loss_evaluator.py:54-57
        def __init__(self, batches):
            self.batch_source = batches

        def update_loss(self, loss, batch_loss, batch_utt, *_):
            new = loss['count'] + batch_utt
            loss['loss'] = (loss['loss'] * loss['count'] + batch_loss * batch_utt) / new
            loss['count'] = new

    val, n = Fixed([(2.0, 4.0, None, None), (1.0, 12.0, None, None)]).evaluate()
    assert n == 2 and abs(val - (2.0 * 4 + 1.0 * 12) / 16) < 1e-12


def test_padding_num_units_to_the_kernel_tile_changes_nothing():
    """engine._BLSTMPadded runs num_units that are not a multiple of 64 on ceil64 units with zero weights for the extra
    ones.  The claim that this is exact (extra units stay at c = h = 0, valid units and every gradient unchanged) is
    checked here with the oracle on the same pad / unpad functions the CUDA path uses."""
    import numpy as np
    import torch
    import oracle as O
    from nabu_b200 import engine
    rng = np.random.default_rng(0)
    B, T, D, H, Hp = 3, 9, 5, 6, 64
    x = rng.standard_normal((B, T, D))
    lens = np.array([9, 6, 8])
    p = O.init_blstm_params(rng, D, H, np.float64)
    pad = lambda w, rows: engine._pad_gates(torch.from_numpy(w), H, Hp, rows).numpy()
    pp = {k: pad(v, k.endswith('kernel')) for k, v in p.items()}
    assert pp['fw_kernel'].shape == (D + Hp, 4 * Hp) and pp['bw_bias'].shape == (4 * Hp,)
    y, cache = O.blstm_fwd(x, lens, p)
    yp, cachep = O.blstm_fwd(x, lens, pp)
    assert np.all(yp[..., H:Hp] == 0) and np.all(yp[..., Hp + H:] == 0)
    assert np.array_equal(np.concatenate([yp[..., :H], yp[..., Hp:Hp + H]], -1), y)
    dy = rng.standard_normal(y.shape)
    dyp = np.zeros(yp.shape)
    dyp[..., :H], dyp[..., Hp:Hp + H] = dy[..., :H], dy[..., H:]
    dx, g = O.blstm_bwd(cache, dy)
    dxp, gp = O.blstm_bwd(cachep, dyp)
    assert np.allclose(dxp, dx, rtol=0, atol=1e-15)
    for k in g:
        got = engine._unpad_gates(torch.from_numpy(gp[k]), H, Hp, k.endswith('kernel')).numpy()
        assert np.allclose(got, g[k], rtol=0, atol=1e-15), k


def test_dy_absmax_hint_only_for_the_unmodified_dx_tensor():
    """engine._dy_absmax_hint (nabu_blstm_bwd_hints): the max |dx| a layer's backward left behind is handed to the next
    backward call only if its dy IS that dx -- same memory, same version counter (views share it, in-place writes bump
    it) -- and only once."""
    import torch
    from nabu_b200 import engine
    dx = torch.zeros(4, 6, 8)
    buf = torch.zeros(128, dtype=torch.int32)

    def store():
        engine._DXMAX['dx'], engine._DXMAX['version'], engine._DXMAX['buf'] = dx, dx._version, buf
    store()
    assert engine._dy_absmax_hint(dx.view(4, 3, 16)) is buf          # pyramid_stack's reshape: same memory
    assert engine._dy_absmax_hint(dx) is None                        # consumed
    store()
    assert engine._dy_absmax_hint(dx.clone()) is None                # another tensor
    store()
    dx.add_(1.0)                                                     # e.g. autograd accumulating a second gradient in place
    assert engine._dy_absmax_hint(dx) is None
    store()
    assert engine._dy_absmax_hint(dx[:, :3]) is None                 # part of it
    assert engine._DXMAX['dx'] is None
