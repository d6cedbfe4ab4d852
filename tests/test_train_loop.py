"""Trainer.train's host logic on the CPU -- the loop, periodic checkpoints in <expdir>/logdir, resume after a crash,
the files SaveAtEnd leaves -- with the CUDA update step replaced by a stand-in (the update itself is what the GPU
parity tests check).  reference: trainers/trainer.py:582-792 (loop), 625-633 (MonitoredTrainingSession checkpoint_dir)."""
import os

import numpy as np
import pytest
import torch

from nabu_b200.neuralnetworks.trainers import standard_trainer
from nabu_b200.processing import tfcheckpoint
from nabu_b200.scripts import read_cfg
from tests.util import write_experiment


class _Crash(Exception):
    pass


class StandardTrainer(standard_trainer.StandardTrainer):      # same name: the defaults file is found by class name
    """update = a deterministic function of (global_step, batch): no kernels"""
    crash_at = None
    seen = None

    def update(self, inputs, input_seq_length, targets, target_seq_length):
        if self.crash_at is not None and self.global_step == self.crash_at:
            raise _Crash()
        frames = float(sum(int(v.sum()) for v in input_seq_length.values()))
        with torch.no_grad():
            self.model.store.theta.add_(1e-3 * (self.global_step + 1))
            self.model.store.m.add_(frames * 1e-6)
            self.model.store.v.add_(1.0)
        self.seen.append(self.global_step)
        lr = self.learning_rate()                     # the real update reads the rate before it advances the step
        self.global_step += 1
        return torch.tensor(frames), lr


def _trainer(expdir, seen, crash_at=None):
    econf = read_cfg(expdir, 'validation_evaluator.cfg')
    econf.set('evaluator', 'evaluator', 'None')               # validation runs the model: GPU tests cover it
    tr = StandardTrainer(read_cfg(expdir, 'trainer.cfg'), read_cfg(expdir, 'database.conf'), read_cfg(expdir, 'model.cfg'),
                     econf, expdir, None, 0, device='cpu')
    tr.seen, tr.crash_at = seen, crash_at
    return tr


def test_train_loop_checkpoints_and_resumes(tmp_path, monkeypatch, capsys):
    monkeypatch.setattr(standard_trainer.StandardTrainer, 'checkpoint_secs', 0.0)          # a checkpoint after every step
    expdir = write_experiment(str(tmp_path / 'a'), num_epochs=3, variable_batch_size=False)
    ref_dir = write_experiment(str(tmp_path / 'b'), num_epochs=3, variable_batch_size=False)

    # uninterrupted run
    seen_ref = []
    ref = _trainer(ref_dir, seen_ref)
    ref.train()
    assert ref.global_step == ref.num_steps == 12 and seen_ref == list(range(12))
    for name in ('network.pt', 'network.ckpt.index', 'network.ckpt.data-00000-of-00001', 'model.pkl'):
        assert os.path.isfile(os.path.join(ref_dir, 'model', name)), name
    # step-suffixed prefix like tf.train.Saver's, the state file names it, older prefixes are gone
    assert sorted(os.listdir(os.path.join(ref_dir, 'logdir'))) == ['checkpoint', 'metrics.jsonl',
                                                                  'model.ckpt-12.data-00000-of-00001', 'model.ckpt-12.index']
    assert tfcheckpoint.latest_checkpoint(os.path.join(ref_dir, 'logdir')) == os.path.join(ref_dir, 'logdir', 'model.ckpt-12')
    import json
    rows = [json.loads(l) for l in open(os.path.join(ref_dir, 'logdir', 'metrics.jsonl'))]
    assert [r['step'] for r in rows] == list(range(12)) and all(r['training_loss'] > 0 for r in rows)
    assert rows[0]['learning_rate'] == 1e-3 and rows[-1]['learning_rate'] < rows[0]['learning_rate']
    saved = dict((n, s) for n, s, _ in tfcheckpoint.list_variables(os.path.join(ref_dir, 'logdir', 'model.ckpt-12')))
    assert saved['global_step'] == () and 'learning_rate_fact' in saved
    assert any(n.endswith('/kernel/Adam_1') for n in saved)

    # the same run killed inside step 5, then started again on the same directory
    seen = []
    with pytest.raises(_Crash):
        _trainer(expdir, seen, crash_at=5).train()
    assert seen == list(range(5))
    assert int(tfcheckpoint.read_checkpoint(tfcheckpoint.latest_checkpoint(os.path.join(expdir, 'logdir')),
                                            names={'global_step'})['global_step']) == 5
    # a save that dies after the bundle is written but before the state file is switched leaves the previous
    # checkpoint in charge (the old scheme replaced .data before .index)
    open(os.path.join(expdir, 'logdir', 'model.ckpt-99.index'), 'wb').write(b'torn')
    assert tfcheckpoint.latest_checkpoint(os.path.join(expdir, 'logdir')).endswith('model.ckpt-5')
    again = _trainer(expdir, seen)
    again.train()
    assert 'resuming from step 5' in capsys.readouterr().out
    assert seen == list(range(12)) and again.global_step == 12          # no step lost, none repeated
    # variables and optimizer slots carried over the restart: theta / v see every step exactly once (m depends on
    # which batch a step got, and the position inside the epoch is not part of a checkpoint -- as in the reference)
    for var in ref.model.store.order:                # (the alignment gaps between variables are not part of a checkpoint)
        sl = slice(var.offset, var.offset + var.numel)
        assert torch.allclose(again.model.store.theta[sl], ref.model.store.theta[sl], rtol=0, atol=1e-6), var.name
        assert torch.equal(again.model.store.v[sl], ref.model.store.v[sl]), var.name
    # a finished experiment started once more does nothing but rewrite the model
    more = _trainer(expdir, seen)
    more.train()
    assert seen == list(range(12)) and more.global_step == 12


def test_validation_state_travels_with_the_checkpoint(tmp_path):
    from nabu_b200.neuralnetworks.trainers.trainer import ValidationController
    expdir = write_experiment(str(tmp_path), num_epochs=1)
    tr = _trainer(expdir, [])
    tr.train(testing=True)
    tr._controller = ValidationController(tr.conf, lambda: None, lambda: None, lambda: None)
    tr._controller.validated_step, tr._controller.best_validation, tr._controller.num_tries = 6, 1.25, 1
    tr.global_step, tr.learning_rate_fact, tr.should_terminate = 7, 0.25, True
    tr.save_checkpoint()
    other = _trainer(expdir, [])
    other.train(testing=True)
    other._controller = ValidationController(other.conf, lambda: None, lambda: None, lambda: None)
    assert other.restore_checkpoint()
    assert (other.global_step, other.learning_rate_fact, other.should_terminate) == (7, 0.25, True)
    other.train()                                   # stopped early before: no further step is taken
    assert other.global_step == 7 and other.seen == []
    assert (other._controller.validated_step, other._controller.best_validation, other._controller.num_tries) == (6, 1.25, 1)
    assert torch.equal(other.model.store.theta, tr.model.store.theta)


def test_best_validated_snapshot_survives_a_restart(tmp_path):
    """the go-back snapshot is also <expdir>/logdir/validated.ckpt (the reference's ValidationSaveHook,
    components/hooks.py:54-86), so a resumed run that validates worse still goes back to the best parameters."""
    from nabu_b200.neuralnetworks.trainers.trainer import ValidationController
    expdir = write_experiment(str(tmp_path), num_epochs=1)
    tr = _trainer(expdir, [])
    tr.train(testing=True)
    tr._controller = ValidationController(tr.conf, tr._save_validated, tr._restore_validated, tr._half_lr)
    tr.global_step = 3
    assert tr._controller.update(2.0, 3) == 'continue'          # better than 1.79e308: saved as the best
    best = tr.model.store.theta.clone()
    assert os.path.isfile(os.path.join(expdir, 'logdir', 'validated.ckpt.index'))
    with torch.no_grad():
        tr.model.store.theta.add_(1.0)                          # training goes on ...
    tr.global_step = 6
    tr.save_checkpoint()                                        # ... and the periodic checkpoint holds the later state
    # a new process on the same directory
    other = _trainer(expdir, [])
    other.conf['go_back'] = 'True'
    other.train(testing=True)
    other._controller = ValidationController(other.conf, other._save_validated, other._restore_validated, other._half_lr)
    assert other.restore_checkpoint() and other.global_step == 6
    assert not torch.equal(other.model.store.theta, best)
    assert other._controller.update(5.0, 6) == 'continue'       # worse: go back to validated.ckpt
    for var in other.model.store.order:
        sl = slice(var.offset, var.offset + var.numel)
        assert torch.equal(other.model.store.theta[sl], best[sl]), var.name
    assert other.global_step == 3 and other._controller.best_validation == 2.0
