"""Trainer (reference: nabu/neuralnetworks/trainers/trainer.py).

Only the update step is the hot path: model forward, loss, backward,
one gradient all-reduce when world_size > 1, elementwise clip(-1,1) + TF-Adam on the flat buffers,
`exponential_decay` learning rate.  The TF session / parameter-server / queue machinery of the
reference (trainer.py:73-510) is replaced by a plain loop over a batch source: one process per
GPU, synchronous data parallelism -- numerically the reference's `non_distributed` run at the
global batch size.  The chief's validation / early-stopping branch (row f2) is `ValidationController`; data come
from a batch source or from the database sections the cfg names (row f1); the final model is saved as a TF checkpoint
under the reference's variable names (row f3).
"""
import os
import time
from abc import ABCMeta, abstractmethod

import torch
import torch.distributed as dist

from ... import engine, parallel
from ...tools.default_conf import apply_defaults, defaults_path
from ..models.model import Model
from . import loss_functions


class ValidationController(object):
    """The chief's validation / early-stopping branch of Trainer.train (reference: trainers/trainer.py:189-265 for the
    state, 646-733 for the control flow) -- SURVEY.md section 8 row f2.

    State = the reference's graph variables: `validated_step` (initialised to -valid_frequency), `best_validation`
    (1.79e308) and the Python-side `num_tries`.  `save` / `restore` are the ValidationSaveHook's (hooks.py:54-86): the
    reference checkpoints ALL global variables, so a restore also winds back global_step, learning_rate_fact,
    validated_step and best_validation -- the callbacks given here must do the same.
    """

    def __init__(self, conf, save, restore, half_lr):
        self.valid_frequency = int(conf['valid_frequency'])
        self.valid_adapt = conf['valid_adapt'] == 'True'
        self.go_back = conf['go_back'] == 'True'
        self.max_tries = None if conf['num_tries'] == 'None' else int(conf['num_tries'])
        self.reset_tries = conf['reset_tries'] == 'True'
        self.validated_step = -self.valid_frequency
        self.best_validation = 1.79e+308
        self.num_tries = 0
        self._save, self._restore, self._half_lr = save, restore, half_lr

    def should_validate(self, global_step):
        return global_step - self.validated_step >= self.valid_frequency          # trainer.py:204-206

    def state(self):
        return {'validated_step': self.validated_step, 'best_validation': self.best_validation}

    def load_state(self, st):
        self.validated_step, self.best_validation = st['validated_step'], st['best_validation']

    def update(self, validation_loss, global_step):
        """Fold one validation result in.  Returns 'terminate' or 'continue' (trainer.py:680-733)."""
        if validation_loss >= self.best_validation:                               # worse (or equal)
            if self.max_tries is not None and self.num_tries == self.max_tries:
                self._restore()
                return 'terminate'
            self.num_tries += 1
            if self.go_back:
                self._restore()
            else:
                self.validated_step = global_step
            if self.valid_adapt:
                self._half_lr()
                self._save()
        else:
            if self.reset_tries:
                self.num_tries = 0
            self.validated_step = global_step
            self.best_validation = validation_loss
            self._save()
        return 'continue'


class Trainer(object, metaclass=ABCMeta):
    """Trainer(conf, dataconf, modelconf, evaluatorconf, expdir, server, task_index)."""

    def __init__(self, conf, dataconf, modelconf, evaluatorconf, expdir, server=None, task_index=0,
                 batch_source=None, device=None, seed=0, val_source=None):
        self.conf = dict(conf.items('trainer'))
        apply_defaults(self.conf, os.path.join(os.path.dirname(os.path.realpath(__file__)), 'defaults',
                                               type(self).__name__.lower() + '.cfg'))
        self.dataconf, self.evaluatorconf = dataconf, evaluatorconf
        self.expdir, self.server, self.task_index = expdir, server, task_index
        if self.conf.get('norm_constraint', 'None') != 'None':
            raise Exception('norm_constraint (MaxNorm) is outside the B200 hot path (SURVEY.md section 8 f4)')
        if int(self.conf['cut_sequence_length']) > 0:
            raise Exception('cut_sequence_length (TBPTT) is outside the B200 hot path (SURVEY.md section 8 f4)')
        self.model = Model(conf=modelconf, trainlabels=int(self.conf['trainlabels']), constraint=None, seed=seed)
        self.loss_fn = loss_functions.factory(self.conf['loss'])
        self.batch_source = batch_source
        self.val_source = val_source
        self.device = torch.device(device if device is not None else 'cuda')
        self.global_step = 0
        self.should_terminate = False
        self.learning_rate_fact = 1.0
        self.num_steps = None      # steps per epoch * num_epochs, set by train()
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # weight-gradient GEMMs of a BLSTM layer overlap the next layer's backward recurrence (NABU_OVERLAP=0 disables)
        if self.device.type == 'cuda' and os.environ.get('NABU_OVERLAP', '1') != '0':
            engine.set_overlap(True)

    # ---- learning rate: trainer.py:161-166 ------------------------------------------------------
    def learning_rate(self):
        lr0 = float(self.conf['initial_learning_rate'])
        decay = float(self.conf['learning_rate_decay'])
        total = self.num_steps if self.num_steps else 1
        return lr0 * decay ** (float(self.global_step) / float(total)) * self.learning_rate_fact

    # ---- the hot path: one minibatch update (trainer.py:169-174, 512-580, 757-767, 787) ---------
    def update(self, inputs, input_seq_length, targets, target_seq_length):
        """Returns (loss tensor on device, learning rate used).  No host synchronisation."""
        model = self.model
        logits, logit_seq_length = model(inputs, input_seq_length, targets, target_seq_length, True)
        loss = self.loss_fn(targets, logits, logit_seq_length, target_seq_length)
        extra = self.aditional_loss()
        if extra is not None:
            loss = loss + extra
        loss.backward()
        engine.side_join()                               # deferred weight gradients land before the reduction
        parallel.allreduce_sum_(model.store.grad)        # the step's only collective
        lr = self.learning_rate()
        # clip AFTER the reduction so the update equals the reference's at the global batch size
        engine.clip_adam_step(model.store, lr, self.global_step + 1, clip=1.0, grad_scale=1.0 / self.world)
        self.global_step += 1
        return loss.detach(), lr

    def train(self, testing=False):
        """Loop over the batch source (trainer.py:582-792).  `testing` builds the model and returns."""
        src = self.batch_source
        if src is None:
            # the reference's _data (trainer.py:286-415): sections named in the trainer cfg for every model input and
            # every target, shuffled file queue, bucketing by the first input's length, variable batch size
            if self.dataconf is None:
                raise Exception('Trainer.train needs a batch_source or a database configuration')
            from ...processing import input_pipeline
            src = input_pipeline.source_from_conf(
                self.conf, self.dataconf, self.model.input_names, [t for t in self.conf['targets'].split(' ') if t],
                device=self.device, numbuckets=int(self.conf['numbuckets']),
                variable_batch_size=self.conf['variable_batch_size'] == 'True', shuffle_seed=0,
                rank=dist.get_rank() if self.world > 1 else 0, world=self.world)
            self.batch_source = src
        steps_per_epoch = len(src)
        self.num_steps = steps_per_epoch * int(self.conf['num_epochs'])
        self.model.build(src.input_dims, self.device)
        if testing:
            return
        # validation (trainer.py:189-265, 646-733): evaluator named in the evaluator cfg, run by the chief every
        # valid_frequency steps on `val_source`
        evaluator, controller = None, None
        if (self.evaluatorconf is not None and (self.val_source is not None or self.dataconf is not None)
                and self.evaluatorconf.get('evaluator', 'evaluator') != 'None'):
            from ..evaluators import evaluator_factory
            evaluator = evaluator_factory.factory(self.evaluatorconf.get('evaluator', 'evaluator'))(
                self.evaluatorconf, self.dataconf, self.model, batch_source=self.val_source)
            controller = ValidationController(self.conf, self._save_validated, self._restore_validated, self._half_lr)
            self._controller = controller
        # MonitoredTrainingSession(checkpoint_dir=<expdir>/logdir) (trainer.py:625-633): every global variable is saved
        # there periodically and training picks up from it when the directory already holds a checkpoint
        if self.restore_checkpoint():
            print('WORKER %d: resuming from step %d (%s)' % (self.task_index, self.global_step, getattr(self, '_restored_from', self._checkpoint_path())))
        last_save = time.time()
        if self.device.type == 'cuda':
            used_mb = lambda: torch.cuda.max_memory_allocated() / 1e6
            total_mb = torch.cuda.get_device_properties(self.device).total_memory / 1e6
        else:
            used_mb, total_mb = (lambda: 0), 0
        terminated = self.should_terminate          # an early-stopped experiment stays stopped when started again
        while self.global_step < self.num_steps and not terminated:
            for batch in src:
                if self.global_step >= self.num_steps:
                    break
                if controller is not None and controller.should_validate(self.global_step):
                    print('WORKER %d: validating model' % self.task_index)
                    validation_loss, _ = evaluator.evaluate()
                    print('WORKER %d: validation loss: %f' % (self.task_index, validation_loss))
                    self._log_scalars(step=self.global_step, validation_loss=float(validation_loss))
                    if validation_loss >= controller.best_validation:
                        print('WORKER %d: validation loss is worse' % self.task_index)
                    if controller.update(validation_loss, self.global_step) == 'terminate':
                        print('WORKER %d: terminating training' % self.task_index)
                        terminated = self.should_terminate = True
                        break
                start = time.time()
                loss, lr = self.update(*batch)
                loss_v = float(loss)          # the reference fetches the loss every step as well
                frames = int(sum(int(v.sum()) for v in batch[1].values()))
                elapsed = time.time() - start
                print('WORKER %d: step %d/%d loss: %f, learning rate: %f \n\t time elapsed: %f sec'
                      '\n\t %.0f frames/sec, peak memory usage: %d/%d MB'
                      % (self.task_index, self.global_step - 1, self.num_steps, loss_v, lr, elapsed,
                         frames / max(elapsed, 1e-9), used_mb(), total_mb))
                self._log_scalars(step=self.global_step - 1, training_loss=loss_v, learning_rate=lr,
                                  frames_per_sec=frames / max(elapsed, 1e-9))
                if time.time() - last_save >= self.checkpoint_secs:
                    self.save_checkpoint()
                    last_save = time.time()
        self.save_checkpoint()
        if self.expdir and self.task_index == 0:
            os.makedirs(os.path.join(self.expdir, 'model'), exist_ok=True)
            torch.save(self.model.store.state_dict(), os.path.join(self.expdir, 'model', 'network.pt'))
            # SaveAtEnd (components/hooks.py:30-52, trainer.py:615-619): the model variables as a TF checkpoint under
            # the reference's variable names -- readable by a nabu / TF-1.8 install and by Recognizer below
            self.model.store.save_tf_checkpoint(os.path.join(self.expdir, 'model', 'network.ckpt'))
            self.model.save(os.path.join(self.expdir, 'model', 'model.pkl'))        # trainer.py:790-792

    def _log_scalars(self, **scalars):
        """the reference's tf.summary scalars (training_loss trainer.py:541, learning rate :269, validation loss :258,
        written by FileWriter(<expdir>/logdir) :636) as one JSON object per line in <expdir>/logdir/metrics.jsonl"""
        if not self.expdir or self.task_index != 0:
            return
        import json
        os.makedirs(os.path.join(self.expdir, 'logdir'), exist_ok=True)
        with open(os.path.join(self.expdir, 'logdir', 'metrics.jsonl'), 'a') as fid:
            fid.write(json.dumps(scalars) + '\n')

    # ---- checkpoint / resume: <expdir>/logdir/model.ckpt, a TF bundle with the Adam slots and the trainer's scalars ----
    checkpoint_secs = float(os.environ.get('NABU_CHECKPOINT_SECS', '600'))     # TF's save_checkpoint_secs default

    def _checkpoint_path(self):
        return os.path.join(self.expdir, 'logdir', 'model.ckpt') if self.expdir else None

    def save_checkpoint(self):
        """chief only.  Crash-safe the way tf.train.Saver is: the bundle goes to a NEW step-suffixed prefix
        (`model.ckpt-<step>`), the `checkpoint` state file is switched to it last and atomically, the previous prefix is
        deleted afterwards -- a run killed at any point leaves a state file that names a complete checkpoint."""
        path = self._checkpoint_path()
        if path is None or self.task_index != 0:
            return
        import glob
        import numpy as np
        from ...processing import tfcheckpoint
        extra = {'learning_rate_fact': np.array(self.learning_rate_fact, np.float32),
                 'should_terminate': np.array(bool(self.should_terminate))}          # trainer.py:97-104
        controller = getattr(self, '_controller', None)
        if controller is not None:
            extra['validated_step'] = np.array(controller.validated_step, np.int64)
            extra['best_validation'] = np.array(controller.best_validation, np.float64)
            extra['num_tries'] = np.array(controller.num_tries, np.int64)
        prefix = '%s-%d' % (path, self.global_step)
        self.model.store.save_tf_checkpoint(prefix, with_adam=True, global_step=self.global_step, extra=extra,
                                            state_file=False)
        tfcheckpoint.write_state_file(prefix)
        for old in glob.glob(path + '-*') + glob.glob(path + '.*'):
            if not old.startswith(prefix + '.'):
                os.remove(old)

    def restore_checkpoint(self):
        """every rank; returns whether a checkpoint was found.  Like the reference's session restore it brings back the
        variables, the optimizer slots, global_step and the validation state; the position inside the epoch is not
        part of it (the reference's input queues restart as well)."""
        path = self._checkpoint_path()
        if path is None:
            return False
        from ...processing import tfcheckpoint
        path = tfcheckpoint.latest_checkpoint(os.path.dirname(path))
        if path is None:
            return False
        self._restored_from = path
        step = self.model.store.load_tf_checkpoint(path, with_adam=True)
        have = set(n for n, _, _ in tfcheckpoint.list_variables(path))
        names = [n for n in ('learning_rate_fact', 'validated_step', 'best_validation', 'num_tries', 'should_terminate')
                 if n in have]
        scalars = tfcheckpoint.read_checkpoint(path, names=set(names)) if names else {}
        self.global_step = int(step) if step is not None else 0
        if 'learning_rate_fact' in scalars:
            self.learning_rate_fact = float(scalars['learning_rate_fact'])
        if 'should_terminate' in scalars:
            self.should_terminate = bool(scalars['should_terminate'])
        controller = getattr(self, '_controller', None)
        if controller is not None and 'validated_step' in scalars:
            controller.validated_step = int(scalars['validated_step'])
            controller.best_validation = float(scalars['best_validation'])
            controller.num_tries = int(scalars['num_tries'])
        self._load_validated_from_disk()
        return True

    # ---- ValidationSaveHook (hooks.py:54-86): every global variable, in memory AND as <expdir>/logdir/validated.ckpt ----
    def _validated_path(self):
        return os.path.join(self.expdir, 'logdir', 'validated.ckpt') if self.expdir else None

    def _save_validated(self):
        st = self.model.store
        self._validated = {'theta': st.theta.clone(), 'm': st.m.clone(), 'v': st.v.clone(),
                           'global_step': self.global_step, 'learning_rate_fact': self.learning_rate_fact,
                           'controller': self._controller.state()}
        # on disk as well (the reference's validated.ckpt): go_back / early stopping must still find the best-validated
        # parameters after a crash and resume
        path = self._validated_path()
        if path is not None and self.task_index == 0:
            import numpy as np
            os.makedirs(os.path.dirname(path), exist_ok=True)
            cs = self._controller.state()
            extra = {'learning_rate_fact': np.array(self.learning_rate_fact, np.float32),
                     'validated_step': np.array(cs['validated_step'], np.int64),
                     'best_validation': np.array(cs['best_validation'], np.float64)}
            st.save_tf_checkpoint(path + '.tmp', with_adam=True, global_step=self.global_step, extra=extra)
            for suffix in ('.data-00000-of-00001', '.index'):
                os.replace(path + '.tmp' + suffix, path + suffix)

    def _load_validated_from_disk(self):
        """validated.ckpt -> the in-memory snapshot (after a resume); False when there is none"""
        path = self._validated_path()
        if path is None or not os.path.isfile(path + '.index'):
            return False
        from ...processing import tfcheckpoint
        st = self.model.store
        keep = (st.theta.clone(), st.m.clone(), st.v.clone())
        step = st.load_tf_checkpoint(path, with_adam=True)
        sc = tfcheckpoint.read_checkpoint(path, names={'learning_rate_fact', 'validated_step', 'best_validation'})
        self._validated = {'theta': st.theta.clone(), 'm': st.m.clone(), 'v': st.v.clone(),
                           'global_step': int(step) if step is not None else 0,
                           'learning_rate_fact': float(sc['learning_rate_fact']),
                           'controller': {'validated_step': int(sc['validated_step']),
                                          'best_validation': float(sc['best_validation'])}}
        st.theta.copy_(keep[0]); st.m.copy_(keep[1]); st.v.copy_(keep[2])
        return True

    def _restore_validated(self):
        snap = getattr(self, '_validated', None)
        if snap is None and self._load_validated_from_disk():
            snap = self._validated
        if snap is None:
            # nothing was ever better than the initial best_validation (1.79e308), so nothing was saved: the reference
            # fails here looking for validated.ckpt; say so instead of silently keeping the current parameters
            print('WORKER %d: no validated checkpoint to go back to; keeping the current parameters' % self.task_index)
            return
        st = self.model.store
        st.theta.copy_(snap['theta']); st.m.copy_(snap['m']); st.v.copy_(snap['v'])
        self.global_step = snap['global_step']
        self.learning_rate_fact = snap['learning_rate_fact']
        self._controller.load_state(snap['controller'])

    def _half_lr(self):
        self.learning_rate_fact = self.learning_rate_fact / 2.0

    @abstractmethod
    def aditional_loss(self):
        """an extra loss term or None"""

    @abstractmethod
    def chief_only_hooks(self, outputs):
        """kept for API compatibility (session hooks have no equivalent here)"""

    @abstractmethod
    def hooks(self, outputs):
        """kept for API compatibility"""
