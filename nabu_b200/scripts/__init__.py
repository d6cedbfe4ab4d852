"""The per-experiment entry points of nabu (reference: nabu/scripts/{train,decode,test}.py): each reads the cfg files
`run train|decode|test` copied into the experiment directory and drives the Trainer / Recognizer / Evaluator mirror.
Only these three are kept -- the `prepare_*` orchestration around them (condor, ssh tunnels, parameter servers,
`run data`) is outside the hot path."""
import configparser
import os


def read_cfg(expdir, *names):
    """ConfigParser of the first of `names` present in the experiment directory (the reference reads database.conf;
    its recipes ship database.cfg)"""
    conf = configparser.ConfigParser()
    for name in names:
        path = os.path.join(expdir, name)
        if os.path.isfile(path):
            conf.read(path)
            return conf
    raise IOError('%s: none of %s found' % (expdir, ', '.join(names)))


def load_model(expdir):
    """model/model.pkl (scripts/decode.py:46-48, scripts/test.py:41-43) -- here the pickled model DESCRIPTION that
    Trainer.train writes next to network.ckpt (a pickled TF-graph builder has no meaning outside TensorFlow); when it
    is missing, e.g. for a model trained by the reference itself, the model is rebuilt from model.cfg + trainer.cfg of
    the experiment directory or of its parent (`run test|decode` work in <expdir>/test, <expdir>/decode with a symlink
    to the training run's model directory).  The variables are restored from model/network.ckpt by the caller."""
    from ..neuralnetworks.models.model import Model
    pkl = os.path.join(expdir, 'model', 'model.pkl')
    if os.path.isfile(pkl):
        try:
            return Model.load(pkl)
        except Exception as e:                      # a reference pickle: fall through to the cfg files
            print('ignoring %s: %s' % (pkl, e))
    for d in (expdir, os.path.dirname(os.path.abspath(expdir))):
        if os.path.isfile(os.path.join(d, 'model.cfg')) and os.path.isfile(os.path.join(d, 'trainer.cfg')):
            model_cfg, trainer_cfg = read_cfg(d, 'model.cfg'), read_cfg(d, 'trainer.cfg')
            return Model(conf=model_cfg, trainlabels=int(trainer_cfg.get('trainer', 'trainlabels')), constraint=None)
    raise IOError('%s: neither model/model.pkl nor model.cfg + trainer.cfg found' % expdir)
