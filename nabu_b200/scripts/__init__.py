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
    """The reference unpickles model/model.pkl (scripts/decode.py:46-48); a pickled TF-graph builder has no meaning
    here, so the model is rebuilt from the experiment's own model.cfg + trainer.cfg (what produced that pickle) and its
    variables are restored from model/network.ckpt by the caller."""
    from ..neuralnetworks.models.model import Model
    model_cfg = read_cfg(expdir, 'model.cfg')
    trainer_cfg = read_cfg(expdir, 'trainer.cfg')
    return Model(conf=model_cfg, trainlabels=int(trainer_cfg.get('trainer', 'trainlabels')), constraint=None)
