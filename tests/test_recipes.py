"""The reference's only check, mirrored: `nabu/scripts/test_recipes.py:7-38` walks config/recipes and, per recipe, runs
train / decode / test with `testing=True`, i.e. builds every graph and returns (trainers/trainer.py:610-611).  Here:
every shipped recipe whose model is on the hot path (LAS/TIMIT, LAS/GP, DBLSTM/TIMIT) is read UNCHANGED from the
reference tree -- model.cfg, trainer.cfg, validation_evaluator.cfg, recognizer.cfg -- and the trainer (model variables
declared from the data's dimensions), the evaluator and the recognizer with its decoder are constructed from them over
data directories in nabu's on-disk format.  The recipe files are not copied into this repo; the test is skipped where
the reference tree is absent (the GPU box)."""
import configparser
import os

import numpy as np
import pytest

from tests.test_processing import _write_stream

RECIPES = '/root/reference/config/recipes'
pytestmark = pytest.mark.skipif(not os.path.isdir(RECIPES), reason='reference tree not present')


def _read(path):
    conf = configparser.ConfigParser()
    conf.read(path)
    return conf


def _database(tmp_path, sections, alphabet, dim=40, recipe=None):
    """one tiny data directory per database section a cfg of the recipe names; the data type of a section is the one
    the recipe's own database.cfg gives it (e.g. `string` for the CTC targets of DBLSTM/TIMIT), when it ships one"""
    rng = np.random.default_rng(0)
    lines = []
    shipped = _read(os.path.join(recipe, 'database.cfg')) if recipe else None
    for sec, kind in sorted(sections.items()):
        if shipped is not None and shipped.has_section(sec):
            kind = shipped.get(sec, 'type')
        d = str(tmp_path / sec)
        lens = [9, 14, 11, 12]
        if kind == 'audio_feature':
            _write_stream(d, 'audio', [('u%d' % i, rng.standard_normal((L, dim)).astype(np.float32))
                                       for i, L in enumerate(lens)], dim=dim)
        else:
            _write_stream(d, 'text', [('u%d' % i, ' '.join(rng.choice(alphabet, size=3))) for i in range(len(lens))],
                          alphabet=alphabet)
        lines.append('[%s]\ndir = %s\ntype = %s\n' % (sec, d, kind))
    conf = configparser.ConfigParser()
    conf.read_string(''.join(lines))
    return conf


def _sections(conf, section, names, kind, out):
    for name in names:
        for sec in conf.get(section, name).split(' '):
            out[sec] = kind


@pytest.mark.parametrize('recipe', ['LAS/TIMIT', 'LAS/GP', 'DBLSTM/TIMIT'])
def test_recipe_builds(recipe, tmp_path):
    from nabu_b200.neuralnetworks.evaluators import evaluator_factory
    from nabu_b200.neuralnetworks.recognizer import Recognizer
    from nabu_b200.neuralnetworks.trainers import trainer_factory
    rdir = os.path.join(RECIPES, recipe)
    mconf, tconf = _read(os.path.join(rdir, 'model.cfg')), _read(os.path.join(rdir, 'trainer.cfg'))
    econf, rconf = _read(os.path.join(rdir, 'validation_evaluator.cfg')), _read(os.path.join(rdir, 'recognizer.cfg'))
    inputs = mconf.get('io', 'inputs').split(' ')
    outputs = mconf.get('io', 'outputs').split(' ')
    dims = [int(d) for d in mconf.get('io', 'output_dims').split(' ')]
    alphabet = ['s%d' % i for i in range(min(dims))]
    sections = {}
    _sections(tconf, 'trainer', inputs, 'audio_feature', sections)
    _sections(tconf, 'trainer', tconf.get('trainer', 'targets').split(' '), 'string_eos', sections)
    _sections(econf, 'evaluator', inputs, 'audio_feature', sections)
    _sections(econf, 'evaluator', econf.get('evaluator', 'targets').split(' '), 'string_eos', sections)
    _sections(rconf, 'recognizer', inputs, 'audio_feature', sections)
    dataconf = _database(tmp_path, sections, alphabet, recipe=rdir)
    tconf.set('trainer', 'batch_size', '2')          # four utterances per section here
    trainer = trainer_factory.factory(tconf.get('trainer', 'trainer'))(
        tconf, dataconf, mconf, econf, str(tmp_path / 'exp'), None, 0, device='cpu')
    trainer.train(testing=True)
    store = trainer.model.store
    assert store.materialised and trainer.num_steps == len(trainer.batch_source) * int(trainer.conf['num_epochs'])
    assert set(trainer.model.output_dims) == set(outputs)
    scope = {'listener': 'Listener', 'dblstm': 'DBLSTM'}[mconf.get('encoder', 'encoder')]
    assert any(v.name.startswith(scope + '/' + inputs[0] + '/layer0/') for v in store.order)
    for out, d in zip(outputs, dims):
        assert trainer.model.output_dims[out] == d + int(trainer.conf['trainlabels'])
    evaluator = evaluator_factory.factory(econf.get('evaluator', 'evaluator'))(econf, dataconf, trainer.model)
    assert evaluator.target_names == econf.get('evaluator', 'targets').split(' ')
    recognizer = Recognizer(trainer.model, rconf, dataconf, str(tmp_path / 'exp'))
    assert type(recognizer.decoder).__name__.lower().replace('_', '') == \
        rconf.get('decoder', 'decoder').replace('_', '')


def test_recipe_outside_the_hot_path_says_so():
    from nabu_b200.neuralnetworks.models.model import Model
    with pytest.raises(Exception, match='(?i)dnn|hot path|scope|unknown|undefined'):
        Model(_read(os.path.join(RECIPES, 'DNN/WSJ/model.cfg')), 0).build({'features': 40}, 'cpu')


def test_scripts_build_from_an_experiment_directory(tmp_path):
    """`run train|decode|test` copy the recipe into the experiment directory and call scripts/{train,decode,test}.py
    on it; with testing=True these build everything and return (scripts/test_recipe.py:21-33)."""
    import shutil
    from nabu_b200.scripts import decode, test, train
    rdir = os.path.join(RECIPES, 'DBLSTM/TIMIT')
    expdir = str(tmp_path / 'exp')
    os.makedirs(expdir)
    for name in ('model.cfg', 'trainer.cfg', 'validation_evaluator.cfg', 'test_evaluator.cfg', 'recognizer.cfg'):
        shutil.copy(os.path.join(rdir, name), expdir)
    tconf, econf = _read(os.path.join(rdir, 'trainer.cfg')), _read(os.path.join(rdir, 'validation_evaluator.cfg'))
    xconf, rconf = _read(os.path.join(rdir, 'test_evaluator.cfg')), _read(os.path.join(rdir, 'recognizer.cfg'))
    sections = {}
    for conf, sec in ((tconf, 'trainer'), (econf, 'evaluator'), (xconf, 'evaluator')):
        _sections(conf, sec, ['features'], 'audio_feature', sections)
        _sections(conf, sec, conf.get(sec, 'targets').split(' '), 'string_eos', sections)
    _sections(rconf, 'recognizer', ['features'], 'audio_feature', sections)
    dims = int(_read(os.path.join(rdir, 'model.cfg')).get('io', 'output_dims'))
    dataconf = _database(tmp_path, sections, ['s%d' % i for i in range(dims)], recipe=rdir)
    assert dataconf.get('traintext', 'type') == 'string'
    with open(os.path.join(expdir, 'database.conf'), 'w') as fid:
        dataconf.write(fid)
    tconf.set('trainer', 'batch_size', '2')          # four utterances per section here
    with open(os.path.join(expdir, 'trainer.cfg'), 'w') as fid:
        tconf.write(fid)
    tr = train.train(expdir, testing=True, device='cpu')
    assert tr.model.store.materialised and tr.num_steps > 0
    rec = decode.decode(expdir, testing=True, device='cpu')
    assert type(rec.decoder).__name__ == 'CTCDecoder'
    ev = test.test(expdir, testing=True, device='cpu')
    assert ev.target_names == ['text']


def test_run_commands_prepare_the_experiment_directory(tmp_path):
    """prepare_train / prepare_test / prepare_decode lay the directories out as the reference does (cfg copies,
    <expdir>/test and <expdir>/decode with a symlink to the training run's model) and the entry points build from
    them; model/model.pkl carries the model description from train to test / decode."""
    from nabu_b200.scripts import decode, load_model, prepare, test, train
    rdir = os.path.join(RECIPES, 'LAS/TIMIT')
    recipe = str(tmp_path / 'recipe')
    os.makedirs(recipe)
    import shutil
    for name in os.listdir(rdir):
        shutil.copy(os.path.join(rdir, name), recipe)
    confs = {n: _read(os.path.join(rdir, n)) for n in ('trainer.cfg', 'validation_evaluator.cfg', 'test_evaluator.cfg',
                                                        'recognizer.cfg')}
    sections = {}
    for name, sec in (('trainer.cfg', 'trainer'), ('validation_evaluator.cfg', 'evaluator'),
                      ('test_evaluator.cfg', 'evaluator')):
        _sections(confs[name], sec, ['features'], 'audio_feature', sections)
        _sections(confs[name], sec, confs[name].get(sec, 'targets').split(' '), 'string_eos', sections)
    _sections(confs['recognizer.cfg'], 'recognizer', ['features'], 'audio_feature', sections)
    dataconf = _database(tmp_path, sections, ['s%d' % i for i in range(39)])
    with open(os.path.join(recipe, 'database.conf'), 'w') as fid:
        dataconf.write(fid)
    confs['trainer.cfg'].set('trainer', 'batch_size', '2')
    with open(os.path.join(recipe, 'trainer.cfg'), 'w') as fid:
        confs['trainer.cfg'].write(fid)
    expdir = str(tmp_path / 'exp')
    assert prepare.prepare_train(expdir, recipe, run=False) == expdir
    assert sorted(os.listdir(expdir)) == ['database.conf', 'model', 'model.cfg', 'trainer.cfg',
                                          'validation_evaluator.cfg']
    tr = train.train(expdir, testing=True, device='cpu')
    tr.model.save(os.path.join(expdir, 'model', 'model.pkl'))               # what Trainer.train does at its end
    tdir = prepare.prepare_test(expdir, recipe, run=False)
    ddir = prepare.prepare_decode(expdir, recipe, run=False)
    assert sorted(os.listdir(tdir)) == ['database.conf', 'model', 'test_evaluator.cfg']
    assert sorted(os.listdir(ddir)) == ['database.conf', 'model', 'recognizer.cfg']
    assert os.path.islink(os.path.join(ddir, 'model'))
    model = load_model(ddir).build({'features': 40}, 'cpu')
    assert [(v.name, v.shape) for v in model.store.order] == [(v.name, v.shape) for v in tr.model.store.order]
    assert type(decode.decode(ddir, testing=True, device='cpu').decoder).__name__ == 'BeamSearchDecoder'
    assert test.test(tdir, testing=True, device='cpu').target_names == ['text']
    with pytest.raises(Exception, match='multi_machine'):
        prepare.prepare_train(str(tmp_path / 'x'), recipe, mode='multi_machine', run=False)
    with pytest.raises(Exception, match='trained model'):
        prepare.prepare_test(str(tmp_path / 'nothing'), recipe, run=False)
