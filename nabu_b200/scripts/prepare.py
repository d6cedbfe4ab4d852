"""`run train|test|decode --recipe=<dir> --expdir=<dir>`: the local legs of nabu's prepare scripts.

reference: nabu/scripts/prepare_train.py:18-123 (copy database.conf, model.cfg, trainer.cfg, validation_evaluator.cfg
into the experiment directory, then train), prepare_test.py:15-75 and prepare_decode.py:27-82 (a sub-directory `test`
/ `decode` holding database.conf, the evaluator / recognizer cfg and a symlink to the training run's `model`, then
test / decode in it).  Of the reference's computing modes only the local ones exist here: `non_distributed` (one
process, one GPU) and `single_machine` (one process per GPU under torch.distributed.run -- synchronous data
parallelism in place of the parameter servers and workers of prepare_train.py:93-123).  Condor, ssh tunnels,
`multi_machine`, `run data` and `run sweep` are outside the hot path."""
import argparse
import os
import shutil
import subprocess
import sys


def _copy(recipe, expdir, names_in, name_out):
    for name in names_in:
        src = os.path.join(recipe, name)
        if os.path.isfile(src):
            shutil.copyfile(src, os.path.join(expdir, name_out))
            return
    raise Exception('cannot find %s in recipe %s' % (' or '.join(names_in), recipe))


def _check(expdir, recipe):
    if expdir is None or recipe is None:
        raise Exception('no expdir or recipe specified. Command usage: '
                        'run <command> --expdir=/path/to/expdir --recipe=/path/to/recipe')
    if not os.path.isdir(recipe):
        raise Exception('cannot find recipe %s' % recipe)


def prepare_train(expdir, recipe, mode='non_distributed', numgpus=None, overwrite=False, run=True):
    """Returns the experiment directory.  An existing directory is resumed from its own cfg files (the reference asks
    resume / overwrite on the terminal; here `overwrite` decides)."""
    _check(expdir, recipe)
    if mode not in ('non_distributed', 'single_machine'):
        raise Exception('unknown or unsupported distributed mode: %s' % mode)
    if os.path.isdir(expdir) and overwrite:
        shutil.rmtree(expdir)
    if not os.path.isdir(expdir):
        os.makedirs(os.path.join(expdir, 'model'))
        _copy(recipe, expdir, ['database.conf', 'database.cfg'], 'database.conf')
        for name in ('model.cfg', 'validation_evaluator.cfg', 'trainer.cfg'):
            _copy(recipe, expdir, [name], name)
    if not run:
        return expdir
    if mode == 'non_distributed':
        from .train import train
        train(expdir)
    else:
        import torch
        n = numgpus or torch.cuda.device_count()
        subprocess.check_call([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
                               '--master-addr', '127.0.0.1', '--master-port', os.environ.get('MASTER_PORT', '29517'),
                               '-m', 'nabu_b200.scripts.train', '--expdir', expdir])
    return expdir


def _prepare_sub(expdir, recipe, sub, cfg):
    _check(expdir, recipe)
    if not os.path.isdir(os.path.join(expdir, 'model')):
        raise Exception('cannot find a trained model in %s' % os.path.join(expdir, 'model'))
    subdir = os.path.join(expdir, sub)
    if os.path.isdir(subdir):
        shutil.rmtree(subdir)
    os.makedirs(subdir)
    _copy(recipe, subdir, ['database.conf', 'database.cfg'], 'database.conf')
    _copy(recipe, subdir, [cfg], cfg)
    os.symlink(os.path.abspath(os.path.join(expdir, 'model')), os.path.join(subdir, 'model'))
    return subdir


def prepare_test(expdir, recipe, run=True):
    subdir = _prepare_sub(expdir, recipe, 'test', 'test_evaluator.cfg')
    if run:
        from .test import test
        test(subdir)
    return subdir


def prepare_decode(expdir, recipe, run=True):
    subdir = _prepare_sub(expdir, recipe, 'decode', 'recognizer.cfg')
    if run:
        from .decode import decode
        decode(subdir)
    return subdir


def main(argv=None):
    ap = argparse.ArgumentParser(prog='run')
    ap.add_argument('command', choices=['train', 'test', 'decode'])
    ap.add_argument('--expdir')
    ap.add_argument('--recipe')
    ap.add_argument('--mode', default='non_distributed', help='non_distributed | single_machine')
    ap.add_argument('--computing', default='standard')
    ap.add_argument('--numgpus', type=int, default=None, help='single_machine: processes to start (default: all GPUs)')
    ap.add_argument('--overwrite', action='store_true')
    args = ap.parse_args(argv)
    if args.computing != 'standard':
        raise Exception('unknown or unsupported computing mode: %s' % args.computing)
    if args.command == 'train':
        prepare_train(args.expdir, args.recipe, args.mode, args.numgpus, args.overwrite)
    elif args.command == 'test':
        prepare_test(args.expdir, args.recipe)
    else:
        prepare_decode(args.expdir, args.recipe)


if __name__ == '__main__':
    main()
