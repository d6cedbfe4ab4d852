"""`python -m nabu_b200.scripts.test --expdir <dir>` (reference: nabu/scripts/test.py:14-110): evaluate the trained
model with the evaluator of test_evaluator.cfg, print the loss and write it to <expdir>/result."""
import argparse
import os

from . import load_model, read_cfg
from ..neuralnetworks.evaluators import evaluator_factory


def test(expdir, testing=False, device='cuda'):
    database_cfg = read_cfg(expdir, 'database.conf', 'database.cfg')
    evaluator_cfg = read_cfg(expdir, 'test_evaluator.cfg')
    model = load_model(expdir)
    model.device = device
    evaluator = evaluator_factory.factory(evaluator_cfg.get('evaluator', 'evaluator'))(
        conf=evaluator_cfg, dataconf=database_cfg, model=model)
    if testing:
        return evaluator
    # LoadAtBegin(model/network.ckpt, model.variables) (scripts/test.py:78-80)
    source = evaluator.source()
    model.build(source.input_dims, device)
    model.store.load_tf_checkpoint(os.path.join(expdir, 'model', 'network.ckpt'))
    loss, _ = evaluator.evaluate()
    loss = float(loss)
    print('loss = %f' % loss)
    with open(os.path.join(expdir, 'result'), 'w') as fid:
        fid.write(str(loss))
    return loss


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--expdir', default='expdir', help='the experiments directory that was used for training')
    test(ap.parse_args().expdir, False)
