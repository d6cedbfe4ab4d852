"""`python -m nabu_b200.scripts.train --expdir <dir>` (reference: nabu/scripts/train.py:13-90).

One process per GPU; under `torchrun` the ranks form a synchronous data-parallel job (NCCL) in place of the reference's
parameter servers: rank r trains on utterances r::world of every minibatch, rank 0 is the chief (validation, saving)."""
import argparse
import os

import torch
import torch.distributed as dist

from . import read_cfg
from ..neuralnetworks.trainers import trainer_factory


def train(expdir, testing=False, device=None):
    database_cfg = read_cfg(expdir, 'database.conf', 'database.cfg')
    model_cfg = read_cfg(expdir, 'model.cfg')
    trainer_cfg = read_cfg(expdir, 'trainer.cfg')
    evaluator_cfg = read_cfg(expdir, 'validation_evaluator.cfg')
    task_index = 0
    if device is None:
        local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        device = torch.device('cuda', local_rank)
        torch.cuda.set_device(device)
    own_group = int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized()
    if own_group:
        dist.init_process_group('nccl' if torch.device(device).type == 'cuda' else 'gloo')
    if dist.is_available() and dist.is_initialized():
        task_index = dist.get_rank()
    tr = trainer_factory.factory(trainer_cfg.get('trainer', 'trainer'))(
        conf=trainer_cfg, dataconf=database_cfg, modelconf=model_cfg, evaluatorconf=evaluator_cfg, expdir=expdir,
        server=None, task_index=task_index, device=device)
    print('starting training')
    tr.train(testing)
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    return tr


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--expdir', default='expdir', help='the experiments directory')
    train(ap.parse_args().expdir, False)
