"""Data-parallel training from nabu data directories, checked against the single-GPU run at the same global batch.

    python tools/dp_train_check.py prepare <root>            # two identical experiment directories under <root>
    python -m nabu_b200.scripts.train --expdir <root>/one/exp                                  # 1 GPU
    torchrun --nproc-per-node 2 ... -m nabu_b200.scripts.train --expdir <root>/two/exp         # 2 GPUs
    python tools/dp_train_check.py compare <root>

Every rank walks the same global batches and keeps utterances rank::2 (BatchSource), the gradients are summed by one
all-reduce and scaled by 1/world inside the clip+Adam kernel, so the two runs differ only by the order of fp32
additions: the checkpoints `model/network.ckpt` must agree to a few 1e-5 after the 8 Adam steps
(tools/gpu_dp_check.sh runs the four commands on a 2-GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    cmd, root = sys.argv[1], sys.argv[2]
    if cmd == 'prepare':
        from tests.util import write_experiment
        for name in ('one', 'two'):
            print(write_experiment(os.path.join(root, name), num_epochs=2, variable_batch_size=False))
        return
    from nabu_b200.processing import tfcheckpoint
    a = tfcheckpoint.read_checkpoint(os.path.join(root, 'one', 'exp', 'model', 'network.ckpt'))
    b = tfcheckpoint.read_checkpoint(os.path.join(root, 'two', 'exp', 'model', 'network.ckpt'))
    assert set(a) == set(b)
    worst = max(float(np.abs(a[k] - b[k]).max()) for k in a)
    moved = max(float(np.abs(a[k]).max()) for k in a)
    print('DP_CHECK variables %d  max |theta_1gpu - theta_2gpu| = %.3e  (max |theta| %.3f)' % (len(a), worst, moved))
    assert worst < 2e-4, worst
    print('DP_CHECK_OK')


if __name__ == '__main__':
    main()
