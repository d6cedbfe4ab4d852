"""world_size-2 gloo worker for test_host_logic.test_data_parallel_gradients_world_size_2_gloo."""
import numpy as np
import torch
import torch.distributed as dist

import oracle as O
from nabu_b200.parallel import shard_batch, allreduce_mean_

dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(0)
B, T, D, H, V = 4, 12, 6, 4, 5
x = rng.standard_normal((B, T, D))
lens = np.array([12, 9, 12, 7])
labels = rng.integers(0, V - 1, size=(B, 3))
ll = np.array([3, 2, 3, 1])
layer = O.init_blstm_params(rng, D, H, np.float64)
lin = O.init_linear_params(rng, 2 * H, V, np.float64)


def grads(xs, ls, labs, lls):
    enc, _, caches = O.dblstm_fwd(xs, ls, [layer])
    logits = O.linear_fwd(enc, lin)
    loss, dlog = O.ctc_loss_mean(logits, ls, labs, lls)        # mean over the LOCAL batch
    denc, glin = O.linear_bwd(enc, lin, dlog)
    _, gl = O.dblstm_bwd(caches, denc)
    return loss, np.concatenate([gl[0]['fw_kernel'].ravel(), gl[0]['bw_bias'].ravel(), glin['weights'].ravel()])


full_loss, full = grads(x, lens, labels, ll)
sx, sl, slab, sll = shard_batch((x, lens, labels, ll), rank, world)
assert sx.shape[0] == B // world
loss, g = grads(sx, sl, slab, sll)
flat = torch.from_numpy(g.copy())
allreduce_mean_(flat)                                          # SUM all-reduce, then 1/world
assert np.abs(flat.numpy() - full).max() < 1e-12, np.abs(flat.numpy() - full).max()
lt = torch.tensor([loss])
allreduce_mean_(lt)
assert abs(lt.item() - full_loss) < 1e-12
if rank == 0:
    print('DP_OK')
dist.destroy_process_group()
