"""Data-parallel plumbing.  Utterances are independent through encoder, decoder
and loss, so a global minibatch shards over ranks with no data-path exchange; the only collective is
ONE all-reduce of the flat gradient buffer per step (NCCL over NVLink on GPUs, gloo in CPU tests).

With equal shard sizes, mean-over-global-batch = (1/world) * sum over ranks of mean-over-shard, so each
rank keeps the reference's per-batch `reduce_mean` loss and the 1/world factor is folded into the
fused clip+Adam kernel (grad_scale) AFTER the reduction -- clipping then sees the global-batch
gradient, like the reference's non_distributed run at that batch size (trainer.py:556-563)."""
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_batch(arrays, rank, world):
    """Rank r takes utterances r::world of every batch-major array."""
    return tuple(a[rank::world] for a in arrays)


def allreduce_sum_(flat):
    if world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def allreduce_mean_(flat):
    """SUM all-reduce followed by 1/world (the CPU-test twin of all-reduce + grad_scale in the kernel)."""
    allreduce_sum_(flat)
    w = world_size()
    if w > 1:
        flat.mul_(1.0 / w)
    return flat
