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


_COMM = {'ready': False}


def init_library_comm():
    """One NCCL communicator owned by libnabu_b200.so (include/nabu_b200.h: nabu_comm_*): rank 0 draws the unique id,
    torch.distributed carries its 128 bytes to the other ranks.  CUDA + NCCL jobs only; NABU_LIB_ALLREDUCE=0 keeps
    torch's own all_reduce (the same ncclAllReduce on torch's communicator)."""
    import ctypes
    import os
    import torch
    from . import lib as L
    if _COMM['ready'] or world_size() == 1 or os.environ.get('NABU_LIB_ALLREDUCE', '1') == '0':
        return _COMM['ready']
    if dist.get_backend() != 'nccl' or not torch.cuda.is_available():
        return False
    lib = L.load()
    buf = ctypes.create_string_buffer(128)
    if dist.get_rank() == 0:
        L.check(lib.nabu_comm_unique_id(buf), 'nabu_comm_unique_id')
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device='cuda')
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    L.check(lib.nabu_comm_init(ctypes.create_string_buffer(raw, 128), dist.get_rank(), dist.get_world_size()),
            'nabu_comm_init')
    _COMM['ready'] = True
    return True


def allreduce_sum_(flat):
    """the step's only collective: nabu_allreduce_grads (ncclAllReduce on the caller's stream) on CUDA jobs, torch's
    all_reduce under gloo (the CPU tests)"""
    if world_size() > 1:
        if flat.is_cuda and init_library_comm():
            from . import lib as L
            L.check(L.load().nabu_allreduce_grads(L.ptr(flat), flat.numel(), L.stream()), 'nabu_allreduce_grads')
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def allreduce_mean_(flat):
    """SUM all-reduce followed by 1/world (the CPU-test twin of all-reduce + grad_scale in the kernel)."""
    allreduce_sum_(flat)
    w = world_size()
    if w > 1:
        flat.mul_(1.0 / w)
    return flat
