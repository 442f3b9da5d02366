"""Data-parallel host logic: batch sharding, replica initialisation and the gradient exchange.

The reference has no distributed code; this is the north-star's "shard the batch,
all-reduce the gradients only".  Every reference loss is SUM-reduced (vessel_analysis/01_train/
train.py:41,46,49,58), so the sum of per-shard gradients equals the gradient of the same loss over
the concatenated batch -- modulo per-shard BatchNorm statistics and the per-shard pos_weight, which
the north-star ("gradients only") leaves per shard.  latent_translator's loss is MEAN-reduced
(latent_translator/engine.py:25-26): its gradients are averaged (SUM then 1/world).

These helpers are device-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's contiguous slice of n samples (remainder spread over the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world):
    lo, hi = shard_bounds(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


def broadcast_module(module, src=0, group=None):
    """identical replicas: parameters and buffers (BN running statistics, counters) from `src`."""
    for p in module.parameters():
        dist.broadcast(p.data, src, group=group)
    for b in module.buffers():
        dist.broadcast(b, src, group=group)


def allreduce_gradients(flat_grad, group=None, mean=False):
    """ONE collective over the flat fp32 gradient vector (56 MB for CausalViTVAE @256^2)."""
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    if mean:
        flat_grad.div_(dist.get_world_size(group))
    return flat_grad
