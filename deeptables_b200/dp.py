"""Data-parallel gradient exchange (reference: tf.distribute.MirroredStrategy around the model
build, deepmodel.py:88-103 -- gradient all-reduce inside TensorFlow, per-replica BatchNorm).

One process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box; gloo in the CPU tests).
The path shards by batch rows only, so the exchange is:

  1. the loss gradient is pre-scaled by 1/world_size (``scale_for_mean``), which turns the SUM
     all-reduce into MirroredStrategy's global-batch mean;
  2. ONE all-reduce bucket for every dense weight gradient (the flat gradient buffer);
  3. ONE all-reduce for the embedding-table gradient (dense [sum V, D] buffer, zero outside the
     rows a rank touched);
  4. an all-gather of the categorical ids, so that every rank's row-wise Adam visits the UNION of
     touched rows and the replicas stay bit-identical.

Nothing else crosses GPUs (no embedding sharding, no all-to-all); scoring needs no collective.
"""
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if is_distributed() else 1


def scale_for_mean(dz):
    """dLoss/dz of the local batch mean -> contribution to the global-batch mean."""
    if is_distributed():
        dz.mul_(1.0 / dist.get_world_size())
    return dz


def exchange(flat_grad, table_grad, cat_ids):
    """All-reduce the gradients in place and return the union of the ranks' id batches
    ([world*B, F]; ``cat_ids`` itself when single-process or when there is no table)."""
    if not is_distributed():
        return cat_ids
    dist.all_reduce(flat_grad)
    if table_grad is None or cat_ids is None:
        return cat_ids
    gathered = [torch.empty_like(cat_ids) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, cat_ids.contiguous())
    dist.all_reduce(table_grad)
    return torch.cat(gathered, dim=0)


def broadcast_parameters(tensors, src=0):
    """Identical replicas at start: rank ``src``'s initial weights win."""
    if is_distributed():
        for t in tensors:
            if t is not None:
                dist.broadcast(t, src)
