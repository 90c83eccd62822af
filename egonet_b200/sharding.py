"""Multi-GPU sharding of the crop stream.

Every crop is independent through HC, decode, lifter and pose, so ranks never exchange
activations: rank r of W processes the contiguous block ``partition(n, r, W)`` of the crop stream
with its own engine handles, and the only collective is ONE all-gather of the per-rank pose
records ([n_r, 7] fp64: Euler x,y,z | translation | alpha) -- NCCL over NVLink on GPUs, gloo in
the CPU tests.
"""
import torch


def partition(n, rank, world):
    """Contiguous block partition: rank r gets [r*n//W, (r+1)*n//W)."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world of %d' % (rank, world))
    return (rank * n) // world, ((rank + 1) * n) // world


def counts(n, world):
    return [partition(n, r, world)[1] - partition(n, r, world)[0] for r in range(world)]


def all_gather_records(local, n_total, dist=None):
    """Gather per-rank record blocks ``local`` [n_r, F] into the full [n_total, F] tensor (same on
    every rank, global crop order).  Equal blocks use one ``all_gather_into_tensor``; ragged blocks
    are padded to the largest block so that a single collective still suffices."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        if local.shape[0] != n_total:
            raise ValueError('single-process gather expects all %d records, got %d' % (n_total, local.shape[0]))
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    cs = counts(n_total, world)
    if local.shape[0] != cs[rank]:
        raise ValueError('rank %d holds %d records, partition says %d' % (rank, local.shape[0], cs[rank]))
    width = max(cs)
    feat = local.shape[1:]
    if min(cs) == width:
        out = torch.empty((n_total,) + tuple(feat), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros((width,) + tuple(feat), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * width,) + tuple(feat), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    return torch.cat([buf[r * width:r * width + cs[r]] for r in range(world)], dim=0)
