"""Frame sharding for multi-GPU decode.

Frames are independent given (weights, t, [embedding]), so decode shards embarrassingly: weights are
replicated, rank r decodes frames {i : i mod world == r}, and there is NO collective on the data path.
This mirrors the reference's DistributedSampler sharding (train_nerv_all.py:176) and its metric
averaging by all_reduce (hnerv_utils.py:213-229, train_nerv_all.py:554-556), which is the only
communication: one tiny SUM all-reduce of (sum_psnr, n_frames) at the end.
"""
import torch
import torch.distributed as dist


def frame_indices(n_frames, rank, world):
    """Round-robin shard of range(n_frames); the union over ranks is exactly range(n_frames)."""
    return list(range(rank, n_frames, world))


def norm_index(i, n_frames):
    """norm_idx of frame i as the dataset produces it — hnerv_utils.py:47 ((idx+1)/N), float64 after collate."""
    return float(i + 1) / n_frames


def reduce_metric(sum_value, count, device=None):
    """Mean of a per-frame metric over all ranks: all_reduce(SUM) of (sum, count).  Works on any backend."""
    t = torch.tensor([float(sum_value), float(count)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return (t[0] / t[1].clamp_min(1)).item(), int(t[1].item())
