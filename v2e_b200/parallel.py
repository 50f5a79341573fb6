"""Multi-GPU plumbing for the two hot paths. One process per GPU (torchrun), NCCL on
the GPU box, gloo in the CPU tests.

The paths shard without a data-path collective:
  * independent clips (BASELINE config 4): clips are dealt round-robin to ranks, every rank runs the
    whole SloMo + pixel-model path on its clips;
  * the only exchange is the merge of the packed event streams at the end (`gather_event_streams`).
Nothing here touches model arithmetic.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_clips(n_clips, rank, world):
    """Indices of the clips rank `rank` owns: round-robin so that clips of similar cost spread evenly."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_clips, world))


def row_band(height, rank, world, align=1):
    """[y0, y1) of the pixel rows rank `rank` owns when ONE clip's pixel model is sharded over ranks
    (BASELINE config 5). Bands differ by at most `align` rows; empty bands are allowed."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    units = (height + align - 1) // align
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return min(u0 * align, height), min(u1 * align, height)


def gather_event_streams(rows, clip_ids=None, dst=0, group=None):
    """Gathers every rank's packed event rows ([N_r, 4] float32: t, x, y, p) on rank `dst`.

    Row counts differ per rank, so counts are all-gathered first and rows are padded to the largest
    count for the (fixed-size) gather. Returns on `dst` a list with one [N_r, 4] tensor per rank (in rank
    order, on the input's device), elsewhere None. Works with NCCL (CUDA tensors) and gloo (CPU)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rows.dim() != 2 or rows.shape[1] != 4:
        raise ValueError("rows must be [N, 4]")
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx, 4), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=dst, group=group)
        return [b[:c] for b, c in zip(bufs, counts)]
    dist.gather(pad, None, dst=dst, group=group)
    return None


def merge_by_time(streams):
    """Merges per-rank streams of ONE clip (pixel-sharded) into a single stream with non-decreasing
    timestamps (stable: ties keep rank order, like concatenating the bands of each timestamp group)."""
    if not streams:
        return torch.zeros((0, 4), dtype=torch.float32)
    allr = torch.cat(streams, 0)
    order = torch.argsort(allr[:, 0], stable=True)
    return allr[order]


def allreduce_max_int(value, device, group=None):
    """max over ranks of a small integer (frame-global max_n when one clip is pixel-sharded,
    emulator.py:773-775)."""
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())
