"""Frame sharding for the multi-GPU path: frames are independent units, so a batch is cut into
contiguous blocks, one per rank, and the only exchange is an all-gather of fixed-size top-K record buffers."""
import numpy as np


def shard_range(n_frames, world_size, rank):
    """Contiguous block [lo, hi) of frames for `rank`; blocks differ in size by at most one frame."""
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def records_per_rank(boxes_per_frame, world_size, topk):
    """Static message size: the largest per-rank record count (ranks pad with valid == 0 records)."""
    n = len(boxes_per_frame)
    best = 0
    for r in range(world_size):
        lo, hi = shard_range(n, world_size, r)
        best = max(best, int(sum(boxes_per_frame[lo:hi])))
    return best * topk


def pad_records(recs, n_slots):
    """Flatten a rank's (n_boxes x topk) record array into the fixed-size gather slot."""
    flat = np.zeros(n_slots, recs.dtype)
    flat[:recs.size] = recs.reshape(-1)
    return flat


def unpack_gathered(gathered, boxes_per_frame, world_size, topk):
    """Inverse of shard + pad: (world x n_slots) gathered records -> list per frame of (n_boxes x topk) arrays."""
    n = len(boxes_per_frame)
    out = []
    for r in range(world_size):
        lo, hi = shard_range(n, world_size, r)
        off = 0
        for f in range(lo, hi):
            nb = int(boxes_per_frame[f])
            out.append(gathered[r, off:off + nb * topk].reshape(nb, topk))
            off += nb * topk
    return out
