"""Stream sharding across GPUs: streams are fully independent, so rank r of W owns the
contiguous range [r*S/W, (r+1)*S/W) and runs its own engine; there is NO collective on the data path.
The only exchange is the final host-side gather of decoded message records (a few hundred bytes each),
done here with torch.distributed object collectives (gloo or nccl, whatever the group uses)."""
from __future__ import annotations


def shard_range(total_streams: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the global stream ids owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or total_streams < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_messages(local_msgs, group=None, dst: int = 0):
    """Collect (stream, freq, bbbb, text) records from every rank on `dst`, ordered by (stream, arrival).
    Returns the merged list on dst and None elsewhere.  Works without an initialised process group (W = 1)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return sorted(local_msgs, key=lambda m: m[0])
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    buckets = [None] * world if rank == dst else None
    dist.gather_object(list(local_msgs), buckets, dst=dst, group=group)
    if rank != dst:
        return None
    merged = [m for b in buckets for m in b]
    merged.sort(key=lambda m: m[0])          # stable: per-stream arrival order is preserved
    return merged
