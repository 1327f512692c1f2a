"""Multi-GPU plumbing for the hot path (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Rows (sampled completions) are independent: each rank owns a contiguous block of rows, weights are broadcast
ONCE from rank 0 as one flat tensor, and fixed-size per-row outputs are all-gathered once per batch.  No collective runs
inside the AR loop.  The reference's equivalent is rank-strided dataset indices with per-rank output files
(xgutils/plutil.py:123-139,187-189)."""
import torch
import torch.distributed as dist


def row_block(total_rows, rank, world, group=1):
    """Contiguous rows [lo, hi) of `rank`; blocks are multiples of `group` (all sample_n rows of a shape stay together)."""
    groups = total_rows // group
    if groups * group != total_rows:
        raise ValueError("total_rows must be a multiple of the group size")
    per, extra = divmod(groups, world)
    lo = rank * per + min(rank, extra)
    hi = lo + per + (1 if rank < extra else 0)
    return lo * group, hi * group


def broadcast_parameters(tensors, src=0):
    """One broadcast for a list of same-dtype tensors (flattened, sent, scattered back in place)."""
    tensors = list(tensors)
    if not dist.is_initialized() or dist.get_world_size() == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.broadcast(flat, src)
    o = 0
    for t in tensors:
        t.copy_(flat[o:o + t.numel()].view_as(t))
        o += t.numel()


def gather_rows(local, out_list=None):
    """All-gather per-rank row tensors (dim 0 = rows; row counts may differ between ranks, as row_block yields unequal blocks
    when the group count is not a multiple of the world size); returns the list ordered by rank, each trimmed to its rank's
    row count.  `out_list` (same-shaped preallocated buffers) is used only when every rank holds the same number of rows."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c) for c in counts]
    top = max(counts)
    if top == min(counts):
        if out_list is None:
            out_list = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(out_list, local.contiguous())
        return out_list
    pad = local.new_zeros((top,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return [b[:c] for b, c in zip(bufs, counts)]
