"""Read sharding across the GPUs of one box: reads are independent units, the graph is replicated
on every GPU, there is no data-path collective.  torch.distributed is used only for the rendezvous, the barrier
around the timed region, a MAX over ranks of the device time and the final gather of fixed-size result records."""
import numpy as np


def shard_bounds(n_items, world):
    """Contiguous, balanced [lo, hi) ranges by index (reads sharded contiguously, as config 4 asks)."""
    base, extra = divmod(n_items, world)
    bounds, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def reduce_max(dist, value, device="cpu"):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, values, device="cpu"):
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


RECORD = np.dtype([("read", "<i8"), ("failed", "<i4"), ("score", "<i4"), ("start", "<u8"), ("end", "<u8"), ("trace_hash", "<u8")])


def gather_records(dist, local_records, n_total, rank, world, device="cpu"):
    """Final gather of per-read result records (global read index + summary) onto rank 0, returned in read order."""
    import torch
    if dist is None or world == 1:
        out = np.zeros(n_total, dtype=RECORD)
        out[local_records["read"]] = local_records
        return out
    counts = [hi - lo for lo, hi in shard_bounds(n_total, world)]
    width = RECORD.itemsize
    mine = torch.from_numpy(np.frombuffer(local_records.tobytes(), dtype=np.uint8).copy()).to(device)
    maxlen = max(counts) * width
    padded = torch.zeros(maxlen, dtype=torch.uint8, device=device)
    padded[:mine.numel()] = mine
    gathered = [torch.zeros(maxlen, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(gathered, padded)
    if rank != 0:
        return None
    out = np.zeros(n_total, dtype=RECORD)
    for r in range(world):
        rec = np.frombuffer(gathered[r].cpu().numpy().tobytes()[:counts[r] * width], dtype=RECORD)
        out[rec["read"]] = rec
    return out
