"""Query sharding over the GPUs of one box — host plumbing only.

The path shards by independent units (characters, sweeps, rays): every rank holds a replica of the mesh + BVH and
processes one contiguous range of the units.  Nothing is exchanged inside the algorithm; the one collective the path
ever needs is the gather of the fixed-size result records when a caller wants them in one place (NCCL over NVLink on
the GPUs, gloo in the CPU tests).  Records travel as raw bytes (`torch.uint8`), so the gathered buffer is byte for byte
what a single rank would have produced for the whole batch.
"""
import numpy as np


def rank_range(n_units, rank, world_size):
    """Contiguous range [lo, hi) of rank `rank`: lo = rank*n/ws, hi = (rank+1)*n/ws (integer division), so ranges tile
    [0, n) exactly, differ in size by at most one unit and may be empty when n < world_size."""
    if world_size <= 0 or not 0 <= rank < world_size or n_units < 0:
        raise ValueError("rank_range: need 0 <= rank < world_size and n_units >= 0")
    return rank * n_units // world_size, (rank + 1) * n_units // world_size


def shard_sizes(n_units, world_size):
    return [rank_range(n_units, r, world_size)[1] - rank_range(n_units, r, world_size)[0] for r in range(world_size)]


def gather_records(local, n_units, record_bytes, group=None, out=None):
    """All-gather the result records of a sharded batch.

    local: 1-D torch.uint8 tensor (device tensor under NCCL, CPU tensor under gloo) holding this rank's
    `rank_range(n_units, rank, ws)` records of `record_bytes` bytes each.  Returns a uint8 tensor of
    n_units*record_bytes bytes in unit order on every rank.  Equal shards (n_units % ws == 0: every configuration of
    BASELINE.json) go through ONE `all_gather_into_tensor` straight into the result; ragged shards are padded to the
    largest shard and trimmed after the collective.  Without an initialised process group it returns `local`."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        if local.numel() != n_units * record_bytes:
            raise ValueError("gather_records: local shard does not hold the whole batch and no process group exists")
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_units, ws)
    if local.dtype != torch.uint8 or local.dim() != 1 or local.numel() != sizes[rank] * record_bytes:
        raise ValueError("gather_records: local must be a flat uint8 tensor of %d records x %d bytes"
                         % (sizes[rank], record_bytes))
    total = n_units * record_bytes
    if out is None:
        out = torch.empty(total, dtype=torch.uint8, device=local.device)
    elif out.numel() != total or out.dtype != torch.uint8:
        raise ValueError("gather_records: out must be a flat uint8 tensor of n_units*record_bytes bytes")
    if min(sizes) == max(sizes):
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    widest = max(sizes) * record_bytes
    padded = torch.zeros(widest, dtype=torch.uint8, device=local.device)
    padded[:local.numel()] = local
    staged = torch.empty(ws * widest, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(staged, padded, group=group)
    at = 0
    for r in range(ws):
        nb = sizes[r] * record_bytes
        out[at:at + nb] = staged[r * widest:r * widest + nb]
        at += nb
    return out


def records_from_bytes(buf, dtype):
    """View a gathered uint8 tensor (moved to the host if needed) as a numpy record array of `dtype`."""
    return np.frombuffer(buf.cpu().numpy().tobytes(), dtype=dtype)
