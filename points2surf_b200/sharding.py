"""Host-side sharding helpers.  The path shards over independent units -- shapes, or
contiguous slabs of one shape's ordered query list -- so there is no data-path collective; the only
communication is the timing reduction of bench.py and (optionally) gathering per-shape results."""
import torch
import torch.distributed as dist


def shapes_for_rank(num_shapes, rank, world):
    """Round-robin shape assignment used by points2surf_b200.eval."""
    return [i for i in range(num_shapes) if i % world == rank]


def query_slab(num_queries, rank, world):
    """Contiguous slab [first, first+count) of the ordered query list for tile-level sharding of one shape;
    feed it to Engine.reconstruct(first_query=..., num_queries=...).  Slabs differ by at most one query."""
    base, rem = divmod(num_queries, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def reduce_timing(ms_local, units_local, device=None):
    """-> (max over ranks of the elapsed ms, sum over ranks of the processed units).  Works on the default
    process group (NCCL on GPUs, gloo on CPU); identity when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(ms_local), float(units_local)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def gather_band(sdf_slab, counts, device=None):
    """Tile mode: every rank holds the SDF of its slab; returns the full band on every rank (all_gather of
    padded slabs).  `counts` = slab sizes of all ranks (from query_slab)."""
    if not (dist.is_available() and dist.is_initialized()):
        return sdf_slab
    world = dist.get_world_size()
    m = max(counts)
    pad = torch.zeros(m, dtype=sdf_slab.dtype, device=sdf_slab.device)
    pad[:sdf_slab.numel()] = sdf_slab
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)])
