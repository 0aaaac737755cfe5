"""Host-side sharding helpers.  The path shards over independent units -- shapes, or
contiguous slabs of one shape's ordered query list -- so there is no data-path collective; the only
communication is the timing reduction of bench.py and (optionally) gathering per-shape results."""
import torch
import torch.distributed as dist


def shapes_for_rank(num_shapes, rank, world, loads=None):
    """Shape assignment used by points2surf_b200.eval: greedy LPT by `loads` (candidate-query counts) when given,
    round-robin otherwise."""
    if loads is not None:
        return lpt_assign(loads, world)[0][rank]
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


def lpt_assign(loads, world):
    """Greedy longest-processing-time assignment of independent shapes to ranks: shapes sorted by
    decreasing load (the candidate-query count Q of the cheap grid kernel), each to the currently least loaded rank.
    Deterministic (ties by shape index / lowest rank), so every rank computes the same table without communication.
    -> (list of shape-index lists per rank, per-rank load)."""
    order = sorted(range(len(loads)), key=lambda i: (-float(loads[i]), i))
    bins, tot = [[] for _ in range(world)], [0.0] * world
    for i in order:
        r = min(range(world), key=lambda k: (tot[k], k))
        bins[r].append(i)
        tot[r] += float(loads[i])
    return [sorted(b) for b in bins], tot


def _p2p_gather(tensor, counts_rows, dst):
    """Rows of `tensor` from every rank to `dst` by point-to-point transfers (NCCL send/recv over NVLink on GPUs, gloo on
    CPU): only `dst` receives, and only the real bytes move -- no padding, no replication to ranks that do not need it.
    `counts_rows[r]` = rows held by rank r (known to everyone).  Returns the list of per-rank tensors on dst, [] elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == dst:
        out, reqs = [], []
        for r in range(world):
            if r == dst:
                out.append(tensor)
            else:
                buf = torch.empty((counts_rows[r],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
                out.append(buf)
                if counts_rows[r] > 0:
                    reqs.append(dist.irecv(buf, src=r))
        for q in reqs:
            q.wait()
        return out
    if counts_rows[rank] > 0:
        dist.send(tensor.contiguous(), dst=dst)
    return []


def gather_band(sdf_slab, counts, dst=None, device=None):
    """Tile mode: every rank holds the SDF of its slab of one shape's ordered query list.  dst=None: the full band on
    every rank (one padded all_gather).  dst=r: only rank r -- the one that runs sign propagation and marching cubes --
    receives it (point-to-point, Q*4 bytes in total); other ranks get None.  `counts` = slab sizes of all ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return sdf_slab
    world = dist.get_world_size()
    if dst is not None:
        parts = _p2p_gather(sdf_slab, counts, dst)
        return torch.cat(parts) if dist.get_rank() == dst else None
    m = max(counts)
    pad = torch.zeros(m, dtype=sdf_slab.dtype, device=sdf_slab.device)
    pad[:sdf_slab.numel()] = sdf_slab
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)])


def gather_meshes(meshes, dst=0):
    """The "final mesh gather" of the shape-sharded run: every rank holds the meshes of its own
    shapes as a list of (shape_index, verts [V,3] fp32, faces [F,3] int32) tensors on its device; rank `dst` receives
    all of them, ordered by shape index, other ranks receive [].  One small all_gather of the size table (so that dst can
    size its receive buffers), then each rank sends its concatenated vertex and face buffers to dst point-to-point: the
    bytes on the wire are the mesh bytes, once (a padded all_gather would replicate every mesh to every rank)."""
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(meshes, key=lambda m: m[0])
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = meshes[0][1].device if meshes else (torch.device('cuda', torch.cuda.current_device())
                                              if dist.get_backend() == 'nccl' else torch.device('cpu'))
    n_local = torch.tensor([len(meshes)], dtype=torch.int64, device=dev)
    n_all = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(n_all, n_local)
    n_max = max(int(n.item()) for n in n_all)
    table = torch.zeros((max(n_max, 1), 3), dtype=torch.int64, device=dev)
    for i, (idx, v, f) in enumerate(meshes):
        table[i] = torch.tensor([idx, v.shape[0], f.shape[0]], dtype=torch.int64)
    tables = [torch.zeros_like(table) for _ in range(world)]
    dist.all_gather(tables, table)
    tables = [t.cpu() for t in tables]
    n_all = [int(n.item()) for n in n_all]
    tot_v = [int(t[:n, 1].sum()) for t, n in zip(tables, n_all)]
    tot_f = [int(t[:n, 2].sum()) for t, n in zip(tables, n_all)]
    if meshes:
        vcat = torch.cat([m[1].reshape(-1, 3).to(torch.float32) for m in meshes])
        fcat = torch.cat([m[2].reshape(-1, 3).to(torch.int32) for m in meshes])
    else:
        vcat = torch.zeros((0, 3), dtype=torch.float32, device=dev)
        fcat = torch.zeros((0, 3), dtype=torch.int32, device=dev)
    vall = _p2p_gather(vcat, tot_v, dst)
    fall = _p2p_gather(fcat, tot_f, dst)
    if rank != dst:
        return []
    out = []
    for r in range(world):
        vo = fo = 0
        for i in range(n_all[r]):
            idx, nv, nf = (int(x) for x in tables[r][i])
            out.append((idx, vall[r][vo:vo + nv], fall[r][fo:fo + nf]))
            vo += nv
            fo += nf
    return sorted(out, key=lambda m: m[0])
