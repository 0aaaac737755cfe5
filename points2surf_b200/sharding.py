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


def gather_meshes(meshes, dst=0):
    """The "final mesh gather" of the shape-sharded run: every rank holds the meshes of its own
    shapes as a list of (shape_index, verts [V,3] fp32, faces [F,3] int32) tensors on its device; rank `dst` receives
    all of them, ordered by shape index, other ranks receive [].  One all_gather of the per-rank size table, then one
    padded all_gather per buffer kind (NCCL over NVLink on GPUs, gloo on CPU) -- meshes are O(res^2) small, so a single
    padded exchange beats per-mesh point-to-point sends."""
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(meshes, key=lambda m: m[0])
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = meshes[0][1].device if meshes else (torch.device('cuda', torch.cuda.current_device())
                                              if dist.get_backend() == 'nccl' else torch.device('cpu'))
    # size table: [n_meshes, then (shape_index, V, F) per mesh], padded to the largest mesh count
    n_local = torch.tensor([len(meshes)], dtype=torch.int64, device=dev)
    n_all = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(n_all, n_local)
    n_max = max(int(n.item()) for n in n_all)
    table = torch.zeros((max(n_max, 1), 3), dtype=torch.int64, device=dev)
    for i, (idx, v, f) in enumerate(meshes):
        table[i] = torch.tensor([idx, v.shape[0], f.shape[0]], dtype=torch.int64)
    tables = [torch.zeros_like(table) for _ in range(world)]
    dist.all_gather(tables, table)
    tot_v = [int(t[:int(n.item()), 1].sum()) for t, n in zip(tables, n_all)]
    tot_f = [int(t[:int(n.item()), 2].sum()) for t, n in zip(tables, n_all)]
    vbuf = torch.zeros((max(max(tot_v), 1), 3), dtype=torch.float32, device=dev)
    fbuf = torch.zeros((max(max(tot_f), 1), 3), dtype=torch.int32, device=dev)
    if meshes:
        vcat = torch.cat([m[1].reshape(-1, 3).to(torch.float32) for m in meshes])
        fcat = torch.cat([m[2].reshape(-1, 3).to(torch.int32) for m in meshes])
        vbuf[:vcat.shape[0]] = vcat
        fbuf[:fcat.shape[0]] = fcat
    vall = [torch.empty_like(vbuf) for _ in range(world)]
    fall = [torch.empty_like(fbuf) for _ in range(world)]
    dist.all_gather(vall, vbuf)
    dist.all_gather(fall, fbuf)
    if rank != dst:
        return []
    out = []
    for r in range(world):
        vo = fo = 0
        for i in range(int(n_all[r].item())):
            idx, nv, nf = (int(x) for x in tables[r][i])
            out.append((idx, vall[r][vo:vo + nv].clone(), fall[r][fo:fo + nf].clone()))
            vo += nv
            fo += nf
    return sorted(out, key=lambda m: m[0])
