"""CPU, world_size 2 over gloo: the N>1 host logic (shape / slab sharding, timing reduction, band gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from points2surf_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        Q = 1001
        first, count = sharding.query_slab(Q, rank, world)
        full = torch.arange(Q, dtype=torch.float32) * 0.5
        slab = full[first:first + count].clone()                      # stand-in for Engine.reconstruct(first, count)
        counts = [sharding.query_slab(Q, r, world)[1] for r in range(world)]
        band = sharding.gather_band(slab, counts)
        band_dst = sharding.gather_band(slab, counts, dst=1)          # only the meshing rank receives
        assert (band_dst is None) if rank != 1 else bool(torch.equal(band_dst, full))
        # LPT: deterministic, identical on every rank, balanced
        bins, tot = sharding.lpt_assign([90, 10, 40, 60, 50, 50], world)
        assert sorted(sum(bins, [])) == list(range(6)) and abs(tot[0] - tot[1]) <= 10
        assert sharding.shapes_for_rank(6, rank, world, loads=[90, 10, 40, 60, 50, 50]) == bins[rank]
        ms, units = sharding.reduce_timing(10.0 + rank, count)
        # final mesh gather: rank r owns shapes r, r+2, ... of 5; mesh i has i+1 vertices and 2i faces (shape 0: no faces)
        mine = [(i, torch.full((i + 1, 3), float(i)), torch.full((2 * i, 3), i, dtype=torch.int32))
                for i in sharding.shapes_for_rank(5, rank, world)]
        got = sharding.gather_meshes(mine, dst=0)
        if rank == 0:
            mesh_ok = [g[0] for g in got] == list(range(5)) and all(
                g[1].shape == (i + 1, 3) and g[2].shape == (2 * i, 3) and bool((g[1] == i).all()) and bool((g[2] == i).all())
                for i, g in enumerate(got))
        else:
            mesh_ok = got == []
        q.put((rank, sharding.shapes_for_rank(7, rank, world), first, count,
               bool(torch.equal(band, full)) and mesh_ok, ms, units))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, f0, c0, ok0, ms0, u0), (r1, s1, f1, c1, ok1, ms1, u1) = res
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5]
    assert (f0, c0) == (0, 501) and (f1, c1) == (501, 500)
    assert ok0 and ok1
    assert ms0 == ms1 == 11.0 and u0 == u1 == 1001.0


def test_slabs_partition_every_size():
    for Q in (0, 1, 5, 152637):
        for world in (1, 2, 3, 8):
            slabs = [sharding.query_slab(Q, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and sum(c for _, c in slabs) == Q
            for (f, c), (f2, _) in zip(slabs, slabs[1:]):
                assert f + c == f2
            assert max(c for _, c in slabs) - min(c for _, c in slabs) <= 1


def _train_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers_train import TorchPrims, make_train_batch
    from points2surf_b200 import synth
    from points2surf_b200.train import TrainStep
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sd = synth.make_state_dict('max', seed=4)
        ts = TrainStep(sd, 0, 0, points_per_patch=16, sub_sample_size=24, device='cpu', prims=TorchPrims(), dtype=torch.float64)
        batch = {k: t.double() for k, t in make_train_batch(4, 16, 24, seed=50 + rank).items()}
        ts.step(batch)
        q.put((rank, ts.flat_params.clone().numpy(), ts.flat_grads.clone().numpy()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_step_gloo():
    """Two ranks, different shards: after the step both hold the same parameters, and the applied gradient is the
    mean of the two per-rank gradients (computed again in this process without a process group)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers_train import TorchPrims, make_train_batch
    from points2surf_b200 import synth
    from points2surf_b200.train import TrainStep
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    local = []
    for rank in range(2):
        ts = TrainStep(synth.make_state_dict('max', seed=4), 0, 0, points_per_patch=16, sub_sample_size=24, device='cpu',
                       prims=TorchPrims(), dtype=torch.float64)
        ts.step({k: t.double() for k, t in make_train_batch(4, 16, 24, seed=50 + rank).items()})
        local.append(ts.flat_grads.clone().numpy())
    np.testing.assert_allclose(res[0][2], 0.5 * (local[0] + local[1]), rtol=1e-12, atol=1e-14)
