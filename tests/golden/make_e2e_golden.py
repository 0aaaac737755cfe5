"""Generate the end-to-end fixtures (tests/golden/e2e_*.npz, points2surf_b200/fitted_fc4.npz).

Run in the build container only (uses /root/reference for config 1):
    python tests/golden/make_e2e_golden.py [fit] [e2e64] [e2e128] [config1] [states]

1. `fit`: a rand-init checkpoint gives a shape-unaware SDF, so a reconstructed "mesh" is noise.  To make the
   end-to-end tests (and the bench's mesh stage) reconstruct a surface, the LAST layer (fc4: 2 x 128 + 2) of the
   synthetic checkpoints is fitted by ridge regression on the fp32 oracle's fc3 activations of ~1000 queries around a
   10k-point sphere (targets: the analytic signed distance, reference conventions: + inside, |d| = tanh(l0)^2 r).
   Everything in front of fc4 stays rand-init.  Written to points2surf_b200/fitted_fc4.npz (a checkpoint piece).
2. `e2e64` / `e2e128`: the reference PATH (oracle: cKDTree patches, RandomState sub-sample stream consumed
   sequentially like source/data_loader.py:272-277,358-362, fp32 CPU network) over EVERY query of one shape;
   the fixture keeps logits + radius per query and a checksum of the sub-sample ids (the ids themselves are
   regenerated from the seed at test time: 48 MB otherwise).
3. `config1`: BASELINE config 1 literally -- abc_minimal test shape, res 32, eps 3, seed 40938661, through the
   UNMODIFIED reference dataset + model (shimmed imports), asserted equal to the oracle in the same run.
"""
import os
import sys
import tempfile
import time
import warnings

import numpy as np

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from oracle import p2s_oracle as orc  # noqa: E402
from points2surf_b200 import synth  # noqa: E402
sys.path.insert(0, os.path.dirname(HERE))
from helpers import ids_checksum  # noqa: E402

torch.set_grad_enabled(False)
FIT_FILE = os.path.join(ROOT, 'points2surf_b200', 'fitted_fc4.npz')
MODEL_SEEDS = {'vanilla': 6, 'max': 4}


def oracle_shape(sd, variant, cloud, qpts, seed, chunk=256, log=None):
    """Reference path over every query of a shape: sequential RandomState stream, fp32 CPU network."""
    v = synth.VARIANTS[variant]
    kd = orc.make_kdtree(cloud)
    rng = np.random.RandomState(seed)
    Q = len(qpts)
    logits = np.empty((Q, 2), np.float32)
    radius = np.empty((Q,), np.float32)
    ids = np.empty((Q, 1000), np.int32)
    t0 = time.time()
    for b in range(0, Q, chunk):
        items = [orc.assemble_query(cloud, kd, qpts[i], 300, 1000, rng, bool(v['uniform_subsample'])) for i in range(b, min(Q, b + chunk))]
        patch = np.stack([it['patch_pts_ps'] for it in items])
        sub = np.stack([it['pts_sub_sample_ms'] for it in items])
        radius[b:b + len(items)] = [it['patch_radius_ms'] for it in items]
        ids[b:b + len(items)] = np.stack([it['sub_sample_ids'] for it in items])
        logits[b:b + len(items)] = orc.model_forward(sd, patch, sub, qpts[b:b + len(items)], v['use_point_stn'], v['shared_transformer'])
        if log and (b // chunk) % 8 == 0:
            print('  %s: %d / %d queries, %.0f s' % (log, b, Q, time.time() - t0), flush=True)
    return logits, radius, ids


def fit():
    out = {}
    cloud = synth.make_cloud('sphere', 10000, seed=0)
    kd = orc.make_kdtree(cloud)
    for variant, seed in MODEL_SEEDS.items():
        v = synth.VARIANTS[variant]
        sd = synth.make_state_dict(variant, seed)
        rng = np.random.RandomState(100 + seed)
        n = 1024
        dirs = rng.standard_normal((n, 3))
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        q = (dirs * (0.5 + rng.uniform(-0.06, 0.06, n))[:, None]).astype(np.float32)
        d = 0.5 - np.linalg.norm(q.astype(np.float64), axis=1)          # + inside (reference convention, sdf.py:204)
        rs = np.random.RandomState(200 + seed)
        items = [orc.assemble_query(cloud, kd, q[i], 300, 1000, rs, bool(v['uniform_subsample'])) for i in range(n)]
        r = np.array([it['patch_radius_ms'] for it in items], np.float64)
        feats = []
        for b in range(0, n, 128):
            _, aux = orc.model_forward(sd, np.stack([it['patch_pts_ps'] for it in items[b:b + 128]]),
                                       np.stack([it['pts_sub_sample_ms'] for it in items[b:b + 128]]), q[b:b + 128],
                                       v['use_point_stn'], v['shared_transformer'], return_aux=True)
            feats.append(aux['fc3_out'])
        F = np.concatenate(feats).astype(np.float64)
        A = np.concatenate([F, np.ones((n, 1))], 1)
        t_sign = np.where(d >= 0, 1.0, -1.0) * np.minimum(np.abs(d) / 0.01, 3.0)
        t_mag = np.arctanh(np.sqrt(np.clip(np.abs(d) / r, 0.0, 0.98)))
        ntr = 896
        W = np.linalg.solve(A[:ntr].T @ A[:ntr] + 0.1 * np.eye(129), A[:ntr].T @ np.stack([t_mag[:ntr], t_sign[:ntr]], 1))
        pred = A @ W
        acc = ((pred[ntr:, 1] >= 0) == (d[ntr:] >= 0)).mean()
        print('fit', variant, 'held-out sign accuracy %.3f' % acc, 'train %.3f' % ((pred[:ntr, 1] >= 0) == (d[:ntr] >= 0)).mean())
        out[variant + '_weight'] = W[:128].T.astype(np.float32).copy()   # [2,128]
        out[variant + '_bias'] = W[128].astype(np.float32).copy()
        out[variant + '_seed'] = np.array(seed)
        out[variant + '_heldout_sign_accuracy'] = np.array(acc)
    np.savez_compressed(FIT_FILE, **out)
    print('written', FIT_FILE)


def e2e(res, variants):
    cloud = synth.make_cloud('sphere', 10000, seed=0)
    eps, seed = 3, 40938661
    qpts = orc.query_grid(cloud, res, eps)
    for variant in variants:
        sd = synth.make_state_dict(variant, MODEL_SEEDS[variant], fitted=True)
        t0 = time.time()
        logits, radius, ids = oracle_shape(sd, variant, cloud, qpts, seed, log='e2e%d %s' % (res, variant))
        sdf = orc.post_process(logits, radius)
        print('e2e', res, variant, 'Q', len(qpts), 'positive fraction %.3f' % (sdf > 0).mean(), '%.0f s' % (time.time() - t0))
        np.savez_compressed(os.path.join(HERE, 'e2e_%s_res%d.npz' % (variant, res)), res=res, eps=eps, seed=seed, points=10000,
                            cloud_seed=0, model_seed=MODEL_SEEDS[variant], Q=len(qpts), logits=logits, radius=radius,
                            ids_checksum=np.array(ids_checksum(ids), dtype=np.uint64), ids_head=ids[:2].copy())


def config1():
    from oracle import ref_shims
    ref_shims.install()
    from source import data_loader as ref_dl
    from source.points_to_surf_model import PointsToSurfModel
    src = '/root/reference/datasets/abc_minimal/04_pts/00994122_57d9d4755722f9d2d7436f0a_trimesh_000.xyz.npy'   # the shape of abc_minimal/testset.txt
    cloud = np.ascontiguousarray(np.load(src).astype(np.float32)[:, :3])
    res, eps, k, S, seed = 32, 3, 300, 1000, 40938661
    variant = 'vanilla'
    v = synth.VARIANTS[variant]
    sd = synth.make_state_dict(variant, MODEL_SEEDS[variant], fitted=True)
    m = PointsToSurfModel(net_size_max=1024, num_points=300, output_dim=2, use_point_stn=v['use_point_stn'], use_feat_stn=1,
                          sym_op='max', use_query_point=True, sub_sample_size=1000, do_augmentation=False,
                          single_transformer=0, shared_transformation=v['shared_transformer'])
    m.load_state_dict(sd)
    m.eval()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, '04_pts'))
        np.save(os.path.join(d, '04_pts', 'shape.xyz.npy'), cloud)
        with open(os.path.join(d, 'testset.txt'), 'w') as f:
            f.write('shape\n')
        ds = ref_dl.PointcloudPatchDataset(
            root=d, shape_list_filename='testset.txt', points_per_patch=k, patch_radius=0.0,
            patch_features=['imp_surf_magnitude', 'imp_surf_sign', 'patch_pts_ids', 'p_index'],
            epsilon=eps, seed=seed, center='mean', cache_capacity=5, pre_processed_patches=True,
            query_grid_resolution=res, sub_sample_size=S, reconstruction=True,
            uniform_subsample=v['uniform_subsample'], fixed_subsample=0, num_workers=0)
        Q = len(ds)
        logits = np.empty((Q, 2), np.float32)
        radius = np.empty((Q,), np.float32)
        subs = []
        for b in range(0, Q, 96):                          # batches like the eval loop (workers=0: sequential stream)
            items = [ds[i] for i in range(b, min(Q, b + 96))]
            x = {kk: torch.stack([it[kk] for it in items]) for kk in ('patch_pts_ps', 'pts_sub_sample_ms', 'imp_surf_query_point_ms')}
            subs.append(x['pts_sub_sample_ms'].numpy().copy())
            logits[b:b + len(items)] = m(x).numpy()
            radius[b:b + len(items)] = [float(it['patch_radius_ms']) for it in items]
    qpts = orc.query_grid(cloud, res, eps)
    assert len(qpts) == Q
    lo, ro, ids = oracle_shape(sd, variant, cloud, qpts, seed)
    assert np.array_equal(np.concatenate(subs), cloud[ids]), 'oracle sub-sample stream != reference'
    assert np.array_equal(ro, radius)
    err = np.abs(lo - logits).max()
    assert err < 2e-4, err
    print('config1: Q', Q, 'oracle-vs-reference max |logit| err', err, 'positive fraction', (logits[:, 1] >= 0).mean())
    np.savez_compressed(os.path.join(HERE, 'e2e_config1_abc_minimal_res32.npz'), cloud=cloud, res=res, eps=eps, seed=seed,
                        model_seed=MODEL_SEEDS[variant], Q=Q, logits=logits, radius=radius,
                        ids_checksum=np.array(ids_checksum(ids), dtype=np.uint64), ids_head=ids[:2].copy())


RNG_CHUNK = 1024


def add_rng_states():
    """Post-process every e2e_*.npz: replay the sequential RandomState sub-sample stream (no network) and store the
    generator state at every RNG_CHUNK-th query, so that the tests can regenerate the ids chunk-parallel."""
    import glob
    for path in sorted(glob.glob(os.path.join(HERE, 'e2e_*.npz'))):
        g = dict(np.load(path))
        if 'rng_keys' in g:
            continue
        if 'cloud' in g:
            cloud, variant = g['cloud'], 'vanilla'
        else:
            cloud = synth.make_cloud('sphere', int(g['points']), seed=int(g['cloud_seed']))
            variant = os.path.basename(path).split('_')[1]
        uniform = bool(synth.VARIANTS[variant]['uniform_subsample'])
        qpts = orc.query_grid(cloud, int(g['res']), int(g['eps']))
        rng = np.random.RandomState(int(g['seed']))
        keys, pos = [], []
        ids = np.empty((len(qpts), 1000), np.int32)
        for i in range(len(qpts)):
            if i % RNG_CHUNK == 0:
                st = rng.get_state()
                assert st[0] == 'MT19937' and st[3] == 0
                keys.append(st[1].copy()); pos.append(st[2])
            ids[i] = orc.sub_sample_ids(1000, cloud, qpts[i], rng, uniform=uniform)
        assert ids_checksum(ids) == int(g['ids_checksum']), path
        g['rng_keys'] = np.stack(keys).astype(np.uint32)
        g['rng_pos'] = np.array(pos, dtype=np.int32)
        g['rng_chunk'] = np.array(RNG_CHUNK)
        np.savez_compressed(path, **g)
        print('rng states added to', os.path.basename(path), len(keys), 'chunks')


if __name__ == '__main__':
    what = sys.argv[1:] or ['fit', 'e2e64', 'config1', 'e2e128', 'states']
    torch.set_num_threads(int(os.environ.get('P2S_GOLDEN_THREADS', '6')))
    for w in what:
        if w == 'fit':
            fit()
        elif w == 'e2e64':
            e2e(64, ['vanilla', 'max'])
        elif w == 'e2e128':
            e2e(128, ['vanilla'])
        elif w == 'config1':
            config1()
        elif w == 'states':
            add_rng_states()
        else:
            raise SystemExit('unknown step ' + w)
