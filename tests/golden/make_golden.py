"""Generate tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Every fixture is produced by reference code (imported through oracle/ref_shims.py) and, in the
same run, compared with the oracle restatement (oracle/p2s_oracle.py); the script fails if
they disagree, so a committed fixture certifies "oracle == reference" on that input.
"""
import os
import sys
import tempfile
import warnings

import numpy as np

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
ref_shims.install()
import torch  # noqa: E402
from oracle import p2s_oracle as orc  # noqa: E402
from points2surf_b200 import synth  # noqa: E402

from source import sdf as ref_sdf  # noqa: E402
from source import sdf_nn as ref_sdf_nn  # noqa: E402
from source import data_loader as ref_dl  # noqa: E402
from source.points_to_surf_model import PointsToSurfModel  # noqa: E402

torch.set_grad_enabled(False)
MODEL_SEEDS = {'vanilla': 6, 'max': 4, 'uniform': 8}


def ref_model(variant, seed, fc4_bias=None):
    v = synth.VARIANTS[variant]
    m = PointsToSurfModel(net_size_max=1024, num_points=300, output_dim=2,
                          use_point_stn=v['use_point_stn'], use_feat_stn=1, sym_op='max',
                          use_query_point=True, sub_sample_size=1000, do_augmentation=False,
                          single_transformer=0, shared_transformation=v['shared_transformer'])
    sd = synth.make_state_dict(variant, seed)
    if fc4_bias is not None:
        sd['fc4.bias'] = torch.from_numpy(np.asarray(fc4_bias, dtype=np.float32))
    m.load_state_dict(sd)
    m.eval()
    return m


def gen_model():
    for variant, seed in MODEL_SEEDS.items():
        v = synth.VARIANTS[variant]
        inp = synth.make_model_inputs(8, seed=100 + seed)
        # calibrate the output bias on this batch so that the sign classes are mixed
        # (a rand-init net is otherwise almost always one-signed)
        x = {k: torch.from_numpy(a.copy()) for k, a in inp.items()}
        raw = ref_model(variant, seed)(x).numpy()
        fc4_bias = (synth.make_state_dict_numpy(variant, seed)['fc4.bias'] - np.median(raw, axis=0)).astype(np.float32)
        m = ref_model(variant, seed, fc4_bias)
        x = {k: torch.from_numpy(a.copy()) for k, a in inp.items()}
        logits = m(x).numpy()
        # the reference centred the sub-sample in place (model.py:303)
        assert np.allclose(x['pts_sub_sample_ms'].numpy(),
                           inp['pts_sub_sample_ms'] - inp['imp_surf_query_point_ms'][:, None, :])
        sd = synth.make_state_dict(variant, seed)
        sd['fc4.bias'] = torch.from_numpy(fc4_bias)
        o, aux = orc.model_forward(sd, inp['patch_pts_ps'], inp['pts_sub_sample_ms'],
                                   inp['imp_surf_query_point_ms'], v['use_point_stn'],
                                   v['shared_transformer'], return_aux=True)
        err = np.abs(o - logits).max()
        assert err < 1e-4, (variant, err)
        radius = np.linspace(0.05, 0.3, 8).astype(np.float32)
        pred = torch.from_numpy(logits.copy())
        mag = ref_sdf_nn.post_process_magnitude(pred[:, 0:1]) * torch.from_numpy(radius).unsqueeze(1)
        sgn = ref_sdf_nn.post_process_sign(pred[:, 1:2])
        sdf_ref = (mag.squeeze() * sgn.squeeze()).numpy()
        assert np.array_equal(orc.post_process(logits, radius), sdf_ref)
        print('model', variant, 'oracle-vs-reference max abs err', err, 'pos frac', (logits[:, 1] >= 0).mean())
        np.savez_compressed(os.path.join(HERE, 'model_%s.npz' % variant), seed=seed, input_seed=100 + seed, fc4_bias=fc4_bias,
                            logits=logits, radius=radius, sdf=sdf_ref,
                            feat_global_max=aux['feat_global_max'][:2], feat_local_max=aux['feat_local_max'][:2],
                            input_checksum=np.float64(sum(float(np.abs(a).sum()) for a in inp.values())))


def gen_assembly():
    src = '/root/reference/datasets/abc_minimal/04_pts/00011084_fddd53ce45f640f3ab922328_trimesh_019.xyz.npy'
    full = np.load(src).astype(np.float32)[:, :3]
    cloud = full[np.random.RandomState(7).choice(len(full), 6000, replace=False)].copy()
    res, eps, k, S, seed = 32, 3, 300, 1000, 40938661
    out = dict(cloud=cloud, res=res, eps=eps, k=k, S=S, seed=seed)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, '04_pts'))
        np.save(os.path.join(d, '04_pts', 'shape.xyz.npy'), cloud)
        with open(os.path.join(d, 'testset.txt'), 'w') as f:
            f.write('shape\n')
        for uniform in (0, 1):
            ds = ref_dl.PointcloudPatchDataset(
                root=d, shape_list_filename='testset.txt', points_per_patch=k, patch_radius=0.0,
                patch_features=['imp_surf_magnitude', 'imp_surf_sign', 'patch_pts_ids', 'p_index'],
                epsilon=eps, seed=seed, center='mean', cache_capacity=5, pre_processed_patches=True,
                query_grid_resolution=res, sub_sample_size=S, reconstruction=True,
                uniform_subsample=uniform, fixed_subsample=0, num_workers=0)
            Q = len(ds)
            qsel = list(range(6)) + [Q // 2, Q - 1]
            shape = ds.shape_cache.get(0)
            qpts = shape.imp_surf_query_point_ms
            # oracle: candidate grid
            idx = orc.query_grid_indices(cloud, res, eps)
            assert np.array_equal(idx, orc.query_grid_indices_shifts(cloud, res, eps))
            assert np.array_equal(orc.volume_space_to_model_space(idx, res).astype(np.float32), qpts)
            kd = orc.make_kdtree(cloud)
            rng_o = np.random.RandomState(seed)
            items, o_items = [], []
            # the reference consumes its global-sample RNG sequentially: replay queries 0..max in order
            for qi in range(6):
                items.append(ds[qi])
                o_items.append(orc.assemble_query(cloud, kd, qpts[qi], k, S, rng_o, bool(uniform)))
            for it, oi in zip(items, o_items):
                assert np.array_equal(it['patch_pts_ps'].numpy(), oi['patch_pts_ps'])
                assert np.float32(it['patch_radius_ms'].numpy()) == oi['patch_radius_ms']
                assert np.array_equal(it['pts_sub_sample_ms'].numpy(), oi['pts_sub_sample_ms'])
                assert np.array_equal(it['imp_surf_query_point_ms'].numpy(), oi['imp_surf_query_point_ms'])
                bid, bd2 = orc.knn_bruteforce(cloud, oi['imp_surf_query_point_ms'], k)
                assert set(bid.tolist()) == set(oi['patch_pts_ids'].tolist())
            tag = 'uni' if uniform else 'wgt'
            out['query_idx'] = idx.astype(np.int16)
            out['patch_ids'] = np.stack([oi['patch_pts_ids'] for oi in o_items])
            out['patch_ps'] = np.stack([it['patch_pts_ps'].numpy() for it in items])
            out['radius'] = np.array([np.float32(it['patch_radius_ms'].numpy()) for it in items])
            out['sub_ids_' + tag] = np.stack([oi['sub_sample_ids'] for oi in o_items]).astype(np.int32)
            print('assembly', tag, 'Q', Q, 'radius', out['radius'][:3])
    np.savez_compressed(os.path.join(HERE, 'assembly.npz'), **out)


def gen_ball():
    """Ball-query patches (radius ablations) through the UNMODIFIED reference dataset (patch_radius > 0), compared with
    orc.ball_patch in the same run.  Cloud = the 6 000-point abc_minimal subset of assembly.npz."""
    cloud = np.load(os.path.join(HERE, 'assembly.npz'))['cloud']
    res, eps, S, seed = 32, 3, 1000, 40938661
    out = dict(res=res, eps=eps, seed=seed)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, '04_pts'))
        np.save(os.path.join(d, '04_pts', 'shape.xyz.npy'), cloud)
        with open(os.path.join(d, 'testset.txt'), 'w') as f:
            f.write('shape\n')
        for tag, radius, P in (('small', 0.05, 10), ('large', 0.2, 300)):
            ds = ref_dl.PointcloudPatchDataset(
                root=d, shape_list_filename='testset.txt', points_per_patch=P, patch_radius=radius,
                patch_features=['imp_surf_magnitude', 'imp_surf_sign', 'patch_pts_ids', 'p_index'],
                epsilon=eps, seed=seed, center='mean', cache_capacity=5, pre_processed_patches=True,
                query_grid_resolution=res, sub_sample_size=S, reconstruction=True,
                uniform_subsample=1, fixed_subsample=0, num_workers=0)
            Q = len(ds)
            qpts = ds.shape_cache.get(0).imp_surf_query_point_ms
            kd = orc.make_kdtree(cloud)
            rng_o = np.random.RandomState(seed)
            sel = np.linspace(0, Q - 1, 24).astype(np.int64)
            counts, patches, radii = [], [], []
            for qi in sel:                      # the dataset rng is consumed in call order: same order on both sides
                it = ds[int(qi)]
                oid, ops_, cnt = orc.ball_patch(cloud, kd, qpts[qi], radius, P, rng_o)
                assert np.array_equal(it['patch_pts_ps'].numpy(), ops_), (tag, qi)
                assert np.float32(it['patch_radius_ms'].numpy()) == np.float32(radius)
                counts.append(cnt); patches.append(ops_); radii.append(np.float32(it['patch_radius_ms'].numpy()))
            counts = np.array(counts)
            print('ball', tag, 'radius', radius, 'P', P, 'in-ball counts', counts.min(), '..', counts.max(),
                  'padded', int((counts < P).sum()), 'sub-set', int((counts > P).sum()))
            assert (counts < P).any() and (counts > P).any() or tag == 'large'
            out[tag + '_radius'] = radius
            out[tag + '_P'] = P
            out[tag + '_query_sel'] = sel
            out[tag + '_counts'] = counts
            out[tag + '_patch_ps'] = np.stack(patches)
    np.savez_compressed(os.path.join(HERE, 'ball.npz'), **out)
    print('ball.npz written: oracle == reference')


def gen_rotation():
    """The non-reconstruction pass of the reference dataset (reconstruction=False: per-sample random rotation of patch,
    sub-sample and query, source/data_loader.py:381-393) on the 6 000-point abc subset: rotated patches and query points of
    the first queries.  (trimesh is absent: random_rotation_matrix / transform_points are the shim's restatement of
    trimesh.transformations; what this pins is the reference's stream consumption, application order and conventions.)"""
    cloud = np.load(os.path.join(HERE, 'assembly.npz'))['cloud']
    seed, k, S, n = 40938661, 300, 1000, 8
    rng = np.random.RandomState(3)
    qpts = (cloud[rng.choice(len(cloud), 40, replace=False)] + rng.normal(0, 0.02, (40, 3))).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        for sub in ('04_pts', '05_query_pts', '05_query_dist'):
            os.makedirs(os.path.join(d, sub))
        np.save(os.path.join(d, '04_pts', 'shape.xyz.npy'), cloud)
        np.save(os.path.join(d, '05_query_pts', 'shape.ply.npy'), qpts)
        np.save(os.path.join(d, '05_query_dist', 'shape.ply.npy'), rng.normal(0, 0.01, 40).astype(np.float32))
        with open(os.path.join(d, 'testset.txt'), 'w') as f:
            f.write('shape\n')
        ds = ref_dl.PointcloudPatchDataset(
            root=d, shape_list_filename='testset.txt', points_per_patch=k, patch_radius=0.0,
            patch_features=['imp_surf_magnitude', 'imp_surf_sign', 'patch_pts_ids', 'p_index'],
            epsilon=3, seed=seed, center='mean', cache_capacity=5, pre_processed_patches=True,
            query_grid_resolution=None, sub_sample_size=S, reconstruction=False,
            uniform_subsample=1, fixed_subsample=0, num_workers=0)
        items = [ds[i] for i in range(n)]
    from points2surf_b200 import eval as my_eval
    R = my_eval._random_rotations(np.random.RandomState(seed), n)
    kd = orc.make_kdtree(cloud)
    for i, it in enumerate(items):
        _, patch_ps, radius = orc.knn_patch(cloud, kd, qpts[i], k)
        want = (R[i].astype(np.float64) @ patch_ps.astype(np.float64).T).T
        assert np.abs(it['patch_pts_ps'].numpy() - want).max() < 1e-6, i
        assert np.abs(it['imp_surf_query_point_ms'].numpy() - R[i].astype(np.float64) @ qpts[i].astype(np.float64)).max() < 1e-6
        # the rotated sub-sample is a rotation of cloud rows: un-rotate and look the rows up
        back = (R[i].astype(np.float64).T @ it['pts_sub_sample_ms'].numpy().astype(np.float64).T).T
        dist, _ = kd.query(back, 1)
        assert dist.max() < 1e-5
    np.savez_compressed(os.path.join(HERE, 'rotation.npz'), seed=seed, k=k, query_pts=qpts[:n],
                        patch_rot=np.stack([it['patch_pts_ps'].numpy() for it in items]),
                        query_rot=np.stack([it['imp_surf_query_point_ms'].numpy() for it in items]),
                        radius=np.array([np.float32(it['patch_radius_ms'].numpy()) for it in items]),
                        dist_sign=np.array([float(it['imp_surf_dist_sign_ms'].numpy()[0]) for it in items]))
    print('rotation.npz written: mirror rotations == reference dataset (reconstruction=False)')


def gen_volume():
    out = {}
    for name, res, noise in (('sphere', 32, 0.0), ('noisy', 40, 0.15)):
        cloud = synth.make_cloud('sphere', 3000, seed=5)
        qpts = ref_sdf.get_voxel_centers_grid_smaller_pc(cloud, res, 3)
        assert np.array_equal(qpts, orc.query_grid(cloud, res, 3))
        d = (0.5 - np.linalg.norm(qpts, axis=1)).astype(np.float32)   # + inside (trimesh convention)
        rng = np.random.RandomState(11)
        flip = rng.random_sample(len(d)) < noise
        d[flip] = -d[flip]
        d[rng.randint(0, len(d), 5)] = 0.0                                  # exact zeros in the band
        vol_ref = np.zeros((res,) * 3)
        vol_ref = ref_sdf.add_samples_to_volume(vol_ref, qpts, d)
        vol_scatter = vol_ref.copy()
        vol_ref = ref_sdf.propagate_sign(vol_ref, 5, 13)
        vol_o = np.zeros((res,) * 3)
        vol_o = orc.add_samples_to_volume(vol_o, qpts, d)
        assert np.array_equal(vol_o, vol_scatter)
        vol_o, iters = orc.propagate_sign(vol_o, 5, 13)
        assert np.array_equal(vol_o, vol_ref), name
        print('volume', name, 'res', res, 'Q', len(d), 'iterations', iters,
              'zeros left', int((vol_ref == 0).sum()))
        out[name + '_res'] = res
        out[name + '_qpts'] = qpts
        out[name + '_dist'] = d
        out[name + '_vol'] = vol_ref.astype(np.float32)     # values are float32 dists or -1/0/+1: exact
        out[name + '_iters'] = iters
        # a different (sigma, threshold) pair as well
        v2 = ref_sdf.propagate_sign(vol_scatter.copy(), 3, 5)
        v2o, _ = orc.propagate_sign(vol_scatter.copy(), 3, 5)
        assert np.array_equal(v2, v2o)
        out[name + '_vol_s3t5'] = v2.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'volume.npz'), **out)


def gen_grid():
    out = {}
    for name, kind, n, res, eps in (('sphere64e3', 'sphere', 10000, 64, 3), ('torus48e4', 'torus', 4000, 48, 4),
                                    ('box40e2', 'box', 3000, 40, 2), ('sphere32e5', 'sphere', 2000, 32, 5)):
        cloud = synth.make_cloud(kind, n, seed=2)
        q = ref_sdf.get_voxel_centers_grid_smaller_pc(cloud, res, eps)
        idx = orc.query_grid_indices(cloud, res, eps)
        assert np.array_equal(orc.volume_space_to_model_space(idx, res).astype(np.float32), q)
        assert np.array_equal(idx, orc.query_grid_indices_shifts(cloud, res, eps)), name
        out[name + '_count'] = len(q)
        out[name + '_idxsum'] = idx.astype(np.int64).sum(axis=0)
        out[name + '_lin_xor'] = np.bitwise_xor.reduce((idx[:, 0] * res + idx[:, 1]) * res + idx[:, 2])
        out[name + '_first'] = q[:4]
        out[name + '_last'] = q[-4:]
        print('grid', name, len(q))
    np.savez_compressed(os.path.join(HERE, 'grid.npz'), **out)


def gen_evaluation():
    """eval_predictions / compare_predictions_binary_tensors / print_list_of_dicts of source/base/evaluation.py
    (imports unmodified) on seeded inputs; the report text is the fixture."""
    import contextlib
    import io
    from source.base import evaluation as ref_eval
    rng = np.random.RandomState(11)
    names = ['aa_b', 'cc', 'shape_with_long_name']
    pred = {n: (rng.randn(2000) * (rng.rand(2000) > 0.1)).astype(np.float32) for n in names}
    gt = {n: rng.randn(2000).astype(np.float32) for n in names}
    out = {'names': np.array(names)}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(d + '/pred'); os.makedirs(d + '/gt')
        for n in names:
            np.save(d + '/pred/' + n + '.xyz.npy', pred[n]); np.save(d + '/gt/' + n + '.ply.npy', gt[n])
            out['pred_' + n] = pred[n]; out['gt_' + n] = gt[n]
        for uns in (0, 1):
            with contextlib.redirect_stdout(io.StringIO()):
                ref_eval.eval_predictions(d + '/pred', d + '/gt', d + '/rep.csv', unsigned=bool(uns))
            out['report_unsigned%d' % uns] = np.array(open(d + '/rep.csv').read())
    a, b = torch.from_numpy(rng.randn(64, 3)), torch.from_numpy(rng.randn(64, 3))
    res = ref_eval.compare_predictions_binary_tensors(a, b, 'cmp')
    out['bin_a'], out['bin_b'] = a.numpy(), b.numpy()
    keys = sorted(k for k in res if k != 'comp_name')
    out['bin_keys'] = np.array(keys); out['bin_vals'] = np.array([res[k] for k in keys], dtype=np.float64)
    rows = [{'file': 'abc_def_ghi_jkl', 'mse': 1.23456789, 'x': 2.0}, {'file': 'zz', 'mse': 0.5, 'x': -1.0}]
    for mode in ('latex', 'csv'):
        with contextlib.redirect_stdout(io.StringIO()):
            out['table_' + mode] = np.array('\n'.join(ref_eval.print_list_of_dicts(rows, None, mode)))
    np.savez_compressed(os.path.join(HERE, 'evaluation.npz'), **out)
    print('evaluation.npz written')


TRAIN_SEEDS = {'vanilla': 21, 'max': 22, 'uniform': 23}


def train_digest_indices(name, numel, count=48):
    """Deterministic sample positions inside a tensor (shared with tests/helpers.py)."""
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return np.random.RandomState(h).randint(0, numel, size=min(count, numel))


def gen_train():
    """One training iteration of the UNMODIFIED reference (PointsToSurfModel.train(), compute_loss, optim.SGD with
    momentum 0.9, source/points_to_surf_train.py:406,441-461,537-563) on a seeded batch of 32 queries
    (tests/helpers_train.py:make_train_batch; only its checksum is stored).  The fixture keeps logits, losses and a
    digest (norm + 48 sampled entries) of every gradient / updated tensor / running statistic.
    oracle/train_oracle.py must reproduce all of it (asserted here).  Gradients are compared in the L2 sense: max-pool
    arg-max and ReLU decisions flip under fp32 rounding, which moves single entries by O(1/batch) of their size."""
    from oracle import train_oracle
    from source import points_to_surf_train as ref_train
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers_train import make_train_batch
    B = 32
    for variant, seed in TRAIN_SEEDS.items():
        v = synth.VARIANTS[variant]
        torch.manual_seed(0)
        m = PointsToSurfModel(net_size_max=1024, num_points=300, output_dim=2, use_point_stn=v['use_point_stn'],
                              use_feat_stn=1, sym_op='max', use_query_point=0, sub_sample_size=1000, do_augmentation=0,
                              single_transformer=0, shared_transformation=v['shared_transformer'])
        sd = synth.make_state_dict(variant, seed=seed)
        m.load_state_dict(sd)
        m.train()
        batch = make_train_batch(B, 300, 1000, seed=seed)
        opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
        opt.zero_grad()
        bd = {k: t.clone() for k, t in batch.items()}
        with torch.enable_grad():
            pred = m(bd)
            loss = ref_train.compute_loss(pred=pred, batch_data=bd, outputs=['imp_surf_magnitude', 'imp_surf_sign'],
                                          output_loss_weights={'imp_surf_magnitude': 1.0, 'imp_surf_sign': 1.0},
                                          fixed_radius=False)
            sum(loss).backward()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        opt.step()
        new_sd = {k: t.detach().clone() for k, t in m.state_dict().items()}
        orc_out = train_oracle.train_iteration(sd, batch, v['use_point_stn'], v['shared_transformer'], lr=0.01, momentum=0.9)
        assert np.allclose(orc_out['logits'].numpy(), pred.detach().numpy(), atol=2e-5), variant
        assert abs(orc_out['losses'][0] - float(loss[0])) < 1e-6 and abs(orc_out['losses'][1] - float(loss[1])) < 1e-6
        nscale = max(float(g.norm()) for g in grads.values())
        for k, g in grads.items():
            err = float((orc_out['grads'][k] - g).norm())
            assert err <= 2e-2 * float(g.norm()) + 1e-4 * nscale, (variant, k, err, float(g.norm()))
        for k, t in new_sd.items():
            if k.endswith('num_batches_tracked'):
                continue
            assert float((orc_out['new_state'][k] - t).norm()) <= 1e-3 * (1e-3 + float(t.norm())), (variant, k)
        out = {'batch': np.array(B), 'input_checksum': np.array(sum(float(t.double().abs().sum()) for t in batch.values()))}
        out['logits'] = pred.detach().numpy()
        out['losses'] = np.array([float(loss[0]), float(loss[1])])
        names = sorted(grads)
        out['names'] = np.array(names)
        out['grad_norm'] = np.array([float(grads[k].double().norm()) for k in names])
        out['grad_max'] = np.array([float(grads[k].abs().max()) for k in names])
        out['grad_samples'] = np.stack([np.pad(grads[k].reshape(-1).numpy()[train_digest_indices(k, grads[k].numel())],
                                               (0, 48 - min(48, grads[k].numel()))) for k in names])
        out['new_samples'] = np.stack([np.pad(new_sd[k].reshape(-1).numpy()[train_digest_indices(k, new_sd[k].numel())],
                                              (0, 48 - min(48, new_sd[k].numel()))) for k in names])
        bn = sorted(k for k in new_sd if k.endswith('running_mean') or k.endswith('running_var'))
        out['buffer_names'] = np.array(bn)
        out['buffer_samples'] = np.stack([np.pad(new_sd[k].numpy()[train_digest_indices(k, new_sd[k].numel())],
                                                 (0, 48 - min(48, new_sd[k].numel()))) for k in bn])
        np.savez_compressed(os.path.join(HERE, 'train_%s.npz' % variant), **out)
        print('train_%s.npz written: losses %s, oracle == reference' % (variant, out['losses']))


def gen_samplers():
    """Index sequences of the reference's samplers (source/data_loader.py:71-176) on a fake data source."""
    from types import SimpleNamespace
    from points2surf_b200 import samplers as mine
    ds = SimpleNamespace(shape_names=['a', 'b', 'c', 'd'], shape_patch_count=[50, 7, 120, 33])
    out = {'shape_patch_count': np.array(ds.shape_patch_count)}
    cases = {
        'seq': (lambda m: m.SequentialPointcloudPatchSampler(ds)),
        'ssr_seq_shapes': (lambda m: m.SequentialShapeRandomPointcloudPatchSampler(ds, patches_per_shape=20, seed=40938661, sequential_shapes=True, identical_epochs=False)),
        'ssr_perm_shapes': (lambda m: m.SequentialShapeRandomPointcloudPatchSampler(ds, patches_per_shape=40, seed=3627473, sequential_shapes=False, identical_epochs=False)),
        'ssr_identical': (lambda m: m.SequentialShapeRandomPointcloudPatchSampler(ds, patches_per_shape=10, seed=5, sequential_shapes=False, identical_epochs=True)),
        'random': (lambda m: m.RandomPointcloudPatchSampler(ds, patches_per_shape=25, seed=3627473, identical_epochs=False)),
        'random_identical': (lambda m: m.RandomPointcloudPatchSampler(ds, patches_per_shape=1000, seed=9, identical_epochs=True)),
    }
    for name, make in cases.items():
        r, m = make(ref_dl), make(mine)
        assert len(r) == len(m)
        for epoch in range(2):
            a, b = np.array(list(iter(r)), dtype=np.int64), np.array(list(iter(m)), dtype=np.int64)
            assert np.array_equal(a, b), (name, epoch)
            out['%s_epoch%d' % (name, epoch)] = a
            if hasattr(r, 'shape_patch_inds') and r.shape_patch_inds is not None:
                for si in range(4):
                    assert np.array_equal(np.asarray(r.shape_patch_inds[si]), np.asarray(m.shape_patch_inds[si]))
                    out['%s_epoch%d_inds%d' % (name, epoch, si)] = np.asarray(r.shape_patch_inds[si], dtype=np.int64)
        out[name + '_len'] = np.array(len(r))
    np.savez_compressed(os.path.join(HERE, 'samplers.npz'), **out)
    print('samplers.npz written, mirror == reference')


def gen_train_init():
    """Default initialisation of PointsToSurfModel under a fixed torch seed (per-tensor sum + first entries) and the
    defaults of the reference's training argument parser; points2surf_b200.points_to_surf_train must reproduce both."""
    import json
    from source import points_to_surf_train as ref_train
    from points2surf_b200 import points_to_surf_train as mine
    out = {}
    for variant in ('vanilla', 'max', 'uniform'):
        v = synth.VARIANTS[variant]
        torch.manual_seed(3627473)
        m = PointsToSurfModel(net_size_max=1024, num_points=300, output_dim=2, use_point_stn=v['use_point_stn'], use_feat_stn=True,
                              sym_op='max', use_query_point=True, sub_sample_size=1000, do_augmentation=True,
                              single_transformer=0, shared_transformation=v['shared_transformer'])
        ref_sd = m.state_dict()
        torch.manual_seed(3627473)
        my_sd = mine.initial_state_dict(v['use_point_stn'], v['shared_transformer'], 1024, 2)
        assert list(ref_sd.keys()) == list(my_sd.keys()), variant
        for k in ref_sd:
            assert ref_sd[k].shape == my_sd[k].shape and torch.equal(ref_sd[k], my_sd[k]), (variant, k)
        names = [k for k in ref_sd if ref_sd[k].is_floating_point()]
        out[variant + '_names'] = np.array(names)
        out[variant + '_sums'] = np.array([float(ref_sd[k].double().sum()) for k in names])
        out[variant + '_first'] = np.array([float(ref_sd[k].reshape(-1)[0]) for k in names])
    ref_defaults = vars(ref_train.parse_arguments([]))
    my_defaults = vars(mine.parse_arguments([]))
    assert ref_defaults == my_defaults, {k: (ref_defaults.get(k), my_defaults.get(k)) for k in set(ref_defaults) | set(my_defaults)
                                          if ref_defaults.get(k) != my_defaults.get(k)}
    out['parser_defaults_json'] = np.array(json.dumps(ref_defaults, sort_keys=True))
    np.savez_compressed(os.path.join(HERE, 'train_init.npz'), **out)
    print('train_init.npz written: initialisation and parser defaults == reference')


if __name__ == '__main__':
    gen_grid()
    gen_train_init()
    gen_samplers()
    gen_train()
    gen_evaluation()
    gen_volume()
    gen_assembly()
    gen_model()
    print('golden fixtures written to', HERE)
