"""GPU: the drop-in boundary -- the reference's entry points, file formats and output tree,
on the analogue of BASELINE config 1 (full_eval plumbing: one shape, rand-init vanilla checkpoint, small grid)."""
import os

import numpy as np
import pytest
import torch

from oracle import p2s_oracle as orc
from points2surf_b200 import synth, mesh_io
from points2surf_b200 import eval as p2s_eval
from points2surf_b200 import sdf as p2s_sdf
from points2surf_b200.model import PointsToSurfModel
from helpers import load_golden, calibrated_state_dict, golden_model_case

pytestmark = pytest.mark.gpu


def _make_dataset(tmp_path, variant='vanilla'):
    g = load_golden('assembly.npz')
    root = tmp_path / 'data'
    (root / '04_pts').mkdir(parents=True)
    np.save(root / '04_pts' / 'shape_a.xyz.npy', g['cloud'])
    np.save(root / '04_pts' / 'shape_b.xyz.npy', np.concatenate([synth.make_cloud('torus', 4000, seed=3), np.zeros((4000, 2), np.float32)], 1))
    (root / 'testset.txt').write_text('shape_a\nshape_b\n\n')
    models = tmp_path / 'models'
    models.mkdir()
    sd = calibrated_state_dict(variant, 31)
    torch.save({'module.' + k: v for k, v in sd.items()}, models / 'p2s_test_model.pth')     # DataParallel-style keys
    torch.save(synth.make_train_opt(variant), models / 'p2s_test_params.pth')                  # pickled Namespace
    return root, models, sd


def test_full_eval_plumbing(tmp_path):
    root, models, sd = _make_dataset(tmp_path)
    out = tmp_path / 'results'
    opt = p2s_eval.parse_arguments(['--indir', str(root), '--outdir', str(out), '--modeldir', str(models), '--models', 'p2s_test',
                                    '--dataset', 'testset.txt', '--query_grid_resolution', '24', '--epsilon', '3',
                                    '--certainty_threshold', '13', '--sigma', '5', '--batchSize', '501', '--workers', '0'])
    assert opt.dataset == 'testset.txt' and opt.seed == 40938661
    opt.reconstruction = True
    p2s_eval.points_to_surf_eval(opt)
    rec = out / 'rec'
    for name, cloud in (('shape_a', load_golden('assembly.npz')['cloud']), ('shape_b', synth.make_cloud('torus', 4000, seed=3))):
        q = np.load(rec / 'query_pts_ms' / (name + '.xyz.npy'))
        d = np.load(rec / 'dist_ms' / (name + '.xyz.npy'))
        assert np.array_equal(q, orc.query_grid(cloud, 24, 3))            # same query set and order as the reference
        assert d.shape == (len(q),) and d.dtype == np.float32 and np.isfinite(d).all()
        assert (d > 0).any() and (d < 0).any()
        assert np.array_equal(np.load(rec / 'eval' / (name + '.xyz.npy')), d)
        assert np.loadtxt(rec / 'eval' / (name + '.xyz.txt')).shape == d.shape
        for sub in ('vis', 'query_pts_ms_vis'):
            v, f = mesh_io.read_ply(str(rec / sub / (name + '.ply')))
            assert v.shape == q.shape and len(f) == 0
    # volume -> mesh stage (full_eval.py:56-62)
    p2s_sdf.implicit_surface_to_mesh_directory(str(rec / 'dist_ms'), str(rec / 'query_pts_ms'), str(rec / 'vol'), str(rec / 'mesh'),
                                               24, 5, 13, 0)
    for name in ('shape_a', 'shape_b'):
        assert (rec / 'vol' / (name + '.off')).read_text().startswith('COFF')
        v, f = mesh_io.read_ply(str(rec / 'mesh' / (name + '.ply')))
        assert len(v) > 0 and len(f) > 0 and f.max() < len(v) and np.abs(v).max() <= 1.0
    # second call is a no-op (mtime rule of file_utils.call_necessary)
    t0 = os.path.getmtime(rec / 'mesh' / 'shape_a.ply')
    p2s_sdf.implicit_surface_to_mesh_directory(str(rec / 'dist_ms'), str(rec / 'query_pts_ms'), str(rec / 'vol'), str(rec / 'mesh'), 24, 5, 13, 0)
    assert os.path.getmtime(rec / 'mesh' / 'shape_a.ply') == t0
    # the non-reconstruction pass needs 05_query_pts; emulate it with a few grid points
    (root / '05_query_pts').mkdir()
    for name in ('shape_a', 'shape_b'):
        np.save(root / '05_query_pts' / (name + '.ply.npy'), np.load(rec / 'query_pts_ms' / (name + '.xyz.npy'))[:50])
    opt.reconstruction = False
    p2s_eval.points_to_surf_eval(opt)
    assert np.load(out / 'eval' / 'eval' / 'shape_a.xyz.npy').shape == (50,)


def test_model_module_is_a_drop_in():
    sd, inp, g = golden_model_case('vanilla')
    m = PointsToSurfModel(net_size_max=1024, num_points=300, output_dim=2, use_point_stn=True, use_feat_stn=True, sym_op='max',
                          use_query_point=True, sub_sample_size=1000, do_augmentation=False, single_transformer=0,
                          shared_transformation=True, precision='fp32')
    ref_keys = list(synth.make_state_dict('vanilla', 0).keys())
    assert list(m.state_dict().keys()) == ref_keys
    m.load_state_dict({'module.' + k: v for k, v in sd.items()})        # checkpoint saved from the DataParallel wrapper
    m.eval()
    x = {k: torch.from_numpy(v.copy()).cuda() for k, v in inp.items()}
    with torch.no_grad():
        y = m(x)
    assert np.abs(y.cpu().numpy() - g['logits']).max() < 2e-3
    # like the reference, forward centres the caller's sub-sample in place (points_to_surf_model.py:303)
    assert np.allclose(x['pts_sub_sample_ms'].cpu().numpy(), inp['pts_sub_sample_ms'] - inp['imp_surf_query_point_ms'][:, None, :], atol=1e-6)
    with pytest.raises(ValueError):
        PointsToSurfModel(sym_op='sum', output_dim=2)
    with pytest.raises(ValueError):
        PointsToSurfModel(single_transformer=True, output_dim=2)


def test_radius_ablation_checkpoint_reconstructs(tmp_path):
    """train_opt.patch_radius > 0 (experiments/train_p2s_*_radius.sh): ball-query patches through the eval entry point."""
    root, models, _ = _make_dataset(tmp_path)
    opt_ns = synth.make_train_opt('vanilla')
    opt_ns.patch_radius = 0.2
    torch.save(opt_ns, models / 'p2s_test_params.pth')
    opt = p2s_eval.parse_arguments(['--indir', str(root), '--outdir', str(tmp_path / 'r'), '--modeldir', str(models), '--models', 'p2s_test',
                                    '--query_grid_resolution', '16', '--epsilon', '3'])
    opt.reconstruction = True
    p2s_eval.points_to_surf_eval(opt)
    d = np.load(tmp_path / 'r' / 'rec' / 'dist_ms' / 'shape_a.xyz.npy')
    assert np.isfinite(d).all() and np.abs(d).max() <= 1.0        # tanh^2, not rescaled by a radius (eval.py:364-368)


def test_fixed_uniform_subsample_checkpoint_reconstructs_like_the_reference(tmp_path):
    """train_opt.fixed_subsample = 1 with uniform_subsample = 1 (experiments/train_p2s_vanilla_uniform_subsample.sh): every query of a
    shape sees the sub-sample pts[RandomState(42).randint(0, N, S)] (utils.py:210-216).  The eval entry point runs this variant
    stage by stage; a few queries are replayed on the CPU oracle with exactly those ids (fp32 engine, same tolerance as the
    fused-pipeline replay)."""
    root, models, sd = _make_dataset(tmp_path)
    opt_ns = synth.make_train_opt('vanilla')
    opt_ns.uniform_subsample = 1
    opt_ns.fixed_subsample = 1
    torch.save(opt_ns, models / 'p2s_test_params.pth')
    res, eps = 16, 3
    opt = p2s_eval.parse_arguments(['--indir', str(root), '--outdir', str(tmp_path / 'r'), '--modeldir', str(models), '--models', 'p2s_test',
                                    '--query_grid_resolution', str(res), '--epsilon', str(eps), '--precision', 'fp32', '--batchSize', '300'])
    opt.reconstruction = True
    p2s_eval.points_to_surf_eval(opt)
    cloud = load_golden('assembly.npz')['cloud'][:, :3].astype(np.float32)
    qpts = orc.query_grid(cloud, res, eps)
    assert np.array_equal(np.load(tmp_path / 'r' / 'rec' / 'query_pts_ms' / 'shape_a.xyz.npy'), qpts)
    d = np.load(tmp_path / 'r' / 'rec' / 'dist_ms' / 'shape_a.xyz.npy')
    assert d.shape == (len(qpts),) and np.isfinite(d).all()
    ids = np.random.RandomState(42).randint(low=0, high=len(cloud), size=1000)
    sel = [0, 1, len(qpts) // 2, len(qpts) - 1]
    kd = orc.make_kdtree(cloud)
    patches, radii = zip(*[(orc.knn_patch(cloud, kd, qpts[i], 300)[1:]) for i in sel])
    sub = np.repeat(cloud[ids][None], len(sel), axis=0)
    logits = orc.model_forward(sd, np.stack(patches), sub, qpts[sel], True, True)
    np.testing.assert_allclose(d[sel], orc.post_process(logits, np.array(radii)), rtol=0, atol=2e-4)
    # the distance-weighted sub-sample has no fixed variant here
    opt_ns.uniform_subsample = 0
    torch.save(opt_ns, models / 'p2s_test_params.pth')
    with pytest.raises(ValueError):
        p2s_eval.points_to_surf_eval(opt)


def test_unsupported_options_raise(tmp_path):
    root, models, _ = _make_dataset(tmp_path)
    opt_ns = synth.make_train_opt('vanilla')
    opt_ns.sym_op = 'sum'
    torch.save(opt_ns, models / 'p2s_test_params.pth')
    opt = p2s_eval.parse_arguments(['--indir', str(root), '--outdir', str(tmp_path / 'r'), '--modeldir', str(models), '--models', 'p2s_test',
                                    '--query_grid_resolution', '16', '--epsilon', '3'])
    opt.reconstruction = True
    with pytest.raises(ValueError):
        p2s_eval.points_to_surf_eval(opt)


def test_rotation_augmented_pass_matches_reference_dataset():
    """full_eval's non-reconstruction pass (full_eval.py:31-41): the reference dataset rotates patch, sub-sample and query by
    a per-sample random rotation (data_loader.py:381-393).  The mirror's rotations (stream, order, conventions) applied
    to the GPU kNN patches reproduce tests/golden/rotation.npz, made by the unmodified dataset class; float32 rotation
    here vs float64-then-cast there: tolerance 1e-6."""
    g = load_golden('rotation.npz')
    cloud = load_golden('assembly.npz')['cloud']
    q = torch.from_numpy(g['query_pts']).cuda()
    pts = torch.from_numpy(cloud).cuda()
    from points2surf_b200 import ops
    _, patch, radius = ops.knn_patch(pts, q, int(g['k']))
    assert np.array_equal(radius.cpu().numpy(), g['radius'])
    R = torch.from_numpy(p2s_eval._random_rotations(np.random.RandomState(int(g['seed'])), q.shape[0])).cuda()
    sub = ops.gather_points(pts, ops.subsample(pts, q, 1000, True, 1))
    patch_r, sub_r, q_r = p2s_eval._rotate_inputs(patch, sub, q, R)
    assert np.abs(patch_r.cpu().numpy() - g['patch_rot']).max() < 1e-6
    assert np.abs(q_r.cpu().numpy() - g['query_rot']).max() < 1e-6
    # rotations preserve the geometry the network sees: distances to the (rotated) query are unchanged
    d0 = (sub - q[:, None, :]).norm(dim=2)
    d1 = (sub_r - q_r[:, None, :]).norm(dim=2)
    assert torch.allclose(d0, d1, atol=1e-5)
