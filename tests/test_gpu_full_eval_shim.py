"""The `source/` drop-in package: the call sequence of the reference's full_eval.full_eval (full_eval.py:17-75) executed
through `from source import points_to_surf_eval, sdf` / `from source.base import evaluation`, on a small dataset with
ground-truth query distances (so the non-reconstruction pass and rme_comp_res.csv are produced too) and reference meshes
(so hausdorff_dist_pred_rec.csv is)."""
import os
import sys

import numpy as np
import pytest
import torch

from points2surf_b200 import synth, ops, mesh_io

pytestmark = pytest.mark.gpu


def _dataset(root, models):
    names = ['ball_a', 'ball_b']
    for sub in ('04_pts', '05_query_pts', '05_query_dist', '03_meshes'):
        os.makedirs(os.path.join(root, sub))
    os.makedirs(models)
    rng = np.random.RandomState(0)
    # reference meshes: the analytic sphere, meshed by the MC kernel
    R = 48
    g = (np.arange(R) + 0.5) / R * 2 - 1
    X, Y, Z = np.meshgrid(g, g, g, indexing='ij')
    vol = torch.from_numpy(np.clip(0.5 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2), -1, 1).astype(np.float32)).cuda()
    v, f = ops.marching_cubes(vol, 0.0)
    for i, n in enumerate(names):
        cloud = synth.make_cloud('sphere', 6000, seed=i)
        q = (cloud[rng.choice(6000, 200, replace=False)] + rng.normal(0, 0.02, (200, 3))).astype(np.float32)
        np.save(os.path.join(root, '04_pts', n + '.xyz.npy'), cloud)
        np.save(os.path.join(root, '05_query_pts', n + '.ply.npy'), q)
        np.save(os.path.join(root, '05_query_dist', n + '.ply.npy'), (0.5 - np.linalg.norm(q, axis=1)).astype(np.float32))
        mesh_io.write_ply(os.path.join(root, '03_meshes', n + '.ply'), v.cpu().numpy(), f.cpu().numpy())
    with open(os.path.join(root, 'testset.txt'), 'w') as fp:
        fp.write('\n'.join(names) + '\n')
    sd = synth.make_state_dict('vanilla', 6, fitted=True)
    torch.save({'module.' + k: t for k, t in sd.items()}, os.path.join(models, 'p2s_shim_model.pth'))
    torch.save(synth.make_train_opt('vanilla'), os.path.join(models, 'p2s_shim_params.pth'))
    return names


def test_full_eval_body_through_the_source_shim(tmp_path):
    sys.modules.pop('source', None)
    from source import points_to_surf_eval
    from source.base import evaluation
    from source import sdf
    import source
    assert os.path.dirname(os.path.abspath(source.__file__)).startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    root, models, out = str(tmp_path / 'data'), str(tmp_path / 'models'), str(tmp_path / 'results')
    names = _dataset(root, models)
    opt = points_to_surf_eval.parse_arguments([
        '--indir', root, '--outdir', out, '--modeldir', models, '--models', 'p2s_shim', '--dataset', 'testset.txt',
        '--query_grid_resolution', '32', '--epsilon', '3', '--certainty_threshold', '13', '--sigma', '5', '--batchSize', '501',
        '--workers', '0'])
    # ---- the body of full_eval.full_eval, for one dataset entry
    indir_root = opt.indir
    outdir_root = os.path.join(opt.outdir, opt.models + os.path.splitext(opt.modelpostfix)[0])
    dataset = opt.dataset
    opt.indir = os.path.join(indir_root, os.path.dirname(dataset))
    opt.outdir = os.path.join(outdir_root, os.path.dirname(dataset))
    opt.dataset = os.path.basename(dataset)
    assert os.path.exists(os.path.join(opt.indir, '05_query_dist'))
    opt.reconstruction = False
    points_to_surf_eval.points_to_surf_eval(opt)
    res_dir_eval = os.path.join(opt.outdir, 'eval')
    evaluation.eval_predictions(os.path.join(res_dir_eval, 'eval'), os.path.join(opt.indir, '05_query_dist'),
                                os.path.join(res_dir_eval, 'rme_comp_res.csv'), unsigned=False)
    opt.reconstruction = True
    points_to_surf_eval.points_to_surf_eval(opt)
    res_dir_rec = os.path.join(opt.outdir, 'rec')
    sdf.implicit_surface_to_mesh_directory(os.path.join(res_dir_rec, 'dist_ms'), os.path.join(res_dir_rec, 'query_pts_ms'),
                                           os.path.join(res_dir_rec, 'vol'), os.path.join(res_dir_rec, 'mesh'),
                                           opt.query_grid_resolution, opt.sigma, opt.certainty_threshold, opt.workers)
    csv_file = os.path.join(res_dir_rec, 'hausdorff_dist_pred_rec.csv')
    evaluation.mesh_comparison(new_meshes_dir_abs=os.path.join(res_dir_rec, 'mesh'), ref_meshes_dir_abs=os.path.join(opt.indir, '03_meshes'),
                               num_processes=opt.workers, report_name=csv_file, samples_per_model=10000,
                               dataset_file_abs=os.path.join(opt.indir, opt.dataset))
    # ---- the output tree of SURVEY section 10
    base = os.path.join(out, 'p2s_shim_model')
    for n in names:
        for rel in ('eval/eval/%s.xyz.npy', 'eval/eval/%s.xyz.txt', 'eval/vis/%s.ply', 'rec/eval/%s.xyz.npy', 'rec/vis/%s.ply',
                    'rec/query_pts_ms/%s.xyz.npy', 'rec/dist_ms/%s.xyz.npy', 'rec/query_pts_ms_vis/%s.ply', 'rec/vol/%s.off', 'rec/mesh/%s.ply'):
            assert os.path.isfile(os.path.join(base, rel % n)), rel % n
        assert np.load(os.path.join(base, 'eval/eval/%s.xyz.npy' % n)).shape == (200,)
    rows = open(os.path.join(base, 'eval', 'rme_comp_res.csv')).read().strip().split('\n')
    assert len(rows) >= 3
    rep = open(csv_file).read().strip().split('\n')
    assert rep[0].startswith('in mesh,ref mesh,Hausdorff dist new-ref') and len(rep) == 3
    for line in rep[1:]:
        cols = line.split(',')
        cd = float(cols[5])
        # fitted last layer: the reconstruction of the r = 0.5 sphere stays within a few voxels (2 / 32) of the reference mesh
        assert 0.0 < cd / 20000.0 < 0.1, cd
