"""GPU tests of the mesh acceptance metric: area-weighted sampling, nearest-neighbour
distances, Chamfer / Hausdorff and the `mesh_comparison` report, through the C ABI, against the oracle
(oracle/p2s_oracle.py: chamfer, hausdorff -- scipy.spatial.cKDTree, the library the reference calls)."""
import os

import numpy as np
import pytest
import scipy.spatial as spatial
import torch

from oracle import p2s_oracle as orc
from points2surf_b200 import ops, synth, mesh_io, evaluation as ev

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize('na,nb', [(1, 1), (37, 5), (3000, 5000), (10000, 10000), (20000, 1500)])
def test_nn_distance_matches_ckdtree(na, nb):
    rng = np.random.RandomState(na + nb)
    a = rng.uniform(-1, 1, (na, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (nb, 3)).astype(np.float32)
    dist, idx = ops.nn_distance(cu(a), cu(b))
    dist, idx = dist.cpu().numpy(), idx.cpu().numpy()
    d_ref, i_ref = spatial.cKDTree(b).query(a, 1)
    # fp32 differences / squares vs the kd-tree's f64: 1e-6 absolute on distances of O(0.01 .. 1)
    np.testing.assert_allclose(dist, d_ref, atol=1e-6, rtol=1e-6)
    # the index may differ only where two targets are equidistant to fp32 precision
    bad = idx != i_ref
    if bad.any():
        alt = np.linalg.norm(a[bad].astype(np.float64) - b[idx[bad]].astype(np.float64), axis=1)
        np.testing.assert_allclose(alt, d_ref[bad], atol=1e-6)


def test_nn_distance_ties_pick_lowest_index():
    a = np.zeros((3, 3), np.float32)
    b = np.tile(np.array([[1.0, 0, 0]], np.float32), (5000, 1))   # all targets equidistant, several slabs
    dist, idx = ops.nn_distance(cu(a), cu(b))
    assert (idx.cpu().numpy() == 0).all() and np.allclose(dist.cpu().numpy(), 1.0)


def test_mesh_sample_is_area_weighted_and_on_surface():
    # two triangles with area ratio 1:8 in different planes, plus a degenerate face that must never be hit
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [4, 0, 1], [0, 4, 1], [2, 2, 2]], np.float32)
    faces = np.array([[0, 1, 2], [6, 6, 6], [3, 4, 5]], np.int32)
    n = 200000
    s, fid = ops.mesh_sample(cu(verts), cu(faces), n, seed=5, return_face_ids=True)
    s, fid = s.cpu().numpy(), fid.cpu().numpy()
    assert not (fid == 1).any()
    frac = (fid == 2).mean()
    assert abs(frac - 16.0 / 17.0) < 4 * np.sqrt((16 / 17) * (1 / 17) / n)
    # every sample lies inside its triangle
    for f in (0, 2):
        p = s[fid == f]
        a, b, c = verts[faces[f]]
        m = np.stack([b - a, c - a], axis=1)
        uv, res, _, _ = np.linalg.lstsq(m, (p - a).T, rcond=None)
        assert np.abs(m @ uv - (p - a).T).max() < 1e-5
        assert uv.min() > -1e-6 and (uv.sum(0)).max() < 1 + 1e-6
        # uniform inside the triangle: mean barycentric coordinates 1/3
        assert np.abs(uv.mean(1) - 1 / 3).max() < 0.01
    # deterministic in the seed, different across seeds
    s2 = ops.mesh_sample(cu(verts), cu(faces), n, seed=5).cpu().numpy()
    s3 = ops.mesh_sample(cu(verts), cu(faces), n, seed=6).cpu().numpy()
    assert np.array_equal(s, s2) and not np.array_equal(s, s3)


def _sphere_mesh(res, radius):
    cloud = synth.make_cloud('sphere', 10000, seed=0, noise=0.0) * (radius / 0.5)
    lin = ops.query_grid(cu(cloud.astype(np.float32)), res, 3)
    q = ops.query_points(lin, res).cpu().numpy()
    d = (radius - np.linalg.norm(q, axis=1)).astype(np.float32)
    vol, _ = ops.sdf_to_volume(lin, cu(d), res, 5, 13.0)
    v, f = ops.marching_cubes(vol, 0.0)
    return v, f


def test_chamfer_hausdorff_match_oracle_on_mesh_samples():
    v1, f1 = _sphere_mesh(48, 0.5)
    v2, f2 = _sphere_mesh(64, 0.45)
    s1 = ops.mesh_sample(v1, f1, 10000, seed=1)
    s2 = ops.mesh_sample(v2, f2, 10000, seed=2)
    r = ops.chamfer_hausdorff(s1, s2)
    a, b = s1.cpu().numpy().astype(np.float64), s2.cpu().numpy().astype(np.float64)
    c_ref = orc.chamfer(a, b)
    h_ref = orc.hausdorff(a, b)
    # sums of 2 x 10^4 fp32 distances accumulated in f64: 1e-5 relative
    assert abs(r['chamfer'] - c_ref) / c_ref < 1e-5, (r, c_ref)
    np.testing.assert_allclose([r['hausdorff_ab'], r['hausdorff_ba'], r['hausdorff']], h_ref, rtol=1e-5)
    # two concentric spheres 0.05 apart: mean nearest-sample distance ~0.05 per direction
    assert 0.045 < r['chamfer'] / 20000 < 0.06
    # device sampler vs the oracle's sampler on the same mesh: same surface integral (Chamfer to the other mesh)
    so = orc.sample_mesh_surface(v1.cpu().numpy(), f1.cpu().numpy(), 10000, np.random.RandomState(3))
    c_o = orc.chamfer(so, b)
    assert abs(c_o - c_ref) / c_ref < 0.02


def test_mesh_comparison_report(tmp_path):
    new_dir, ref_dir = tmp_path / 'rec' / 'mesh', tmp_path / '03_meshes'
    v1, f1 = _sphere_mesh(48, 0.5)
    v2, f2 = _sphere_mesh(64, 0.5)
    mesh_io.write_ply(str(new_dir / 'ball.ply'), v1.cpu().numpy(), f1.cpu().numpy())
    mesh_io.write_ply(str(ref_dir / 'ball.ply'), v2.cpu().numpy(), f2.cpu().numpy())
    mesh_io.write_off(str(new_dir / 'extra.off'), v1.cpu().numpy(), f1.cpu().numpy())     # not in the set file
    mesh_io.write_ply(str(ref_dir / 'lonely.ply'), v2.cpu().numpy(), f2.cpu().numpy())    # no reconstruction
    (tmp_path / 'testset.txt').write_text('ball\nlonely\n')
    rep = tmp_path / 'rep' / 'hausdorff_dist_pred_rec.csv'
    ev.mesh_comparison(str(new_dir), str(ref_dir), 3, str(rep), samples_per_model=10000,
                       dataset_file_abs=str(tmp_path / 'testset.txt'))
    lines = rep.read_text().split('\n')
    assert lines[0] == ('in mesh,ref mesh,Hausdorff dist new-ref,Hausdorff dist ref-new,Hausdorff dist,'
                        'Chamfer dist(-1: no input; -2: no reference)')
    assert len(lines) == 3
    ball = lines[1].split(',')
    assert ball[0].endswith('ball.ply') and ball[1].endswith('ball.ply')
    h_nr, h_rn, h, ch = (float(x) for x in ball[2:])
    assert h == max(h_nr, h_rn) and 0 < h < 0.05           # same sphere meshed at two resolutions
    assert 0 < ch / 20000 < 0.01
    assert lines[2].split(',')[2:] == ['-1', '-1', '-1', '-1'] and 'lonely' in lines[2]
    # single-file helpers keep the reference's return shapes and its -1 sentinels for unreadable meshes
    r = ev._chamfer_distance_single_file(str(new_dir / 'ball.ply'), str(ref_dir / 'ball.ply'), 10000)
    assert r[0].endswith('ball.ply') and abs(r[2] - ch) < 1e-9
    assert ev._chamfer_distance_single_file(str(new_dir / 'nope.ply'), str(ref_dir / 'ball.ply'), 100)[2] == -1.0
    assert ev._hausdorff_distance_single_file(str(new_dir / 'nope.ply'), str(ref_dir / 'ball.ply'), 100)[2:] == (-1.0, -1.0, -1.0)
