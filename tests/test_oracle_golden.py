"""CPU: the oracle restatement against the golden vectors generated from the real reference
(tests/golden/make_golden.py).  Keeps the oracle pinned on boxes where /root/reference is absent."""
import numpy as np
import pytest

from oracle import p2s_oracle as orc
from points2surf_b200 import synth
from helpers import load_golden, golden_model_case


@pytest.mark.parametrize('variant', ['vanilla', 'max', 'uniform'])
def test_model_forward_matches_reference(variant):
    sd, inp, g = golden_model_case(variant)
    v = synth.VARIANTS[variant]
    out, aux = orc.model_forward(sd, inp['patch_pts_ps'], inp['pts_sub_sample_ms'], inp['imp_surf_query_point_ms'],
                                 v['use_point_stn'], v['shared_transformer'], return_aux=True)
    np.testing.assert_allclose(out, g['logits'], rtol=0, atol=2e-4)   # fp32 CPU, different BLAS threads/blocking
    np.testing.assert_allclose(aux['feat_global_max'][:2], g['feat_global_max'], rtol=1e-4, atol=1e-4)
    assert ((out[:, 1] >= 0) == (g['logits'][:, 1] >= 0)).all()
    np.testing.assert_allclose(orc.post_process(g['logits'], g['radius']), g['sdf'], rtol=1e-6, atol=1e-7)
    assert 0 < (g['logits'][:, 1] >= 0).sum() < 8   # mixed sign classes: the test is not vacuous


def test_forward_does_not_mutate_inputs():
    sd, inp, g = golden_model_case('vanilla')
    before = inp['pts_sub_sample_ms'].copy()
    orc.model_forward(sd, inp['patch_pts_ps'], inp['pts_sub_sample_ms'], inp['imp_surf_query_point_ms'], 1, 1)
    assert np.array_equal(before, inp['pts_sub_sample_ms'])


def test_query_grid_golden():
    g = load_golden('grid.npz')
    for name, kind, n, res, eps in (('sphere64e3', 'sphere', 10000, 64, 3), ('torus48e4', 'torus', 4000, 48, 4),
                                    ('box40e2', 'box', 3000, 40, 2), ('sphere32e5', 'sphere', 2000, 32, 5)):
        cloud = synth.make_cloud(kind, n, seed=2)
        idx = orc.query_grid_indices(cloud, res, eps)
        assert len(idx) == int(g[name + '_count'])
        assert np.array_equal(idx.astype(np.int64).sum(axis=0), g[name + '_idxsum'])
        assert np.array_equal(idx, orc.query_grid_indices_shifts(cloud, res, eps))
        q = orc.query_grid(cloud, res, eps)
        assert np.array_equal(q[:4], g[name + '_first']) and np.array_equal(q[-4:], g[name + '_last'])


def test_assembly_golden():
    g = load_golden('assembly.npz')
    cloud, res, eps, k, S, seed = g['cloud'], int(g['res']), int(g['eps']), int(g['k']), int(g['S']), int(g['seed'])
    idx = orc.query_grid_indices(cloud, res, eps)
    assert np.array_equal(idx, g['query_idx'].astype(np.int64))
    qpts = orc.volume_space_to_model_space(idx, res).astype(np.float32)
    kd = orc.make_kdtree(cloud)
    for uniform, tag in ((0, 'wgt'), (1, 'uni')):
        rng = np.random.RandomState(seed)
        for qi in range(6):
            item = orc.assemble_query(cloud, kd, qpts[qi], k, S, rng, bool(uniform))
            assert np.array_equal(item['patch_pts_ps'], g['patch_ps'][qi])
            assert item['patch_radius_ms'] == g['radius'][qi]
            assert np.array_equal(item['sub_sample_ids'], g['sub_ids_' + tag][qi])
            bid, _ = orc.knn_bruteforce(cloud, qpts[qi], k)
            assert set(bid.tolist()) == set(g['patch_ids'][qi].tolist())


def test_sign_propagation_golden():
    g = load_golden('volume.npz')
    for name in ('sphere', 'noisy'):
        res = int(g[name + '_res'])
        vol = orc.sdf_to_volume(g[name + '_dist'], g[name + '_qpts'], res, 5, 13)
        ref = np.clip(g[name + '_vol'].astype(np.float64), -1.0, 1.0)
        assert np.array_equal(vol, ref)
        v2 = np.zeros((res,) * 3)
        v2 = orc.add_samples_to_volume(v2, g[name + '_qpts'], g[name + '_dist'])
        v2, _ = orc.propagate_sign(v2, 3, 5)
        assert np.array_equal(v2.astype(np.float32), g[name + '_vol_s3t5'])


def test_all_zero_distances_are_skipped():
    q = orc.query_grid(synth.make_cloud('sphere', 500, seed=1), 16, 3)
    assert orc.sdf_to_volume(np.zeros(len(q), np.float32), q, 16, 5, 13) is None


def test_marching_cubes_oracle_sphere_is_closed_and_accurate():
    from oracle import mc_oracle as mc
    R = 24
    g = (np.arange(R) + 0.5) / R * 2 - 1
    X, Y, Z = np.meshgrid(g, g, g, indexing='ij')
    vol = (0.6 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2)).astype(np.float32)
    v, f = mc.marching_cubes(vol)
    assert mc.mesh_is_closed(f) and len(v) - 3 * len(f) // 2 + len(f) == 2       # Euler characteristic of a sphere
    r = np.linalg.norm(v, axis=1)
    assert abs(r - 0.6).max() < 0.5 * (2.0 / R) ** 2 + 1e-3                      # linear interpolation error
    a, b, c = (v[f[:, i]].astype(np.float64) for i in range(3))
    assert np.einsum('ij,ij->i', a, np.cross(b, c)).sum() > 0                    # outward orientation
    # Chamfer (reference definition) between the mesh and the analytic sphere, 10k samples each side
    rng = np.random.RandomState(0)
    s_mesh = orc.sample_mesh_surface(v, f, 10000, rng)
    d = rng.standard_normal((10000, 3))
    s_ref = 0.6 * d / np.linalg.norm(d, axis=1, keepdims=True)
    assert orc.chamfer(s_mesh, s_ref) / 20000 < 0.02


def test_marching_cubes_oracle_handles_zeros_and_noise():
    from oracle import mc_oracle as mc
    rng = np.random.RandomState(0)
    vol = rng.standard_normal((12, 12, 12)).astype(np.float32)
    vol[[0, -1], :, :] = -1; vol[:, [0, -1], :] = -1; vol[:, :, [0, -1]] = -1
    vol[3, 3, 3] = 0; vol[5, 5, 5] = 0
    v, f = mc.marching_cubes(vol)
    assert mc.mesh_is_closed(f) and np.isfinite(v).all()
    v2, f2 = mc.marching_cubes(np.full((8, 8, 8), -1.0, np.float32))
    assert len(v2) == 0 and len(f2) == 0


@pytest.mark.parametrize('variant', ['vanilla', 'max', 'uniform'])
def test_train_oracle_reproduces_reference_iteration(variant):
    """oracle/train_oracle.py against the digest of one training iteration of the unmodified reference
    (tests/golden/train_*.npz: losses, logits, every gradient, updated parameters, BatchNorm running statistics)."""
    from oracle import train_oracle
    from helpers import TRAIN_SEEDS, check_train_digest, train_fixture_batch
    v = synth.VARIANTS[variant]
    sd = synth.make_state_dict(variant, seed=TRAIN_SEEDS[variant])
    out = train_oracle.train_iteration(sd, train_fixture_batch(variant), v['use_point_stn'], v['shared_transformer'],
                                       lr=0.01, momentum=0.9)
    check_train_digest(variant, out['grads'], out['new_state'], out['losses'], out['logits'].numpy(), tol=2e-2)


def test_ball_patch_oracle_matches_reference_golden():
    """orc.ball_patch against tests/golden/ball.npz (made by the unmodified reference dataset with patch_radius > 0)."""
    g = load_golden('ball.npz')
    cloud = load_golden('assembly.npz')['cloud']
    kd = orc.make_kdtree(cloud)
    qall = orc.query_grid(cloud, int(g['res']), int(g['eps']))
    for tag in ('small', 'large'):
        rng = np.random.RandomState(int(g['seed']))
        radius, P = float(g[tag + '_radius']), int(g[tag + '_P'])
        for j, qi in enumerate(g[tag + '_query_sel']):
            ids, patch, cnt = orc.ball_patch(cloud, kd, qall[qi], radius, P, rng)
            assert cnt == int(g[tag + '_counts'][j])
            assert np.array_equal(patch, g[tag + '_patch_ps'][j])


def test_marching_cubes_table_oracle_matches_table_free_restatement():
    """oracle/mc_oracle.py (table from tools/gen_mc_tables.py, shared with the CUDA kernel) against oracle/mc_topo.py (no
    table: polygons traced on the cell values) on every corner-sign configuration with random magnitudes, and on noise."""
    from oracle import mc_oracle as mc, mc_topo
    rng = np.random.RandomState(0)
    noise = rng.standard_normal((17, 17, 17)).astype(np.float32)
    noise[[0, -1]] = -1; noise[:, [0, -1]] = -1; noise[:, :, [0, -1]] = -1
    for vol in (mc_topo.all_cases_volume(0), mc_topo.all_cases_volume(3), noise):
        v0, f0 = mc.marching_cubes(vol, 0.0)
        v1, f1, st = mc_topo.marching_cubes(vol, 0.0, 'asymptotic', return_stats=True)
        assert np.array_equal(v0, v1) and np.array_equal(f0, f1)
        assert mc.mesh_is_closed(f0) and st['ambiguous_face_cells'] > 100
        # the classic reading of ambiguous faces is a different surface on these volumes (documented in DESIGN.md)
        _, f2 = mc_topo.marching_cubes(vol, 0.0, 'separate_positive')
        assert mc.mesh_is_closed(f2) and not np.array_equal(mc_topo.triangle_set(f1), mc_topo.triangle_set(f2))
    # exact zeros are not positive: an isolated zero voxel in a negative region produces no surface
    z = np.full((6, 6, 6), -1.0, np.float32); z[3, 3, 3] = 0.0
    assert mc_topo.marching_cubes(z, 0.0)[1].shape[0] == 0 and mc.marching_cubes(z, 0.0)[1].shape[0] == 0
