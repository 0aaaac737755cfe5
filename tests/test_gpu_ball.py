"""Ball-query patches (the radius ablations): the CUDA kernel through the C ABI against the fixture made by
the UNMODIFIED reference dataset with patch_radius > 0 (tests/golden/ball.npz) and against the oracle
(scipy cKDTree.query_ball_point).  Membership is exact (float64 distances on float32 coordinates, inclusive radius);
patch-space coordinates are bit-exact float32; the random k-subset of an over-full ball follows the same law as the
reference's rng.choice (uniform without replacement) on a different stream."""
import numpy as np
import pytest
import torch

from oracle import p2s_oracle as orc
from points2surf_b200 import synth, ops
from helpers import load_golden, calibrated_state_dict

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rows_sorted(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


@pytest.mark.parametrize('tag', ['small', 'large'])
def test_ball_patch_golden(tag):
    g = load_golden('ball.npz')
    cloud = load_golden('assembly.npz')['cloud']
    res, radius, P = int(g['res']), float(g[tag + '_radius']), int(g[tag + '_P'])
    qpts = orc.query_grid(cloud, res, int(g['eps']))[g[tag + '_query_sel']]
    ids, patch, rad, counts = ops.ball_patch(cu(cloud), cu(qpts), P, radius, seed=5)
    ids, patch, counts = ids.cpu().numpy(), patch.cpu().numpy(), counts.cpu().numpy()
    assert np.array_equal(counts, g[tag + '_counts'])
    assert (rad.cpu().numpy() == np.float32(radius)).all()
    kd = orc.make_kdtree(cloud)
    for i in range(len(qpts)):
        c = int(counts[i])
        ball = np.sort(np.array(kd.query_ball_point(qpts[i], radius), dtype=np.int64))
        assert len(ball) == c
        if c <= P:
            assert np.array_equal(ids[i, :c], ball)                               # ascending id, nothing dropped
            assert (ids[i, c:] == 0).all() and (patch[i, c:] == 0).all()            # padded with the query point itself
            assert np.array_equal(rows_sorted(patch[i]), rows_sorted(g[tag + '_patch_ps'][i]))   # same rows as the reference, bit for bit
        else:
            assert len(set(ids[i].tolist())) == P and set(ids[i].tolist()) <= set(ball.tolist())
            ref = ((cloud[ids[i]] - qpts[i][None, :]) / np.float32(radius)).astype(np.float32)
            assert np.array_equal(patch[i], ref)


def test_ball_patch_subset_is_uniform_and_stream_keyed():
    # one query whose ball holds ~4x k points, drawn under 3000 different stream keys: every ball member is kept with
    # probability k / count (binomial std 0.008 -> 5 sigma = 0.04); the draw depends on (seed, query index) only
    cloud = synth.make_cloud('sphere', 20000, seed=3)
    q = cloud[:1].copy()
    radius, k = 0.08, 32
    kd = orc.make_kdtree(cloud)
    ball = np.array(kd.query_ball_point(q[0], radius))
    assert len(ball) > 3 * k
    trials = 3000
    qq = cu(np.repeat(q, trials, axis=0))
    ids, _, _, counts = ops.ball_patch(cu(cloud), qq, k, radius, seed=9)
    ids = ids.cpu().numpy()
    assert (counts.cpu().numpy() == len(ball)).all()
    assert all(len(set(r.tolist())) == k for r in ids[:50])
    freq = np.bincount(ids.ravel(), minlength=len(cloud))[ball] / trials
    assert np.abs(freq - k / len(ball)).max() < 0.045, np.abs(freq - k / len(ball)).max()
    again, _, _, _ = ops.ball_patch(cu(cloud), qq[100:110], k, radius, seed=9, query_index_base=100)
    assert np.array_equal(again.cpu().numpy(), ids[100:110])


def test_ball_patch_overfull_candidate_buffer():
    # more points in the ball than the 2048-entry candidate buffer: histogram selection path
    rng = np.random.RandomState(0)
    cloud = (rng.standard_normal((30000, 3)) * 0.05).clip(-0.9, 0.9).astype(np.float32)
    q = np.zeros((3, 3), np.float32)
    ids, patch, _, counts = ops.ball_patch(cu(cloud), cu(q), 300, 0.1, seed=1)
    ball = set(orc.make_kdtree(cloud).query_ball_point(q[0], 0.1))
    assert len(ball) > 2048 and int(counts[0]) == len(ball)
    for r in ids.cpu().numpy():
        assert len(set(r.tolist())) == 300 and set(r.tolist()) <= ball
    assert not np.array_equal(ids[0].cpu().numpy(), ids[1].cpu().numpy())       # different stream keys


def test_reconstruct_with_ball_patches_matches_stagewise():
    """Fused pipeline with patch_radius > 0: fixed-radius normalisation, |d| not rescaled (points_to_surf_eval.py:364-368)."""
    variant = 'uniform'                                    # the radius ablations use the per-branch QSTN topology
    v = synth.VARIANTS[variant]
    sd = calibrated_state_dict(variant, 31)
    cloud = synth.make_cloud('sphere', 3000, seed=6)
    res, eps, seed, radius = 16, 3, 77, 0.3
    eng = ops.Engine(sd, v['use_point_stn'], v['shared_transformer'], precision='fp32')
    lin, sdf = eng.reconstruct(cu(cloud), res, eps, 1, seed, patch_radius=radius)
    qd = ops.query_points(lin, res)
    _, patch, _, counts = ops.ball_patch(cu(cloud), qd, 300, radius, seed)
    assert int(counts.max()) > 300 and int(counts.min()) < 300      # both the sub-set and the padding rule are exercised
    sub = ops.gather_points(cu(cloud), ops.subsample(cu(cloud), qd, 1000, True, seed))
    logits = eng.forward(patch, sub, qd)
    ref = ops.sdf_from_logits(logits, None)
    assert torch.allclose(sdf, ref, atol=1e-6)
    # and the oracle on a few queries with the GPU's own ids
    sel = [0, len(qd) // 2, len(qd) - 1]
    lo = orc.model_forward(sd, patch[sel].cpu().numpy(), sub[sel].cpu().numpy(), qd[sel].cpu().numpy(), v['use_point_stn'], v['shared_transformer'])
    want = np.tanh(lo[:, 0]) ** 2 * np.where(lo[:, 1] >= 0, 1.0, -1.0)
    np.testing.assert_allclose(sdf.cpu().numpy()[sel], want, atol=2e-4)
    eng.close()
