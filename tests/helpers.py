"""Shared test helpers (CPU side)."""
import os

import numpy as np

from points2surf_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def ids_checksum(ids):
    """Order-sensitive 64-bit checksum of an int id matrix [Q,S] (shared with the tests)."""
    a = np.asarray(ids, dtype=np.uint64)
    w = (np.arange(a.shape[1], dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))[None, :]
    rows = (a * w).sum(axis=1, dtype=np.uint64)
    k = np.arange(a.shape[0], dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(7)
    return int(np.bitwise_xor.reduce(rows * k))


def golden_model_case(variant):
    """-> (state_dict (torch tensors, calibrated fc4 bias), inputs dict (numpy), golden npz)."""
    import torch
    g = load_golden('model_%s.npz' % variant)
    sd = synth.make_state_dict(variant, int(g['seed']))
    sd['fc4.bias'] = torch.from_numpy(g['fc4_bias'].copy())
    inp = synth.make_model_inputs(8, seed=int(g['input_seed']))
    chk = sum(float(np.abs(a).sum()) for a in inp.values())
    assert abs(chk - float(g['input_checksum'])) < 1e-6 * chk, 'synthetic input generator drifted'
    return sd, inp, g


def calibrated_state_dict(variant, seed, inputs=None):
    """Rand-init checkpoint whose output bias is centred on a calibration batch (oracle forward),
    so that sign classes are mixed."""
    import torch
    from oracle import p2s_oracle as orc
    v = synth.VARIANTS[variant]
    sd = synth.make_state_dict(variant, seed)
    inp = inputs if inputs is not None else synth.make_model_inputs(8, seed=1000 + seed)
    raw = orc.model_forward(sd, inp['patch_pts_ps'], inp['pts_sub_sample_ms'], inp['imp_surf_query_point_ms'],
                            v['use_point_stn'], v['shared_transformer'])
    sd['fc4.bias'] = sd['fc4.bias'] - torch.from_numpy(np.median(raw, axis=0).astype(np.float32))
    return sd


TRAIN_SEEDS = {'vanilla': 21, 'max': 22, 'uniform': 23}


def train_digest_indices(name, numel, count=48):
    """Sample positions of the gradient digests in tests/golden/train_*.npz (same rule as make_golden.py)."""
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return np.random.RandomState(h).randint(0, numel, size=min(count, numel))


def train_fixture_batch(variant):
    """The seeded batch of tests/golden/train_<variant>.npz (regenerated, checked against the stored checksum)."""
    from helpers_train import make_train_batch
    g = load_golden('train_%s.npz' % variant)
    batch = make_train_batch(int(g['batch']), 300, 1000, seed=TRAIN_SEEDS[variant])
    chk = sum(float(t.double().abs().sum()) for t in batch.values())
    assert abs(chk - float(g['input_checksum'])) < 1e-9 * chk, 'synthetic batch generator drifted'
    return batch


def check_train_digest(variant, grads, new_state, losses, logits, tol):
    """Compare one training iteration with the digest of the reference's iteration.  `grads` / `new_state`:
    name -> torch tensor (CPU).  Gradients are compared in the L2 sense (norm of every tensor; all sampled entries
    together, each tensor scaled by its largest reference entry): arg-max / ReLU decisions flip under fp32 rounding
    and move single entries by O(1/batch).  Returns (worst norm error, sample-vector relative L2 error)."""
    g = load_golden('train_%s.npz' % variant)
    assert abs(float(losses[0]) - g['losses'][0]) <= 2e-3 * g['losses'][0] + 1e-6
    assert abs(float(losses[1]) - g['losses'][1]) <= 2e-3 * g['losses'][1] + 1e-6
    assert np.abs(np.asarray(logits) - g['logits']).max() <= 5e-3 * np.abs(g['logits']).max() + 1e-5
    names = [str(n) for n in g['names']]
    assert sorted(grads) == names
    nscale = float(g['grad_norm'].max())
    bad, worst_norm = [], 0.0
    num = den = pnum = 0.0
    for i, name in enumerate(names):
        t = grads[name].reshape(-1).double().numpy()
        idx = train_digest_indices(name, t.size)
        nerr = abs(float(np.linalg.norm(t)) - float(g['grad_norm'][i])) / (float(g['grad_norm'][i]) + 1e-3 * nscale)
        worst_norm = max(worst_norm, nerr)
        if nerr > tol:
            bad.append((name, 'norm', round(nerr, 4)))
        scale = float(g['grad_max'][i]) + 1e-3 * float(g['grad_max'].max())
        ref = g['grad_samples'][i][:idx.size]
        num += float((((t[idx] - ref) / scale) ** 2).sum())
        den += float(((ref / scale) ** 2).sum())
        p = new_state[name].reshape(-1).double().numpy()
        pnum += float((((p[idx] - g['new_samples'][i][:idx.size]) / (0.01 * scale)) ** 2).sum())   # lr = 0.01
    assert not bad, 'gradient norm mismatches: %s' % bad
    sample_err, param_err = (num / den) ** 0.5, (pnum / den) ** 0.5
    assert sample_err <= tol, ('sampled gradient entries', sample_err)
    assert param_err <= tol, ('sampled updated parameters', param_err)
    for i, name in enumerate(str(n) for n in g['buffer_names']):
        b = new_state[name].reshape(-1).double().numpy()
        idx = train_digest_indices(name, b.size)
        ref = g['buffer_samples'][i][:idx.size]
        assert np.abs(b[idx] - ref).max() <= 2e-3 * (np.abs(ref).max() + 1e-3), (name, 'running statistic')
    return worst_norm, sample_err


def compare_gradients_l2(grads, ref_grads, tol_tensor, tol_global):
    """Full-tensor comparison: per-tensor and global relative L2 error of `grads` against `ref_grads`."""
    nscale = max(float(r.double().norm()) for r in ref_grads.values())
    num = den = 0.0
    bad, worst = [], 0.0
    for name, r in ref_grads.items():
        d = grads[name].double().reshape(-1) - r.double().reshape(-1)
        e = float(d.norm()) / (float(r.double().norm()) + 1e-3 * nscale)
        worst = max(worst, e)
        if e > tol_tensor:
            bad.append((name, round(e, 4)))
        num += float((d * d).sum())
        den += float((r.double() ** 2).sum())
    assert not bad, 'per-tensor relative L2 errors above %g: %s' % (tol_tensor, bad)
    glob = (num / den) ** 0.5
    assert glob <= tol_global, glob
    return worst, glob
