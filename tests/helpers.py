"""Shared test helpers (CPU side)."""
import os

import numpy as np

from points2surf_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


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


def check_train_digest(variant, grads, new_state, losses, logits, rtol, floor):
    """Compare one training iteration with the reference's digest.  `grads` / `new_state`: name -> torch tensor (CPU).
    Tolerance per tensor: rtol * max|reference gradient| + floor * max over all gradients."""
    g = load_golden('train_%s.npz' % variant)
    assert abs(float(losses[0]) - g['losses'][0]) <= rtol * g['losses'][0] + 1e-6
    assert abs(float(losses[1]) - g['losses'][1]) <= rtol * g['losses'][1] + 1e-6
    assert np.abs(np.asarray(logits) - g['logits']).max() <= rtol * np.abs(g['logits']).max() + 1e-5
    names = [str(n) for n in g['names']]
    assert sorted(grads) == names
    gscale = float(g['grad_max'].max())
    worst = 0.0
    for i, name in enumerate(names):
        t = grads[name].reshape(-1).double().numpy()
        idx = train_digest_indices(name, t.size)
        tol = rtol * float(g['grad_max'][i]) + floor * gscale
        err = np.abs(t[idx] - g['grad_samples'][i][:idx.size]).max()
        assert err <= tol, (name, err, tol)
        nerr = abs(float(np.linalg.norm(t)) - float(g['grad_norm'][i]))
        assert nerr <= rtol * float(g['grad_norm'][i]) + floor * gscale * np.sqrt(t.size), (name, 'norm', nerr)
        worst = max(worst, err / (float(g['grad_max'][i]) + floor * gscale))
        p = new_state[name].reshape(-1).double().numpy()
        perr = np.abs(p[idx] - g['new_samples'][i][:idx.size]).max()
        assert perr <= 0.01 * tol + 1e-6, (name, 'updated parameter', perr)     # lr = 0.01
    for i, name in enumerate(str(n) for n in g['buffer_names']):
        b = new_state[name].reshape(-1).double().numpy()
        idx = train_digest_indices(name, b.size)
        ref = g['buffer_samples'][i][:idx.size]
        assert np.abs(b[idx] - ref).max() <= rtol * (np.abs(ref).max() + 1e-3), (name, 'running statistic')
    return worst
