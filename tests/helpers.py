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
