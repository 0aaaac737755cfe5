"""Host logic of points2surf_b200.points_to_surf_train on the CPU: argument parser and parameter initialisation against
the reference (tests/golden/train_init.npz), then two tiny epochs end to end with the test primitives and an oracle-based
batch assembler standing in for the GPU kernels (both are tested on their own in the GPU suite)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import p2s_oracle as orc
from points2surf_b200 import points_to_surf_train as p2s_train, synth
from helpers import load_golden
from helpers_train import TorchPrims


def test_parser_defaults_and_initialisation_match_reference():
    g = load_golden('train_init.npz')
    assert json.loads(str(g['parser_defaults_json'])) == json.loads(json.dumps(vars(p2s_train.parse_arguments([])), sort_keys=True))
    for variant in ('vanilla', 'max', 'uniform'):
        v = synth.VARIANTS[variant]
        torch.manual_seed(3627473)
        sd = p2s_train.initial_state_dict(v['use_point_stn'], v['shared_transformer'], 1024, 2)
        names = [str(n) for n in g[variant + '_names']]
        assert [k for k in sd if sd[k].is_floating_point()] == names
        sums = np.array([float(sd[k].double().sum()) for k in names])
        first = np.array([float(sd[k].reshape(-1)[0]) for k in names])
        assert np.array_equal(sums, g[variant + '_sums']) and np.array_equal(first, g[variant + '_first'])


class OracleAssembler:
    """CPU stand-in for GpuAssembler built from the oracle's per-query functions."""

    def __init__(self, P, S, uniform, seed):
        self.P, self.S, self.uniform = P, S, bool(uniform)
        self.rng = np.random.RandomState(seed)
        self.trees = {}

    def assemble(self, key, pts, query_pts):
        if key not in self.trees:
            self.trees[key] = orc.make_kdtree(pts)
        items = [orc.assemble_query(pts, self.trees[key], q, self.P, self.S, self.rng, self.uniform) for q in query_pts]
        patch = torch.from_numpy(np.stack([it['patch_pts_ps'] for it in items]).astype(np.float32))
        radius = torch.from_numpy(np.array([it['patch_radius_ms'] for it in items], dtype=np.float32))
        sub = torch.from_numpy(np.stack([it['pts_sub_sample_ms'] for it in items]).astype(np.float32))
        return patch, radius, sub, torch.from_numpy(np.ascontiguousarray(query_pts, dtype=np.float32))


def _make_dataset(root, names, n_pts=400, n_query=12):
    for sub in ('04_pts', '05_query_pts', '05_query_dist'):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    rng = np.random.RandomState(0)
    for i, name in enumerate(names):
        cloud = synth.make_cloud(['sphere', 'torus', 'box'][i % 3], n_pts, seed=i)
        q = (cloud[rng.choice(n_pts, n_query, replace=False)] + rng.normal(0, 0.02, (n_query, 3))).astype(np.float32)
        d = (np.linalg.norm(q, axis=1) - 0.5).astype(np.float32)       # signed distance to the r = 0.5 sphere as stand-in GT
        np.save(os.path.join(root, '04_pts', name + '.xyz.npy'), cloud)
        np.save(os.path.join(root, '05_query_pts', name + '.ply.npy'), q)
        np.save(os.path.join(root, '05_query_dist', name + '.ply.npy'), d)
    with open(os.path.join(root, 'trainset.txt'), 'w') as f:
        f.write('\n'.join(names[:2]) + '\n')
    with open(os.path.join(root, 'testset.txt'), 'w') as f:
        f.write(names[2] + '\n')


@pytest.mark.parametrize('order', ['random', 'random_shape_consecutive'])
def test_two_epochs_end_to_end(tmp_path, capsys, order):
    root = str(tmp_path / 'data')
    _make_dataset(root, ['s0', 's1', 's2'])
    opt = p2s_train.parse_arguments([
        '--name', 'test', '--indir', root, '--outdir', str(tmp_path / 'models'), '--logdir', str(tmp_path / 'logs'),
        '--nepoch', '2', '--batchSize', '5', '--patches_per_shape', '8', '--points_per_patch', '16', '--sub_sample_size', '24',
        '--patch_radius', '0.0', '--lr', '0.01', '--scheduler_steps', '1', '--shared_transformer', '1', '--training_order', order,
        '--outputs', 'imp_surf_magnitude', 'imp_surf_sign', 'patch_pts_ids', 'p_index', '--save_interval', '1'])
    hist = p2s_train.points_to_surf_train(opt, prims=TorchPrims(), assembler=OracleAssembler(16, 24, 0, 1), device=torch.device('cpu'))
    out = capsys.readouterr().out
    assert 'Training set: 16 patches (in 4 batches) | Test set: 8 patches (in 2 batches)' in out
    assert 'LR changed from 0.01 to 0.001 in epoch 0' in out
    assert '[test 0: 0/3]' in out and 'train' in out and 'test' in out
    train_rows = [h for h in hist if h[0] == 'train']
    test_rows = [h for h in hist if h[0] == 'test']
    assert len(train_rows) == 8 and len(test_rows) == 4            # 2 epochs x (4 train + 2 interleaved test batches)
    assert all(np.isfinite(l).all() for _, _, _, l in hist)
    # the test batches are interleaved at the same fraction of the epoch as in the reference loop
    kinds = [h[0] for h in hist if h[1] == 0]
    assert kinds == ['train', 'test', 'train', 'test', 'train', 'train']   # test fraction <= train fraction (train.py:476)
    models = tmp_path / 'models'
    for f in ('test_params.pth', 'test_model.pth', 'test_model_0.pth', 'test_model_1.pth', 'test_description.txt'):
        assert (models / f).exists(), f
    saved_opt = torch.load(models / 'test_params.pth', weights_only=False)
    assert saved_opt.train_shapes == ['s0', 's1'] and saved_opt.test_shapes == ['s2'] and saved_opt.points_per_patch == 16
    sd = torch.load(models / 'test_model.pth')
    assert all(k.startswith('module.') for k in sd)
    ref_keys = ['module.' + k for k in synth.make_state_dict('vanilla', 0).keys()]
    assert sorted(sd.keys()) == sorted(ref_keys)
    assert int(sd['module.bn2.num_batches_tracked']) == 8          # 8 training iterations, none from the eval-mode test batches
    # the checkpoint is a valid model for the inference oracle (the hand-over the reference's eval relies on)
    inp = synth.make_model_inputs(2, points_per_patch=16, sub_sample_size=24, seed=1)
    logits = orc.model_forward({k[7:]: v for k, v in sd.items()}, inp['patch_pts_ps'], inp['pts_sub_sample_ms'],
                               inp['imp_surf_query_point_ms'], True, True)
    assert np.isfinite(logits).all()


def test_unsupported_training_options_raise():
    base = ['--outputs', 'imp_surf_magnitude', 'imp_surf_sign', '--patch_radius', '0.0']
    p2s_train._check_supported(p2s_train.parse_arguments(base))
    p2s_train._check_supported(p2s_train.parse_arguments(base + ['--patch_radius', '0.1']))     # radius ablations: ball-query patches
    # train_p2s_vanilla_uniform_subsample.sh: the fixed sub-sample is supported together with the uniform one only
    p2s_train._check_supported(p2s_train.parse_arguments(base + ['--uniform_subsample', '1', '--fixed_subsample', '1']))
    for extra in (['--sym_op', 'sum'], ['--single_transformer', '1'], ['--training_order', 'bogus'],
                  ['--fixed_subsample', '1'], ['--outputs', 'imp_surf'], ['--outputs', 'normals']):
        with pytest.raises(ValueError):
            p2s_train._check_supported(p2s_train.parse_arguments(base + extra))
    with pytest.raises(ValueError):      # the reference's default outputs include the regression output
        p2s_train._check_supported(p2s_train.parse_arguments(['--patch_radius', '0.0']))
