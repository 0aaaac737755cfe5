"""CPU check of the hand-written backward sequencing in points2surf_b200.train.TrainStep: with the primitives replaced
by plain torch ops (tests/helpers_train.py:TorchPrims) every parameter gradient, the SGD update and the BatchNorm
running statistics must match autograd over the training oracle (oracle/train_oracle.py, which make_golden.py pins
against the unmodified reference).  The CUDA primitives themselves are tested in tests/test_gpu_train.py."""
import numpy as np
import pytest
import torch

from oracle import train_oracle
from points2surf_b200 import synth
from points2surf_b200.train import TrainStep, compute_loss
from helpers_train import TorchPrims, make_train_batch


def _relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize('variant', ['vanilla', 'max', 'uniform'])
@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_hand_written_backward_matches_autograd(variant, dtype):
    v = synth.VARIANTS[variant]
    P, S, B = 20, 30, 5          # small point counts: the sequencing is size-independent
    sd = synth.make_state_dict(variant, seed=3)
    batch = make_train_batch(B, P, S, seed=1)
    ref = train_oracle.train_iteration(sd, batch, v['use_point_stn'], v['shared_transformer'], lr=0.01, momentum=0.9, dtype=dtype)
    ts = TrainStep(sd, v['use_point_stn'], v['shared_transformer'], points_per_patch=P, sub_sample_size=S, lr=0.01,
                   momentum=0.9, device='cpu', prims=TorchPrims(), dtype=dtype)
    losses = ts.step({k: t.clone().to(dtype) for k, t in batch.items()})
    # float64: rounding-free, the sequencing must agree to ~1e-9.  float32: a batch of 5 makes train-mode BatchNorm
    # ill-conditioned, two summation orders differ by ~1e-3 relative after 40 layers.
    tol = 1e-8 if dtype == torch.float64 else 2e-2
    for got, want in zip(losses, ref['losses']):
        assert abs(float(got) - want) <= tol * abs(want)
    assert _relerr(ts.last_logits, ref['logits']) < tol
    grads = ts.named_gradients()
    assert set(grads) == set(ref['grads'])
    scale = max(float(r.abs().max()) for r in ref['grads'].values())
    bad = []
    for name, g in grads.items():
        r = ref['grads'][name]
        assert g.shape == r.shape, name
        # conv / fc biases in front of a BatchNorm have an exactly-zero gradient up to rounding: floor on the scale
        err = float((g - r).abs().max()) / (float(r.abs().max()) + (1e-6 if dtype == torch.float64 else 1e-2) * scale)
        if err > tol:
            bad.append((name, err))
    assert not bad, bad
    new = ts.state_dict()
    for name, r in ref['new_state'].items():
        assert new[name].shape == r.shape, name
        assert float((new[name] - r).abs().max()) <= tol * (1e-3 + float(r.abs().max())), name
    assert int(new['bn2.num_batches_tracked']) == 101


def test_second_step_uses_momentum():
    v = synth.VARIANTS['max']
    P, S, B = 16, 24, 4
    sd = synth.make_state_dict('max', seed=5)
    b1, b2 = make_train_batch(B, P, S, seed=2), make_train_batch(B, P, S, seed=3)
    r1 = train_oracle.train_iteration(sd, b1, 0, 0, dtype=torch.float64)
    sd2 = dict(sd)
    sd2.update(r1['new_state'])
    r2 = train_oracle.train_iteration(sd2, b2, 0, 0, mom_bufs=r1['mom_bufs'], dtype=torch.float64)
    ts = TrainStep(sd, 0, 0, points_per_patch=P, sub_sample_size=S, device='cpu', prims=TorchPrims(), dtype=torch.float64)
    ts.step({k: t.double() for k, t in b1.items()})
    ts.step({k: t.double() for k, t in b2.items()})
    new = ts.state_dict()
    for name, r in r2['new_state'].items():
        assert float((new[name] - r).abs().max()) <= 1e-9 * (1.0 + float(r.abs().max())), name


def test_compute_loss_rejects_unsupported_outputs():
    with pytest.raises(ValueError):
        compute_loss(torch.zeros(2, 2), {}, ['imp_surf'], {}, False, prims=TorchPrims())
    with pytest.raises(ValueError):
        compute_loss(torch.zeros(2, 2), {}, ['imp_surf_magnitude'], {}, False, prims=TorchPrims())
