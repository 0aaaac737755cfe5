"""Diagnostic: how far are three fp32 implementations of one training iteration from the f64 CPU oracle?
  (a) CPU fp32 oracle (torch autograd)   (b) TrainStep over torch CUDA ops (tests/helpers_train.TorchPrims on the GPU)
  (c) TrainStep over the p2s_op_* CUDA primitives.  Prints per-tensor worst and global relative L2 errors."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import train_oracle  # noqa: E402
from points2surf_b200 import synth  # noqa: E402
from points2surf_b200.train import TrainStep  # noqa: E402
from helpers_train import TorchPrims, make_train_batch  # noqa: E402


def err(grads, ref):
    nscale = max(float(r.double().norm()) for r in ref.values())
    num = den = 0.0
    worst, wk = 0.0, None
    per = {}
    for k, r in ref.items():
        d = grads[k].double().cpu().reshape(-1) - r.double().reshape(-1)
        e = float(d.norm()) / (float(r.double().norm()) + 1e-3 * nscale)
        per[k] = e
        if e > worst:
            worst, wk = e, k
        num += float((d * d).sum())
        den += float((r.double() ** 2).sum())
    return worst, wk, (num / den) ** 0.5, per


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for variant in ('vanilla', 'uniform', 'max'):
        v = synth.VARIANTS[variant]
        sd = synth.make_state_dict(variant, seed=21)
        batch = make_train_batch(B, seed=21)
        truth = train_oracle.train_iteration(sd, batch, v['use_point_stn'], v['shared_transformer'], dtype=torch.float64)['grads']
        cpu32 = train_oracle.train_iteration(sd, batch, v['use_point_stn'], v['shared_transformer'])['grads']
        out = {'cpu fp32 autograd': cpu32}
        cb = {k: t.cuda() for k, t in batch.items()}
        csd = {k: t.cuda() for k, t in sd.items()}
        for name, prims in (('torch CUDA ops ', TorchPrims()), ('p2s_op_* kernels', None)):
            ts = TrainStep(csd, v['use_point_stn'], v['shared_transformer'], prims=prims)
            ts.step(cb)
            out[name] = ts.named_gradients()
        pers = {}
        for name, g in out.items():
            w, wk, glob, per = err(g, truth)
            pers[name] = per
            print('%-8s %-18s vs f64: worst tensor %.4f (%s), global %.4f' % (variant, name, w, wk, glob))
        top = sorted(pers['p2s_op_* kernels'].items(), key=lambda kv: -kv[1])[:6]
        for k, e in top:
            print('      %-34s p2s %.4f   torch-cuda %.4f   cpu %.4f' % (k, e, pers['torch CUDA ops '][k], pers['cpu fp32 autograd'][k]))


if __name__ == '__main__':
    main()
