"""Test-only stand-in for points2surf_b200.train_ops.CudaPrims built from plain torch ops, so that the hand-written
backward sequencing of points2surf_b200.train.TrainStep (host logic) can be checked on a CPU-only machine against
autograd.  Never imported by the product."""
import numpy as np
import torch

from oracle.p2s_oracle import quat_to_rotmat
from points2surf_b200 import synth


class TorchPrims:
    name = 'torch-test'

    def gemm_nt(self, A, W, bias=None, relu=False):
        out = torch.matmul(A, W.transpose(-1, -2))
        if bias is not None:
            out = out + bias
        return torch.relu(out) if relu else out

    def gemm_tn(self, A, B, out=None):
        r = torch.matmul(A.transpose(-1, -2), B)
        return r if out is None else out.add_(r.reshape(out.shape))

    def bn_maxpool_forward(self, z, B, npts, gamma, beta, relu, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
        y, mean, invstd = self.bn_forward(z, gamma, beta, relu, running_mean, running_var, eps, momentum)
        out, arg = self.maxpool_fwd(y, B, npts)
        return out, arg, mean, invstd

    def bn_maxpool_backward(self, dout, arg, out, z, mean, invstd, gamma, relu, B, npts):
        dy = self.maxpool_bwd(dout, arg, npts)
        y = None
        if relu:   # the mask only matters at the arg-max rows, where y == out
            y = self.maxpool_bwd(out, arg, npts)
        return self.bn_backward(dy, z, y, mean, invstd, gamma)

    def transpose(self, x):
        return x.transpose(-1, -2).contiguous()

    def bn_forward(self, z, gamma, beta, relu, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
        M = z.shape[0]
        zd = z.double()
        mean = zd.mean(0)
        var = (zd * zd).mean(0) - mean * mean
        invstd = 1.0 / torch.sqrt(var + eps)
        y = (gamma * invstd.to(z.dtype)) * (z - mean.to(z.dtype)) + beta
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean.to(z.dtype))
            running_var.mul_(1 - momentum).add_(momentum * (var * M / max(M - 1, 1)).to(z.dtype))
        return (torch.relu(y) if relu else y), mean.to(z.dtype), invstd.to(z.dtype)

    def bn_apply(self, z, mean, invstd, gamma, beta, relu):
        y = (gamma * invstd) * (z - mean) + beta
        return torch.relu(y) if relu else y

    def bn_backward(self, dy, z, y_mask, mean, invstd, gamma):
        g = dy if y_mask is None else dy * (y_mask > 0)
        M = z.shape[0]
        xhat = (z - mean) * invstd
        s1 = g.double().sum(0)
        s2 = (g * xhat).double().sum(0)
        dz = gamma * invstd * (g - (s1 / M).to(z.dtype) - xhat * (s2 / M).to(z.dtype))
        return dz, s2.to(z.dtype), s1.to(z.dtype)

    def col_sum(self, x):
        return x.double().sum(0).to(x.dtype)

    def maxpool_fwd(self, y, B, npts):
        v, a = y.view(B, npts, -1).max(dim=1)
        return v.contiguous(), a.to(torch.int32)

    def maxpool_bwd(self, dout, arg, npts):
        B, C = dout.shape
        dy = torch.zeros(B, npts, C, dtype=dout.dtype, device=dout.device)
        dy.scatter_(1, arg.long().unsqueeze(1), dout.unsqueeze(1))
        return dy.view(B * npts, C)

    def loss(self, pred, target_mag, radius, target_sign, w_mag, w_sign, fixed_radius=False, need_grad=True):
        p = pred.detach().clone().requires_grad_(True)
        t = target_mag if fixed_radius else target_mag / radius
        with torch.enable_grad():
            l0 = torch.nn.functional.mse_loss(torch.tanh(p[:, 0].abs()), torch.tanh(t.abs())) * w_mag
            l1 = torch.nn.functional.binary_cross_entropy_with_logits(p[:, 1], target_sign, reduction='none').mean() * w_sign
            (l0 + l1).backward()
        return torch.stack([l0.detach().double(), l1.detach().double()]), (p.grad if need_grad else None)

    def quat_to_rot(self, q):
        return quat_to_rotmat(q + q.new_tensor([1, 0, 0, 0]))

    def quat_to_rot_bwd(self, q, dR):
        x = q.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            R = quat_to_rotmat(x + x.new_tensor([1, 0, 0, 0]))
            (R.reshape(-1, 9) * dR.reshape(-1, 9)).sum().backward()
        return x.grad

    def add_row_(self, x, v):
        return x.add_(v)

    def center(self, pts, q):
        return pts - q.unsqueeze(1)

    def axpy_(self, y, x, a=1.0):
        return y.add_(x.reshape(y.shape), alpha=a)

    def sgd_(self, param, grad, buf, lr, momentum, first):
        if first:
            buf.copy_(grad)
        else:
            buf.mul_(momentum).add_(grad)
        param.add_(buf, alpha=-lr)


def make_train_batch(B, P=300, S=1000, seed=0):
    """Synthetic batch with the keys of the reference's DataLoader in training mode (data_loader.py:395-421):
    config 4 of SURVEY.md section 8d."""
    rng = np.random.RandomState(seed)
    inp = synth.make_model_inputs(B, points_per_patch=P, sub_sample_size=S, seed=seed)
    out = {k: torch.from_numpy(np.ascontiguousarray(inp[k])) for k in ('patch_pts_ps', 'pts_sub_sample_ms', 'imp_surf_query_point_ms')}
    out['patch_radius_ms'] = torch.from_numpy(rng.uniform(0.05, 0.3, B).astype(np.float32))
    out['imp_surf_magnitude_ms'] = torch.from_numpy(rng.uniform(0.0, 0.1, B).astype(np.float32))
    out['imp_surf_dist_sign_ms'] = torch.from_numpy((rng.rand(B) < 0.5).astype(np.float32))
    return out
