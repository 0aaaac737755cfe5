"""TEST INFRASTRUCTURE ONLY -- CPU restatement of one training iteration of the reference
(source/points_to_surf_train.py:441-461: zero_grad, train-mode forward, compute_loss :537-563, backward,
optim.SGD(momentum) :406) as functional torch code with autograd.  Pinned against the unmodified reference modules
(PointsToSurfModel.train(), sdf_nn.calc_loss_*, torch.optim.SGD) by tests/golden/make_golden.py:gen_train, which
asserts equality of losses, every gradient and every updated tensor before writing tests/golden/train_*.npz.
Nothing outside tests/, bench.py's CPU legs and __graft_entry__.smoke() may import this module."""
import torch
import torch.nn.functional as F

from oracle.p2s_oracle import quat_to_rotmat

EPS, MOM = 1e-5, 0.1


def _bn(sd, bufs, x, bn, relu):
    y = F.batch_norm(x, bufs[bn + '.running_mean'], bufs[bn + '.running_var'], sd[bn + '.weight'], sd[bn + '.bias'],
                     training=True, momentum=MOM, eps=EPS)
    return F.relu(y) if relu else y


def _conv(sd, bufs, x, conv, bn, relu=True):
    return _bn(sd, bufs, F.conv1d(x, sd[conv + '.weight'], sd[conv + '.bias']), bn, relu)


def _fc(sd, bufs, x, fc, bn=None, relu=True):
    y = F.linear(x, sd[fc + '.weight'], sd[fc + '.bias'])
    return _bn(sd, bufs, y, bn, relu) if bn is not None else y


def _stn(sd, bufs, p, x):
    """conv1-3 + max + fc1-3 of STN / QSTN (source/points_to_surf_model.py:41-62, 100-121), raw fc3 output."""
    x = _conv(sd, bufs, x, p + 'conv1', p + 'bn1')
    x = _conv(sd, bufs, x, p + 'conv2', p + 'bn2')
    x = _conv(sd, bufs, x, p + 'conv3', p + 'bn3')
    x = torch.max(x, dim=2)[0]
    x = _fc(sd, bufs, x, p + 'fc1', p + 'bn4')
    x = _fc(sd, bufs, x, p + 'fc2', p + 'bn5')
    return _fc(sd, bufs, x, p + 'fc3')


def _qstn(sd, bufs, p, x):
    return quat_to_rotmat(_stn(sd, bufs, p, x) + x.new_tensor([1, 0, 0, 0]))


def _feat(sd, bufs, p, x, point_stn):
    trans = None
    if point_stn:                                                   # model.py:181-186
        trans = _qstn(sd, bufs, p + 'stn1.', x)
        x = torch.bmm(trans, x)
    x = _conv(sd, bufs, x, p + 'conv0a', p + 'bn0a')
    x = _conv(sd, bufs, x, p + 'conv0b', p + 'bn0b')
    t2 = _stn(sd, bufs, p + 'stn2.', x)
    t2 = (t2 + torch.eye(64, dtype=x.dtype).view(1, 4096)).view(-1, 64, 64)
    x = torch.bmm(t2, x)
    x = _conv(sd, bufs, x, p + 'conv1', p + 'bn1')
    x = _conv(sd, bufs, x, p + 'conv2', p + 'bn2')
    x = _conv(sd, bufs, x, p + 'conv3', p + 'bn3', relu=False)
    return torch.max(x, dim=2)[0], trans


def forward_train(sd, bufs, batch, use_point_stn, shared_transformer):
    """PointsToSurfModel.forward in training mode (model.py:296-352); `sd` parameters (leaf tensors), `bufs`
    running statistics (updated in place like nn.BatchNorm1d)."""
    patch = batch['patch_pts_ps'].transpose(1, 2)
    shape = batch['pts_sub_sample_ms'].transpose(1, 2) - batch['imp_surf_query_point_ms'].unsqueeze(2)
    if use_point_stn and shared_transformer:
        trans = _qstn(sd, bufs, 'point_stn.', torch.cat((patch, shape), dim=2))
        shape, patch = torch.bmm(trans, shape), torch.bmm(trans, patch)
    sf, trans_g = _feat(sd, bufs, 'feat_global.', shape, bool(use_point_stn and not shared_transformer))
    sf = _fc(sd, bufs, sf, 'fc1_global', 'bn1_global')
    if use_point_stn and not shared_transformer:
        patch = torch.bmm(trans_g, patch)
    pf, _ = _feat(sd, bufs, 'feat_local.', patch, False)
    pf = _fc(sd, bufs, pf, 'fc1_local', 'bn1_local')
    x = torch.cat((pf, sf), dim=1)
    x = _fc(sd, bufs, x, 'fc2', 'bn2')
    x = _fc(sd, bufs, x, 'fc3', 'bn3')
    return _fc(sd, bufs, x, 'fc4')


def losses(pred, batch, w_mag=1.0, w_sign=1.0, fixed_radius=False):
    """compute_loss (points_to_surf_train.py:550-561) with sdf_nn.calc_loss_magnitude / calc_loss_sign (:30-40)."""
    t = batch['imp_surf_magnitude_ms'].reshape(-1)
    if not fixed_radius:
        t = t / batch['patch_radius_ms'].reshape(-1)
    l_mag = F.mse_loss(torch.tanh(torch.abs(pred[:, 0])), torch.tanh(torch.abs(t))) * w_mag
    l_sign = F.binary_cross_entropy_with_logits(pred[:, 1], batch['imp_surf_dist_sign_ms'].reshape(-1),
                                                reduction='none').mean() * w_sign
    return [l_mag, l_sign]


def train_iteration(state_dict, batch, use_point_stn, shared_transformer, lr=0.01, momentum=0.9, mom_bufs=None,
                    dtype=torch.float32):
    """-> dict(logits, losses, grads{name}, new_state{name}, mom_bufs{name}); `state_dict` is not modified.
    dtype=float64 gives the rounding-free reference used to check the sequencing of the hand-written backward."""
    batch = {k: v.to(dtype) for k, v in batch.items()}
    sd, bufs = {}, {}
    for k, v in state_dict.items():
        if k.endswith('num_batches_tracked'):
            continue
        if k.endswith('running_mean') or k.endswith('running_var'):
            bufs[k] = v.detach().clone().to(dtype)
        else:
            sd[k] = v.detach().clone().to(dtype).requires_grad_(True)
    with torch.enable_grad():
        logits = forward_train(sd, bufs, batch, use_point_stn, shared_transformer)
        ls = losses(logits, batch)
        sum(ls).backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items()}
    new_state, new_mom = {}, {}
    for k, v in sd.items():
        buf = grads[k].clone() if mom_bufs is None else momentum * mom_bufs[k] + grads[k]
        new_mom[k] = buf
        new_state[k] = (v.detach() - lr * buf)
    new_state.update(bufs)
    return {'logits': logits.detach(), 'losses': [float(l.detach()) for l in ls], 'grads': grads, 'new_state': new_state,
            'mom_bufs': new_mom}
