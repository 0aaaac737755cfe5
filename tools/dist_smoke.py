"""Multi-GPU smoke of the two host-side loops under torchrun (one process per GPU, NCCL):
  eval : points2surf_b200.eval.points_to_surf_eval in reconstruction mode over 5 shapes -- LPT assignment by candidate count,
         every shape written exactly once, outputs identical to a single-rank run of the same shape (rank 0 re-runs one);
  train: points2surf_b200.points_to_surf_train with rank-sharded batches -- parameters stay identical on all ranks
         (gradient all_reduce), rank 0 alone writes the checkpoints.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_smoke.py
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from points2surf_b200 import synth, ops  # noqa: E402
from points2surf_b200 import eval as p2s_eval  # noqa: E402
from points2surf_b200 import points_to_surf_train as p2s_train  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    base = os.path.join(tempfile.gettempdir(), 'p2s_dist_smoke')
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
        root, models = os.path.join(base, 'data'), os.path.join(base, 'models')
        for sub in ('04_pts', '05_query_pts', '05_query_dist'):
            os.makedirs(os.path.join(root, sub))
        os.makedirs(models)
        names = ['s%d' % i for i in range(5)]
        rng = np.random.RandomState(0)
        for i, n in enumerate(names):
            cloud = synth.make_cloud(['sphere', 'torus', 'box'][i % 3], 3000 + 1500 * i, seed=i)
            q = (cloud[rng.choice(len(cloud), 64, replace=False)] + rng.normal(0, 0.02, (64, 3))).astype(np.float32)
            np.save(os.path.join(root, '04_pts', n + '.xyz.npy'), cloud)
            np.save(os.path.join(root, '05_query_pts', n + '.ply.npy'), q)
            np.save(os.path.join(root, '05_query_dist', n + '.ply.npy'), (0.5 - np.linalg.norm(q, axis=1)).astype(np.float32))
        open(os.path.join(root, 'testset.txt'), 'w').write('\n'.join(names) + '\n')
        open(os.path.join(root, 'trainset.txt'), 'w').write('\n'.join(names[:3]) + '\n')
        open(os.path.join(root, 'valset.txt'), 'w').write('\n'.join(names[3:]) + '\n')
        sd = synth.make_state_dict('vanilla', 6, fitted=True)
        torch.save({'module.' + k: t for k, t in sd.items()}, os.path.join(models, 'p2s_d_model.pth'))
        torch.save(synth.make_train_opt('vanilla'), os.path.join(models, 'p2s_d_params.pth'))
    dist.barrier()
    root, models, out = os.path.join(base, 'data'), os.path.join(base, 'models'), os.path.join(base, 'results')
    # ---- eval, LPT-sharded
    opt = p2s_eval.parse_arguments(['--indir', root, '--outdir', out, '--modeldir', models, '--models', 'p2s_d', '--dataset', 'testset.txt',
                                    '--query_grid_resolution', '32', '--epsilon', '3'])
    opt.reconstruction = True
    p2s_eval.points_to_surf_eval(opt)
    dist.barrier()
    if rank == 0:
        for i in range(5):
            d = np.load(os.path.join(out, 'rec', 'dist_ms', 's%d.xyz.npy' % i))
            assert np.isfinite(d).all() and len(d) > 100
        # a single-rank run of one shape gives the same band (Philox keyed by the query's global rank)
        v = synth.VARIANTS['vanilla']
        eng = ops.Engine(synth.make_state_dict('vanilla', 6, fitted=True), v['use_point_stn'], v['shared_transformer'], device=local,
                         precision='tc', guard_band=0.05)
        pts = torch.from_numpy(np.load(os.path.join(root, '04_pts', 's3.xyz.npy'))).to(dev)
        _, sdf = eng.reconstruct(pts, 32, 3, 0, opt.seed)
        assert np.allclose(sdf.cpu().numpy(), np.load(os.path.join(out, 'rec', 'dist_ms', 's3.xyz.npy')), atol=1e-6)
        print('dist eval: 5 shapes over %d ranks (LPT), outputs complete and identical to a single-rank run' % world)
    dist.barrier()
    # ---- training loop, rank-sharded batches
    topt = p2s_train.parse_arguments([
        '--name', 'test', '--indir', root, '--outdir', os.path.join(base, 'tmodels'), '--logdir', os.path.join(base, 'logs'),
        '--trainset', 'trainset.txt', '--testset', 'valset.txt', '--nepoch', '2', '--batchSize', '16', '--patches_per_shape', '16',
        '--points_per_patch', '300', '--sub_sample_size', '1000', '--patch_radius', '0.0', '--lr', '0.001', '--shared_transformer', '1',
        '--outputs', 'imp_surf_magnitude', 'imp_surf_sign', 'patch_pts_ids', 'p_index'])
    hist = p2s_train.points_to_surf_train(topt)
    assert all(np.isfinite(h[3]).all() for h in hist)
    dist.barrier()
    if rank == 0:
        assert os.path.isfile(os.path.join(base, 'tmodels', 'test_model.pth'))
        print('dist train: %d steps per rank on %d ranks, losses finite, checkpoint written by rank 0' % (len([h for h in hist if h[0] == 'train']), world))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
