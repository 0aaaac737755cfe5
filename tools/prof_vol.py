"""Profiling driver for the volume stage: reconstruct one 10k-point sphere with the fitted checkpoint, then run
scatter -> sign propagation -> marching cubes (ncu: -k regex:propagate_kernel|mc_)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from points2surf_b200 import ops, synth

res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
v = synth.VARIANTS['vanilla']
sd = synth.make_state_dict('vanilla', 6, fitted=True)
eng = ops.Engine(sd, v['use_point_stn'], v['shared_transformer'], precision='tc', guard_band=0.05)
pts = torch.from_numpy(synth.make_cloud('sphere', 10000, seed=0)).cuda()
lin, sdf = eng.reconstruct(pts, res, 3, 0, 40938661)
torch.cuda.synchronize()
for _ in range(reps):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    vol, it = ops.sdf_to_volume(lin, sdf, res, 5, 13.0)
    e1.record()
    mv, mf = ops.marching_cubes(vol, 0.0)
    e2.record()
    torch.cuda.synchronize()
    print('res %d: sdf_to_volume %.3f ms (%d iterations), marching cubes %.3f ms (%d faces)' % (res, e0.elapsed_time(e1), it, e1.elapsed_time(e2), mf.shape[0]))
