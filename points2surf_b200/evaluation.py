"""Mirror of the reference's `source.base.evaluation` for the reconstruction path:
`eval_predictions` and `mesh_comparison` with the same signatures, CSV layout and sentinel values; the sampling and
the nearest-neighbour distances run on the GPU (csrc/meshdist.cu through the C ABI), there is no CPU path.

    reference                                              here
    _chamfer_distance_single_file   evaluation.py:222-256  _chamfer_distance_single_file  (p2s_mesh_sample_dev + p2s_chamfer_hausdorff_dev)
    _hausdorff_distance_single_file evaluation.py:284-304  _hausdorff_distance_single_file
    mesh_comparison                 evaluation.py:307-392  mesh_comparison
    eval_predictions                evaluation.py:84-127   eval_predictions
    print_list_of_dicts             evaluation.py:130-182  print_list_of_dicts
    compare_predictions_binary_tensors evaluation.py:39-81 compare_predictions_binary_tensors

Deviation (documented): the reference samples with trimesh.sample.sample_surface_even (area-weighted sampling followed
by a minimum-distance rejection); trimesh is absent, so the sampler here is the area-weighted part only and the sample
count is exactly `samples_per_model`.  Both estimate the same surface integrals.
"""
import os

import numpy as np
import torch

from . import mesh_io, ops


def calc_accuracy(num_true, num_predictions):
    return float('NaN') if num_predictions == 0 else num_true / num_predictions


def calc_precision(num_true_pos, num_false_pos):
    if isinstance(num_true_pos, (int, float)) and isinstance(num_false_pos, (int, float)) \
            and num_true_pos + num_false_pos == 0:
        return float('NaN')
    return num_true_pos / (num_true_pos + num_false_pos)


def calc_recall(num_true_pos, num_false_neg):
    if isinstance(num_true_pos, (int, float)) and isinstance(num_false_neg, (int, float)) \
            and num_true_pos + num_false_neg == 0:
        return float('NaN')
    return num_true_pos / (num_true_pos + num_false_neg)


def calc_f1(precision, recall):
    if isinstance(precision, (int, float)) and isinstance(recall, (int, float)) and precision + recall == 0:
        return float('NaN')
    return 2.0 * (precision * recall) / (precision + recall)


def compare_predictions_binary_tensors(ground_truth, predicted, prediction_name):
    """Confusion counts of (x > 0) for two dense tensors (evaluation.py:39-81); same keys as the reference."""
    if ground_truth.shape != predicted.shape:
        raise ValueError('The ground truth matrix and the predicted matrix have different sizes!')
    if not isinstance(ground_truth, torch.Tensor) and not isinstance(predicted, torch.Tensor):
        raise ValueError('Both matrices must be dense of type torch.tensor!')
    gt = ground_truth > 0.0
    pr = predicted > 0.0
    n = float(gt.numel())
    res = {'comp_name': prediction_name, 'predictions': n, 'pred_gt': n}
    res['positives'] = float(pr.sum())
    res['pos_gt'] = float(gt.sum())
    res['true_neg'] = float((~pr & ~gt).sum())
    res['negatives'] = n - res['positives']
    res['neg_gt'] = n - res['pos_gt']
    res['true_pos'] = float((pr & gt).sum())
    res['true'] = res['true_pos'] + res['true_neg']
    res['false_pos'] = float((pr & ~gt).sum())
    res['false_neg'] = float((~pr & gt).sum())
    res['false'] = res['false_pos'] + res['false_neg']
    res['accuracy'] = calc_accuracy(res['true'], res['predictions'])
    res['precision'] = calc_precision(res['true_pos'], res['false_pos'])
    res['recall'] = calc_recall(res['true_pos'], res['false_neg'])
    res['f1_score'] = calc_f1(res['precision'], res['recall'])
    return res


def print_list_of_dicts(comp_res, keys_to_print=None, mode='latex'):
    """One line per dict, right-aligned columns, sorted, header first (evaluation.py:130-182)."""
    if len(comp_res) == 0:
        return 'WARNING: comp_res is empty'
    if not keys_to_print:
        keys_to_print = list(comp_res[0].keys())

    def sep(i):
        if mode == 'latex':
            return ' & ' if i < len(keys_to_print) - 1 else ' \\\\'
        return ','

    lines = []
    for d in comp_res:
        line = ''
        for i, key in enumerate(keys_to_print):
            width = max(10, len(key))
            if isinstance(d[key], str):
                line += d[key][:10].replace('_', ' ').rjust(width) + sep(i)
            else:
                line += '{0:.5f}'.format(d[key]).rjust(width) + sep(i)
        lines.append(line)
    lines.sort()
    lines.insert(0, ''.join(key.replace('_', ' ').rjust(10) + sep(i) for i, key in enumerate(keys_to_print)))
    for l in lines:
        print(l)
    return lines


def eval_predictions(pred_path, gt_path, report_file=None, unsigned=False):
    """MSE / mean / variance of predicted vs ground-truth SDF samples per file (evaluation.py:84-127).
    Small per-file arrays (2 000 query points): host arithmetic, no kernel involved."""
    files = [f for f in os.listdir(pred_path) if os.path.isfile(os.path.join(pred_path, f)) and f[-4:] == '.npy']
    results = []
    for f in files:
        mat_gt = np.load(os.path.join(gt_path, f[:-8] + '.ply.npy'))
        mat_rec = np.load(os.path.join(pred_path, f))
        if unsigned:
            mat_gt, mat_rec = np.abs(mat_gt), np.abs(mat_rec)
        nz = (mat_rec != 0.0) | (mat_gt != 0.0)
        diff = mat_rec - mat_gt
        gm, rm = mat_gt.mean(), mat_rec.mean()
        results.append({'file': f, 'mse': (diff * diff)[nz].mean(), 'mean_gt': gm, 'mean_pred': rm,
                        'var_gt': (mat_gt * mat_gt).mean() - gm * gm, 'var_pred': (mat_rec * mat_rec).mean() - rm * rm})
    print('compare_prediction: {} vs {}\n'.format(gt_path, pred_path))
    lines = print_list_of_dicts(results, ['file', 'mse', 'mean_gt', 'mean_pred', 'var_gt', 'var_pred'], mode='csv')
    if report_file is not None:
        mesh_io.make_dir_for_file(report_file)
        with open(report_file, 'w') as fp:
            for l in lines:
                fp.write(l + '\n')


# ------------------------------------------------------------------------------------------------ mesh metrics

def _device():
    if not torch.cuda.is_available():
        raise ops.P2SError('points2surf_b200.evaluation needs a CUDA device (there is no CPU path)')
    return torch.device('cuda', torch.cuda.current_device())


def _seed_for(path):
    # stable per-file Philox seed so that repeated reports are reproducible
    h = 1469598103934665603
    for ch in os.path.basename(path).encode():
        h = ((h ^ ch) * 1099511628211) & (2**64 - 1)
    return h


def sample_mesh(mesh_file, num_samples, seed=None):
    """-> [n,3] fp32 CUDA tensor; an unreadable or empty mesh yields [0,3] like the reference (evaluation.py:229-236).

    Deviation from the reference (parity unpinned, trimesh is not installed): the reference calls
    `trimesh.sample.sample_surface_even`, i.e. area-weighted sampling FOLLOWED by an even-spacing rejection that may
    return fewer than `num_samples` points; this sampler is the area-weighted part only and always returns
    `num_samples`.  The reference's Chamfer distance is a SUM over the samples, so absolute values are comparable
    only at equal sample counts (divide by the sample count for a per-sample mean)."""
    dev = _device()
    try:
        verts, faces = mesh_io.read_mesh(mesh_file)
    except Exception:
        return torch.zeros((0, 3), dtype=torch.float32, device=dev)
    if verts.shape[0] == 0 or faces is None or faces.shape[0] == 0:
        return torch.zeros((0, 3), dtype=torch.float32, device=dev)
    v = torch.from_numpy(np.ascontiguousarray(verts, dtype=np.float32)).to(dev)
    f = torch.from_numpy(np.ascontiguousarray(faces, dtype=np.int32)).to(dev)
    return ops.mesh_sample(v, f, num_samples, _seed_for(mesh_file) if seed is None else seed)


def _chamfer_distance_single_file(file_in, file_ref, samples_per_model, num_processes=1):
    new_s = sample_mesh(file_in, samples_per_model)
    ref_s = sample_mesh(file_ref, samples_per_model)
    if new_s.shape[0] == 0 or ref_s.shape[0] == 0:
        return file_in, file_ref, -1.0
    return file_in, file_ref, ops.chamfer_hausdorff(new_s, ref_s)['chamfer']


def _hausdorff_distance_directed_single_file(file_in, file_ref, samples_per_model):
    new_s = sample_mesh(file_in, samples_per_model)
    ref_s = sample_mesh(file_ref, samples_per_model)
    if new_s.shape[0] == 0 or ref_s.shape[0] == 0:
        return file_in, file_ref, -1.0
    return file_in, file_ref, ops.chamfer_hausdorff(new_s, ref_s)['hausdorff_ab']


def _hausdorff_distance_single_file(file_in, file_ref, samples_per_model):
    new_s = sample_mesh(file_in, samples_per_model)
    ref_s = sample_mesh(file_ref, samples_per_model)
    if new_s.shape[0] == 0 or ref_s.shape[0] == 0:
        return file_in, file_ref, -1.0, -1.0, -1.0
    r = ops.chamfer_hausdorff(new_s, ref_s)
    return file_in, file_ref, r['hausdorff_ab'], r['hausdorff_ba'], r['hausdorff']


def mesh_comparison(new_meshes_dir_abs, ref_meshes_dir_abs, num_processes, report_name, samples_per_model=10000,
                    dataset_file_abs=None):
    """Hausdorff + Chamfer report over two mesh directories (evaluation.py:307-392).  `num_processes` is accepted for
    signature compatibility; the GPU evaluates one pair at a time (a 10k x 10k pair takes well under a millisecond)."""
    if not os.path.isdir(new_meshes_dir_abs):
        print('Warning: dir to check doesn\'t exist: {}'.format(new_meshes_dir_abs))
        return
    new_files = [f for f in os.listdir(new_meshes_dir_abs) if os.path.isfile(os.path.join(new_meshes_dir_abs, f))]
    ref_files = [f for f in os.listdir(ref_meshes_dir_abs) if os.path.isfile(os.path.join(ref_meshes_dir_abs, f))]
    if dataset_file_abs is None:
        # the reference compares full file names of the reference directory against *stems* of the new meshes
        # (evaluation.py:319,349), so without a dataset file only extension-less names ever match; kept as is
        to_compare = set(ref_files)
    else:
        if not os.path.isfile(dataset_file_abs):
            raise ValueError('File does not exist: {}'.format(dataset_file_abs))
        with open(dataset_file_abs) as fp:
            to_compare = set(l.replace('\n', '').split('.')[0] for l in fp.readlines())

    def stem(f):
        return f.split('.')[0]

    def ref_for(new_file):
        return list(set(f for f in ref_files if stem(f) == stem(new_file)))

    pairs = []
    for new_file in new_files:
        if stem(new_file) in to_compare:
            match = ref_for(new_file)
            if match:
                pairs.append((os.path.join(new_meshes_dir_abs, new_file), os.path.join(ref_meshes_dir_abs, match[0])))
    if len(pairs) == 0:
        raise ValueError('Results are empty!')
    results = []
    for file_in, file_ref in pairs:
        h = _hausdorff_distance_single_file(file_in, file_ref, samples_per_model)
        c = _chamfer_distance_single_file(file_in, file_ref, samples_per_model, num_processes)
        results.append((h[0], h[1], str(h[2]), str(h[3]), str(h[4]), str(c[2])))

    for new_file in new_files:   # reconstruction without reference
        if stem(new_file) not in to_compare:
            if dataset_file_abs is None:
                match = ref_for(new_file)
                if match:
                    results.append((os.path.join(new_meshes_dir_abs, new_file),
                                    os.path.join(ref_meshes_dir_abs, match[0]), '-2', '-2', '-2', '-2'))
        else:
            to_compare.discard(stem(new_file))
    for missing in to_compare:     # reference without reconstruction
        results.append((os.path.join(new_meshes_dir_abs, missing), os.path.join(ref_meshes_dir_abs, missing),
                        '-1', '-1', '-1', '-1'))
    results = sorted(results, key=lambda r: r[0])
    mesh_io.make_dir_for_file(report_name)
    lines = ['in mesh,ref mesh,Hausdorff dist new-ref,Hausdorff dist ref-new,Hausdorff dist,'
             'Chamfer dist(-1: no input; -2: no reference)']
    lines += [','.join(r) for r in results]
    with open(report_name, 'w') as fp:
        fp.write('\n'.join(lines))
