"""points2surf_b200.evaluation against the reference's own output (tests/golden/evaluation.npz, generated from the
unmodified source/base/evaluation.py by tests/golden/make_golden.py) -- host-side report logic, CPU only."""
import contextlib
import io
import os

import numpy as np
import torch

from points2surf_b200 import evaluation as ev

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'evaluation.npz'))


def test_eval_predictions_report_matches_reference(tmp_path):
    pred, gt = tmp_path / 'pred', tmp_path / 'gt'
    pred.mkdir(); gt.mkdir()
    for n in GOLD['names']:
        np.save(pred / (str(n) + '.xyz.npy'), GOLD['pred_' + str(n)])
        np.save(gt / (str(n) + '.ply.npy'), GOLD['gt_' + str(n)])
    for uns in (0, 1):
        rep = tmp_path / ('rep%d.csv' % uns)
        with contextlib.redirect_stdout(io.StringIO()):
            ev.eval_predictions(str(pred), str(gt), str(rep), unsigned=bool(uns))
        assert rep.read_text() == str(GOLD['report_unsigned%d' % uns])


def test_binary_comparison_matches_reference():
    res = ev.compare_predictions_binary_tensors(torch.from_numpy(GOLD['bin_a']), torch.from_numpy(GOLD['bin_b']), 'cmp')
    for k, v in zip(GOLD['bin_keys'], GOLD['bin_vals']):
        assert res[str(k)] == v or (np.isnan(v) and np.isnan(res[str(k)])), k
    assert res['comp_name'] == 'cmp'
    try:
        ev.compare_predictions_binary_tensors(torch.zeros(3), torch.zeros(4), 'x')
        assert False
    except ValueError:
        pass


def test_table_printer_matches_reference():
    rows = [{'file': 'abc_def_ghi_jkl', 'mse': 1.23456789, 'x': 2.0}, {'file': 'zz', 'mse': 0.5, 'x': -1.0}]
    for mode in ('latex', 'csv'):
        with contextlib.redirect_stdout(io.StringIO()):
            assert '\n'.join(ev.print_list_of_dicts(rows, None, mode)) == str(GOLD['table_' + mode])
    assert ev.print_list_of_dicts([]) == 'WARNING: comp_res is empty'


def test_mesh_comparison_argument_errors(tmp_path):
    (tmp_path / 'new').mkdir(); (tmp_path / 'ref').mkdir()
    try:
        ev.mesh_comparison(str(tmp_path / 'new'), str(tmp_path / 'ref'), 1, str(tmp_path / 'r.csv'),
                           dataset_file_abs=str(tmp_path / 'missing.txt'))
        assert False
    except ValueError as e:
        assert 'File does not exist' in str(e)
    (tmp_path / 'set.txt').write_text('a\nb\n')
    try:   # nothing to compare -> the reference raises too (evaluation.py:352-353)
        ev.mesh_comparison(str(tmp_path / 'new'), str(tmp_path / 'ref'), 1, str(tmp_path / 'r.csv'),
                           dataset_file_abs=str(tmp_path / 'set.txt'))
        assert False
    except ValueError as e:
        assert 'empty' in str(e)
    assert ev.mesh_comparison(str(tmp_path / 'nope'), str(tmp_path / 'ref'), 1, str(tmp_path / 'r.csv')) is None
