"""CPU: the C-ABI library loads and exports every symbol include/p2s_b200.h declares; the Python
binding declares a prototype for each; host-only helpers behave; there is no CPU compute path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from points2surf_b200 import _lib, weights, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'p2s_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(p2s_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'missing export: ' + s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.p2s_abi_version() == 2


@pytest.mark.parametrize('variant', ['vanilla', 'max', 'uniform'])
def test_blob_size_matches_library(variant):
    lib = _lib.load()
    v = synth.VARIANTS[variant]
    sd = synth.make_state_dict_numpy(variant, 0, module_prefix='module.')
    blob = weights.pack_blob(sd, v['use_point_stn'], v['shared_transformer'])
    cfg = _lib.ModelConfig(v['use_point_stn'], v['shared_transformer'], 300, 1000, 1024)
    assert blob.size == lib.p2s_model_blob_floats(C.byref(cfg))
    assert blob.dtype == np.float32 and np.isfinite(blob).all()


def test_bn_folding_matches_batchnorm():
    import torch
    import torch.nn.functional as F
    sd = synth.make_state_dict('max', 1)
    w, b = weights.fold(sd, 'feat_local.conv2', 'feat_local.bn2')
    x = torch.randn(5, 64, 7)
    y = F.batch_norm(F.conv1d(x, sd['feat_local.conv2.weight'], sd['feat_local.conv2.bias']),
                     sd['feat_local.bn2.running_mean'], sd['feat_local.bn2.running_var'],
                     sd['feat_local.bn2.weight'], sd['feat_local.bn2.bias'], False, 0.0, 1e-5)
    y2 = F.conv1d(x, torch.from_numpy(w.reshape(128, 64, 1)), torch.from_numpy(b))
    assert torch.allclose(y, y2, atol=1e-4)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from points2surf_b200 import ops
    with pytest.raises(_lib.P2SError):
        ops.Engine(synth.make_state_dict('max', 0), 0, 0)
    with pytest.raises(_lib.P2SError):
        ops.query_grid(torch.zeros(10, 3), 16, 3)


def test_ply_off_roundtrip(tmp_path):
    from points2surf_b200 import mesh_io
    v = np.random.RandomState(0).rand(7, 3).astype(np.float32)
    f = np.array([[0, 1, 2], [2, 3, 4], [4, 5, 6]], np.int32)
    mesh_io.write_ply(str(tmp_path / 'm.ply'), v, f)
    v2, f2 = mesh_io.read_ply(str(tmp_path / 'm.ply'))
    assert np.array_equal(v, v2) and np.array_equal(f, f2)
    mesh_io.write_ply(str(tmp_path / 'p.ply'), v, None, colors=np.random.rand(7, 3))
    v3, f3 = mesh_io.read_ply(str(tmp_path / 'p.ply'))
    assert np.array_equal(v, v3) and len(f3) == 0
    mesh_io.write_off(str(tmp_path / 'q.off'), v, np.array([]), colors_vertex=np.random.rand(7, 3))
    lines = (tmp_path / 'q.off').read_text().split('\n')
    assert lines[0] == 'COFF' and lines[1] == '7 0 0' and len(lines[2].split()) == 6


def test_library_is_newer_than_its_sources():
    """A stale libp2s_b200.so travelling to the GPU box silently tests old kernels: rebuild with
    `python -m points2surf_b200.build` whenever csrc/ or include/ changes."""
    import glob
    from points2surf_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    srcs = glob.glob(os.path.join(root, 'points2surf_b200', 'csrc', '*.cu*')) + glob.glob(os.path.join(root, 'include', '*.h'))
    assert srcs
    newest = max(os.path.getmtime(f) for f in srcs)
    assert os.path.getmtime(_lib.LIB_PATH) >= newest, 'libp2s_b200.so is older than its sources'


def test_product_never_imports_the_oracle_or_tests():
    """oracle/ and tests/ are test infrastructure: nothing under points2surf_b200/ (nor bench.py's product arm) may
    import them -- a product path routed through the oracle would void every parity claim."""
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r'^\s*(from|import)\s+(oracle|tests|helpers|helpers_train)\b', re.M)
    for f in glob.glob(os.path.join(root, 'points2surf_b200', '**', '*.py'), recursive=True):
        assert not pat.search(open(f).read()), f
    # bench.py may use the oracle only inside the CPU legs
    src = open(os.path.join(root, 'bench.py')).read()
    body = src[src.index('def run_b200('):src.index("if __name__ == '__main__':")]
    assert 'oracle' not in body.replace('oracle port', ''), 'run_b200 must not touch the oracle'
