"""TEST INFRASTRUCTURE ONLY -- shims that let the *unmodified* reference
(/root/reference, read-only, present only in the build container) be imported
with current NumPy/SciPy and without trimesh/scikit-image, so that golden
vectors can be generated from the reference itself (tests/golden/make_golden.py).

Nothing in the product path (points2surf_b200/) may import this module.
The shims are the ones listed in SURVEY.md section 8(c):
  (1) stub `trimesh` package (imported at module level by source/data_loader.py:8,
      source/sdf.py:4, source/base/point_cloud.py:3),
  (2) `np.int = int` (source/sdf.py:75),
  (3) cKDTree.query(n_jobs=) -> workers= (source/base/point_cloud.py:175,177),
  (4) scipy.ndimage.filters alias (source/sdf.py:54,122).
"""
import sys
import types

import numpy as np

REFERENCE_ROOT = '/root/reference'


def _random_rotation_matrix(rand=None):
    # Same algorithm as trimesh.transformations.random_rotation_matrix /
    # random_quaternion (Shoemake, Graphics Gems III) -- used only by the
    # reference's non-reconstruction eval pass (source/data_loader.py:381-393).
    if rand is None:
        rand = np.random.rand(3)
    r1 = np.sqrt(1.0 - rand[0])
    r2 = np.sqrt(rand[0])
    pi2 = np.pi * 2.0
    t1 = pi2 * rand[1]
    t2 = pi2 * rand[2]
    q = np.array([np.cos(t2) * r2, np.sin(t1) * r1, np.cos(t1) * r1, np.sin(t2) * r2])
    n = np.dot(q, q)
    q = q * np.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([
        [1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0], 0.0],
        [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0], 0.0],
        [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2], 0.0],
        [0.0, 0.0, 0.0, 1.0]])


def _transform_points(points, matrix):
    points = np.asanyarray(points, dtype=np.float64)
    return np.dot(matrix[:3, :3], points.T).T + matrix[:3, 3]


def install():
    """Install the shims and put the reference on sys.path. Idempotent."""
    if 'trimesh' not in sys.modules:
        tm = types.ModuleType('trimesh')

        class Trimesh:  # only needed for annotations / never instantiated by the oracle path
            def __init__(self, *a, **k):
                raise RuntimeError('trimesh stub: Trimesh is not available')
        tm.Trimesh = Trimesh
        for sub in ('transformations', 'proximity', 'repair', 'sample', 'points', 'path'):
            m = types.ModuleType('trimesh.' + sub)
            setattr(tm, sub, m)
            sys.modules['trimesh.' + sub] = m
        tm.transformations.random_rotation_matrix = _random_rotation_matrix
        tm.transformations.transform_points = _transform_points
        sys.modules['trimesh'] = tm
    if not hasattr(np, 'int'):
        np.int = int
    import scipy.spatial as spatial
    if not getattr(spatial.cKDTree, '_p2s_shimmed', False):
        _orig = spatial.cKDTree

        class cKDTree(_orig):
            _p2s_shimmed = True

            def query(self, x, k=1, eps=0, p=2, distance_upper_bound=np.inf, n_jobs=None, workers=1):
                return _orig.query(self, x, k=k, eps=eps, p=p,
                                   distance_upper_bound=distance_upper_bound,
                                   workers=(n_jobs if n_jobs is not None else workers))

            def query_ball_point(self, x, r, p=2., eps=0, n_jobs=None, workers=1, **kw):
                return _orig.query_ball_point(self, x, r, p=p, eps=eps,
                                              workers=(n_jobs if n_jobs is not None else workers), **kw)
        spatial.cKDTree = cKDTree
    import scipy.ndimage
    if 'scipy.ndimage.filters' not in sys.modules:
        try:
            import scipy.ndimage.filters  # noqa: F401  (deprecated alias still present in 1.x)
        except Exception:
            m = types.ModuleType('scipy.ndimage.filters')
            m.convolve = scipy.ndimage.convolve
            sys.modules['scipy.ndimage.filters'] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
