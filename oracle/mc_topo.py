"""TEST INFRASTRUCTURE ONLY -- a second, table-free restatement of marching cubes, independent of tools/gen_mc_tables.py.

PARITY UNPINNED (skimage.measure.marching_cubes_lewiner, source/sdf.py:215, is neither installed nor installable).
The CUDA kernel (points2surf_b200/csrc/mc.cu) and oracle/mc_oracle.py both look triangles up in the table written by
tools/gen_mc_tables.py.  This module shares no table and no code with that generator: it walks every sign-changing cell on
its VALUES, builds the iso-polygons face by face and triangulates them, so that
  * face_rule='separate_positive' reproduces the kernel's rule by an independent route (tests require identical vertex
    arrays and identical triangle sets), and
  * face_rule='asymptotic' resolves ambiguous faces the way Lewiner et al. 2003 ("Efficient implementation of Marching
    Cubes' cases with topological guarantees", JGT 8(2); face test = Nielson & Hamann's asymptotic decider on the bilinear
    face interpolant, Chernyaev's MC33) do, which is what the reference's skimage call does.  The tests report where the two
    rules give different meshes (cells, Euler characteristic, Chamfer distance).
Not restated: the INTERIOR test of MC33 (tunnels between two sheets inside one cell: sub-cases 4.1.2, 6.1.2, 7.4.2, 10.1.2,
12.1.2, 13.5.x) -- both rules here keep the sheets separate; `multi_sheet_cells` in the returned statistics counts the cells
where that test would have been consulted.

Conventions (documented in oracle/mc_oracle.py and followed here so that arrays are comparable element by element):
a corner is positive iff value > level (exact zeros are NOT positive -- scikit-image's Cython port builds the case index with
`value > isovalue`; Lewiner's C++ nudges zeros to +epsilon instead, which would wrap every zero-filled far corner of a
propagated volume, SURVEY section 3.3, in a spurious sheet); one vertex per sign-changing grid edge, numbered by ascending
3*lin(lower end) + axis; linear interpolation in float32; index space -> model space ((v+0.5)/res-0.5)*2 (sdf.py:224);
polygons are fan-triangulated from their smallest cube-edge label; all faces flipped if the signed volume is negative
(trimesh.repair.fix_inversion, sdf.py:226).
"""
import numpy as np

# cube corner c = x + 2y + 4z; cube edge label = 4*axis + (other two coordinates of the lower end, low bit first)


def _edge_label(c0, c1):
    d = c0 ^ c1
    axis = {1: 0, 2: 1, 4: 2}[d]
    lo = min(c0, c1)
    p = [(lo >> i) & 1 for i in range(3)]
    o = [p[i] for i in range(3) if i != axis]
    return 4 * axis + o[0] + 2 * o[1]


def _face_rings():
    """Six faces, corners in counter-clockwise order seen from OUTSIDE the cube."""
    rings = []
    for a in range(3):
        u, v = (a + 1) % 3, (a + 2) % 3                   # (a, u, v) right-handed
        for side in (0, 1):
            ring = []
            for pu, pv in ((0, 0), (1, 0), (1, 1), (0, 1)):    # ccw seen from +a
                p = [0, 0, 0]
                p[a], p[u], p[v] = side, pu, pv
                ring.append(p[0] + 2 * p[1] + 4 * p[2])
            rings.append(ring if side == 1 else ring[::-1])
    return rings


_RINGS = _face_rings()
_EDGE_ENDS = {}
for _c0 in range(8):
    for _b in (1, 2, 4):
        if not _c0 & _b:
            _EDGE_ENDS[_edge_label(_c0, _c0 | _b)] = (_c0, _c0 | _b)


def cell_polygons(val, level, face_rule):
    """val[8] corner values -> list of polygons (lists of cube-edge labels), each starting at its smallest label."""
    pos = [bool(v > level) for v in val]
    nxt = {}
    for ring in _RINGS:
        s = [pos[c] for c in ring]
        n = sum(s)
        if n in (0, 4):
            continue
        ambiguous = (n == 2 and s[0] == s[2])
        connect_pos = False
        if ambiguous and face_rule == 'asymptotic':
            # bilinear interpolant on the face: value at the saddle = (A*C - B*D) / (A + C - B - D), A, C / B, D the diagonals
            A, B, C_, D = (float(val[c]) - float(level) for c in ring)
            num, den = A * C_ - B * D, A + C_ - B - D
            connect_pos = (num / den > 0.0) if den != 0.0 and num != 0.0 else False
        elif face_rule not in ('asymptotic', 'separate_positive'):
            raise ValueError(face_rule)
        if ambiguous and connect_pos:
            # the positive corners are joined through the face: cut off each NEGATIVE corner (walk counter-clockwise round it)
            for k in range(4):
                if not s[k]:
                    nxt[_edge_label(ring[k - 1], ring[k])] = _edge_label(ring[k], ring[(k + 1) % 4])
        else:
            # one segment per maximal run of positive corners: from the edge behind the run to the edge in front of it
            for i in range(4):
                if s[i] and not s[i - 1]:
                    j = i
                    while s[(j + 1) % 4]:
                        j = (j + 1) % 4
                    nxt[_edge_label(ring[j], ring[(j + 1) % 4])] = _edge_label(ring[i - 1], ring[i])
    polys, seen = [], set()
    for e0 in sorted(nxt):
        if e0 in seen:
            continue
        loop, e = [], e0
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == e0 and len(loop) >= 3, 'open iso-polygon'
        polys.append(loop)
    return polys


def marching_cubes(vol, level=0.0, face_rule='asymptotic', return_stats=False):
    """vol [R,R,R] -> (verts [V,3] float32 model space, faces [F,3] int32[, stats])."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    R = vol.shape[0]
    assert vol.shape == (R, R, R)
    lvl = np.float32(level)
    pos = vol > lvl
    # ---- vertices: one per sign-changing grid edge, ascending key 3*lin(lower end) + axis
    keys, pts = [], []
    for axis in range(3):
        a = [slice(None)] * 3
        b = [slice(None)] * 3
        a[axis], b[axis] = slice(0, R - 1), slice(1, R)
        cross = pos[tuple(a)] != pos[tuple(b)]
        idx = np.argwhere(cross)
        v0 = vol[tuple(a)][cross]
        v1 = vol[tuple(b)][cross]
        t = (lvl - v0) / (v1 - v0)
        p = idx.astype(np.float32)
        p[:, axis] += t
        keys.append(3 * ((idx[:, 0] * R + idx[:, 1]) * R + idx[:, 2]) + axis)
        pts.append(p)
    keys = np.concatenate(keys)
    order = np.argsort(keys, kind='stable')
    keys = keys[order]
    p = np.concatenate(pts)[order]
    verts = (((p + np.float32(0.5)) / np.float32(R)) - np.float32(0.5)) * np.float32(2.0)
    vid = {int(k): i for i, k in enumerate(keys)}
    # ---- faces: cell by cell in C order
    code = np.zeros((R - 1,) * 3, dtype=np.int32)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        code |= pos[dx:R - 1 + dx, dy:R - 1 + dy, dz:R - 1 + dz].astype(np.int32) << c
    cells = np.argwhere((code != 0) & (code != 255))
    faces = []
    stats = dict(cells=int(len(cells)), ambiguous_face_cells=0, multi_sheet_cells=0, connected_faces=0)
    for cx, cy, cz in cells:
        val = [vol[cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)] for c in range(8)]
        polys = cell_polygons(val, lvl, face_rule)
        if len(polys) > 1:
            stats['multi_sheet_cells'] += 1
        if return_stats:
            amb = 0
            for ring in _RINGS:
                s = [val[c] > lvl for c in ring]
                amb += int(sum(s) == 2 and s[0] == s[2])
            stats['ambiguous_face_cells'] += int(amb > 0)
        for loop in polys:
            ids = []
            for e in loop:
                c0, _ = _EDGE_ENDS[e]
                g = 3 * (((cx + (c0 & 1)) * R + (cy + ((c0 >> 1) & 1))) * R + (cz + ((c0 >> 2) & 1))) + e // 4
                ids.append(vid[int(g)])
            for i in range(1, len(ids) - 1):
                faces.append((ids[0], ids[i], ids[i + 1]))
    faces = np.array(faces, dtype=np.int32).reshape(-1, 3)
    if len(faces):
        v0, v1, v2 = (verts[faces[:, i]].astype(np.float64) for i in range(3))
        if np.einsum('ij,ij->i', v0, np.cross(v1, v2)).sum() / 6.0 < 0:
            faces = faces[:, [0, 2, 1]].copy()
    verts = verts.astype(np.float32)
    return (verts, faces, stats) if return_stats else (verts, faces)


def triangle_set(faces):
    """Unordered triangles as a sorted array of sorted vertex triples (winding and fan rotation removed)."""
    f = np.sort(np.asarray(faces, dtype=np.int64), axis=1)
    return f[np.lexsort((f[:, 2], f[:, 1], f[:, 0]))]


def euler_characteristic(n_verts, faces):
    f = np.asarray(faces, dtype=np.int64)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    n_edges = len(np.unique(e[:, 0] * (int(f.max()) + 1) + e[:, 1]))
    used = len(np.unique(f))
    return used - n_edges + len(f)


def all_cases_volume(seed=0):
    """A 49^3 volume that contains every one of the 256 corner-sign configurations (each in its own cell, separated by
    all-negative voxels) with random magnitudes, three times with different magnitudes so that ambiguous faces fall on both
    sides of the asymptotic decider; plus two exact zeros."""
    rng = np.random.RandomState(seed)
    n = 16
    R = 3 * n + 1
    vol = np.full((R, R, R), -1.0, dtype=np.float32)
    for rep in range(3):
        bz = 1 + 4 * rep
        for case in range(256):
            bx, by = 3 * (case % n) + 1, 3 * (case // n) + 1
            for c in range(8):
                mag = rng.uniform(0.05, 1.0)
                vol[bx + (c & 1), by + ((c >> 1) & 1), bz + ((c >> 2) & 1)] = mag if (case >> c) & 1 else -mag
    vol[5, 5, 20] = 0.0          # isolated exact zeros: not positive, no surface around them
    vol[7, 8, 1] = 0.0           # an exact zero inside a configuration block
    return vol
