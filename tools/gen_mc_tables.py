"""Generate the marching-cubes case table programmatically (no table is copied from anywhere).

Cube corner c = x + 2y + 4z.  Edge e = 4*axis + (the two other coordinates of its lower end, low bit first):
axis 0: e = 0 + (y + 2z); axis 1: e = 4 + (x + 2z); axis 2: e = 8 + (x + 2y).
Case index = bitmask of corners whose value is > level ("positive").

Construction: on every cube face each maximal run of positive corners contributes one oriented iso-segment from the
edge where the run is left to the edge where it is entered.  An AMBIGUOUS face (+-+-) has two readings: the positive
corners are separated (the rule above, Lorensen & Cline 1987 / Montani et al. 1994) or joined through the face, in
which case each NEGATIVE corner is cut off instead.  Which reading applies is decided at run time from the face's four
VALUES by the asymptotic decider (sign of the bilinear interpolant at its saddle: Nielson & Hamann 1991; the face test
of Lewiner et al. 2003, i.e. of the skimage call the reference makes, source/sdf.py:215).  The decision depends on
the shared face only, so neighbouring cells agree and the surface is watertight.  The table therefore has one row per
(case, decision bits of the case's ambiguous faces in ascending face order).  Segments chain into closed loops, each
loop is fan-triangulated from its smallest edge label.  All cells use the same orientation rule, so the mesh is
consistently oriented; the global sign is fixed afterwards by the signed volume (trimesh.repair.fix_inversion in the
reference, source/sdf.py:226).  The interior test of MC33 (tunnels inside a cell) is not part of the table.

Writes points2surf_b200/csrc/mc_tables.cuh and oracle/mc_tables.py.
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def corner(x, y, z):
    return x + 2 * y + 4 * z


def edge_id(axis, lo):
    """lo = (x,y,z) of the lower end point."""
    o = [lo[i] for i in range(3) if i != axis]
    return 4 * axis + o[0] + 2 * o[1]


def edge_between(c0, c1):
    p0 = [(c0 >> i) & 1 for i in range(3)]
    p1 = [(c1 >> i) & 1 for i in range(3)]
    diff = [i for i in range(3) if p0[i] != p1[i]]
    assert len(diff) == 1
    lo = [min(a, b) for a, b in zip(p0, p1)]
    return edge_id(diff[0], lo)


EDGE_AXIS = [e // 4 for e in range(12)]
EDGE_LO = []
for e in range(12):
    a, r = e // 4, e % 4
    o0, o1 = r & 1, r >> 1
    lo = [0, 0, 0]
    others = [i for i in range(3) if i != a]
    lo[others[0]], lo[others[1]] = o0, o1
    EDGE_LO.append(tuple(lo))
    assert edge_id(a, lo) == e


def faces():
    out = []
    for a in range(3):
        u, v = (a + 1) % 3, (a + 2) % 3
        for s in (0, 1):
            ring = [(0, 0), (1, 0), (1, 1), (0, 1)]
            if s == 0:
                ring = ring[::-1]
            cs = []
            for (pu, pv) in ring:
                p = [0, 0, 0]
                p[a], p[u], p[v] = s, pu, pv
                cs.append(corner(*p))
            out.append(cs)
    return out


FACES = faces()


def ambiguous_faces(mask):
    pos = [(mask >> c) & 1 for c in range(8)]
    return [f for f, ring in enumerate(FACES) if sum(pos[c] for c in ring) == 2 and pos[ring[0]] == pos[ring[2]]]


def case_triangles(mask, joined=()):
    """`joined`: indices of ambiguous faces whose positive corners are joined through the face."""
    pos = [(mask >> c) & 1 for c in range(8)]
    nxt = {}
    for f, ring in enumerate(FACES):
        sg = [pos[c] for c in ring]
        if sum(sg) in (0, 4):
            continue
        if f in joined:
            for k in range(4):      # cut off each negative corner
                if not sg[k]:
                    nxt[edge_between(ring[(k - 1) % 4], ring[k])] = edge_between(ring[k], ring[(k + 1) % 4])
            continue
        for i in range(4):
            # start of a maximal positive run: corner i positive, previous corner negative
            if sg[i] and not sg[(i - 1) % 4]:
                j = i
                while sg[(j + 1) % 4]:
                    j = (j + 1) % 4
                entry = edge_between(ring[(i - 1) % 4], ring[i])
                exit_ = edge_between(ring[j], ring[(j + 1) % 4])
                assert exit_ not in nxt
                nxt[exit_] = entry
    crossing = [e for e in range(12) if pos[corner(*EDGE_LO[e])] != pos[corner(*[EDGE_LO[e][i] + (1 if i == EDGE_AXIS[e] else 0) for i in range(3)])]]
    assert sorted(nxt.keys()) == sorted(crossing) and sorted(nxt.values()) == sorted(crossing)
    tris, seen = [], set()
    for e0 in crossing:
        if e0 in seen:
            continue
        loop, e = [], e0
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == e0 and len(loop) >= 3
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def main():
    rows, base, amb_mask = [], [], []
    for m in range(256):
        amb = ambiguous_faces(m)
        base.append(len(rows))
        amb_mask.append(sum(1 << f for f in amb))
        for bits in range(1 << len(amb)):
            joined = tuple(f for k, f in enumerate(amb) if (bits >> k) & 1)
            rows.append(case_triangles(m, joined))
    max_t = max(len(t) for t in rows)
    assert max_t <= 12 and len(rows) < 65536
    counts = [len(t) for t in rows]
    width = 3 * max_t
    flat = []
    for t in rows:
        row = [i for tri in t for i in tri]
        flat.append(row + [-1] * (width - len(row)))
    hdr = ['// GENERATED by tools/gen_mc_tables.py -- do not edit.',
           '// corner c = x + 2y + 4z; edge e = 4*axis + low bits of the other two coordinates of its lower end;',
           '// case = bitmask of corners with value > level.  Row of (case, decisions) = kMcRowBase[case] + sum over the set bits of',
           '// kMcAmbMask[case] (ascending face index k-th set bit) of (positive corners of that face joined ? 1 << k : 0);',
           '// a row holds kMcTriCount[row] triangles of edge labels.  kMcFaceRing[f] = the four corners of face f in cyclic order.',
           '#pragma once', '#include <cstdint>', 'namespace p2s {',
           'constexpr int kMcRows = %d, kMcMaxTris = %d;' % (len(rows), max_t),
           '__device__ __constant__ uint16_t kMcRowBase[256] = {', '    ' + ', '.join(str(b) for b in base), '};',
           '__device__ __constant__ uint8_t kMcAmbMask[256] = {', '    ' + ', '.join(str(b) for b in amb_mask), '};',
           '__device__ __constant__ uint8_t kMcFaceRing[6][4] = {', '    ' + ', '.join('{%d, %d, %d, %d}' % tuple(r) for r in FACES), '};',
           '__device__ __constant__ int8_t kMcTriTable[kMcRows][%d] = {' % width]
    for row in flat:
        hdr.append('    {' + ', '.join('%2d' % v for v in row) + '},')
    hdr += ['};', '__device__ __constant__ uint8_t kMcTriCount[kMcRows] = {',
            '    ' + ', '.join(str(c) for c in counts), '};', '}  // namespace p2s', '']
    with open(os.path.join(ROOT, 'points2surf_b200', 'csrc', 'mc_tables.cuh'), 'w') as f:
        f.write('\n'.join(hdr))
    with open(os.path.join(ROOT, 'oracle', 'mc_tables.py'), 'w') as f:
        f.write('"""GENERATED by tools/gen_mc_tables.py -- do not edit.  TEST INFRASTRUCTURE (oracle side)."""\n')
        f.write('import numpy as np\n\nROW_BASE = np.array([' + ', '.join(str(b) for b in base) + '], dtype=np.int32)\n')
        f.write('AMB_MASK = np.array([' + ', '.join(str(b) for b in amb_mask) + '], dtype=np.int32)\n')
        f.write('FACE_RING = np.array(' + repr([list(r) for r in FACES]) + ', dtype=np.int32)\n')
        f.write('TRI_TABLE = np.array([\n')
        for row in flat:
            f.write('    [' + ', '.join('%2d' % v for v in row) + '],\n')
        f.write('], dtype=np.int8)\nTRI_COUNT = np.array([' + ', '.join(str(c) for c in counts) + '], dtype=np.int32)\n')
    print('rows', len(rows), 'max tris', max_t, 'total tris', sum(counts), 'table bytes', len(rows) * width)


if __name__ == '__main__':
    main()
