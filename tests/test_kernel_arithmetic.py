"""CPU checks of the integer / bit tricks the byte kernels rely on, exhaustively where the domain is small.  They restate the
arithmetic of the CUDA code (file and construct named per test) in NumPy / pure Python; the kernels themselves are checked
against the oracle on the GPU (tests/test_gpu_parity.py)."""
import random

import numpy as np


def test_packed_byte_vote_equals_the_scalar_rule_for_every_byte_and_threshold():
    # csrc/volume.cu, row-vector path of propagate_kernel: four box sums of biased signs per 32-bit word (bytes 0..250,
    # bias 125); vote = 0 if |n| < thr or n == 0 else sign(n), n = byte - 125, evaluated with carry tricks on 16-bit lanes.
    def packed_vote(word, ithr):
        T = min(max(ithr, 1), 126)
        cpos = ((0x100 - (125 + T)) * 0x00010001) & 0xffffffff
        cneg = ((0x100 + 125 - T) * 0x00010001) & 0xffffffff
        e, o = word & 0x00ff00ff, (word >> 8) & 0x00ff00ff
        pe, po = ((e + cpos) >> 8) & 0x00010001, ((o + cpos) >> 8) & 0x00010001
        ne, no = ((cneg - e) >> 8) & 0x00010001, ((cneg - o) >> 8) & 0x00010001
        nz = 4 - bin(pe | ne).count('1') - bin(po | no).count('1')
        ve, vo = pe | (ne * 3), po | (no * 3)
        return (ve | (vo << 8) | 0x04040404) & 0xffffffff, nz

    def scalar_vote(byte, thr):
        n = byte - 125
        if abs(n) < thr or n == 0:
            return 0
        return 1 if n > 0 else 3          # 2-bit two's complement sign

    rng = random.Random(0)
    for thr in [0.0, 0.5, 1.0, 5.0, 12.5, 13.0, 26.0, 40.0, 125.0, 126.0, 200.0]:
        ithr = int(np.ceil(thr)) if thr > 0 else 0
        # every byte value in every lane (the other lanes random)
        for lane in range(4):
            for b in range(251):
                others = [rng.randrange(251) for _ in range(4)]
                others[lane] = b
                word = others[0] | (others[1] << 8) | (others[2] << 16) | (others[3] << 24)
                cand, nz = packed_vote(word, ithr)
                want = [scalar_vote(x, thr) for x in others]
                got = [(cand >> (8 * q)) & 3 for q in range(4)]
                assert got == want, (thr, word)
                assert all(((cand >> (8 * q)) & 0xfc) == 4 for q in range(4))      # the "unknown at the start" flag travels
                assert nz == sum(1 for x in want if x == 0)


def test_biased_packed_box_sums_never_carry_and_match_the_signed_sums():
    # csrc/volume.cu: signs are held as sign + 1 in {0,1,2}; z sums by funnel shifts, y / x sums by (sliding) word adds.
    # 5^3 taps * 2 = 250 <= 255, so no byte carries into its neighbour; bias 125 restores the signed sum.
    rng = np.random.RandomState(1)
    s = rng.randint(-1, 2, size=(12, 12, 40)).astype(np.int32)          # tile + halo of signs
    biased = (s + 1).astype(np.uint32)
    words = (biased[..., 0::4] | (biased[..., 1::4] << 8) | (biased[..., 2::4] << 16) | (biased[..., 3::4] << 24)).astype(np.uint64)   # [12,12,10]

    def funnel(lo, hi, shift):              # __funnelshift_r(lo, hi, shift): low 32 bits of (hi:lo) >> shift
        return (((hi << np.uint64(32)) | lo) >> np.uint64(shift)) & np.uint64(0xffffffff)

    c = words
    z = np.zeros((12, 12, 8), np.uint64)
    for j in range(8):                       # output word j: window starts at bytes 4j+2 .. 4j+6 of the row (halo word first)
        z[..., j] = (funnel(c[..., j], c[..., j + 1], 16) + funnel(c[..., j], c[..., j + 1], 24) + c[..., j + 1] +
                     funnel(c[..., j + 1], c[..., j + 2], 8) + funnel(c[..., j + 1], c[..., j + 2], 16))
    y = np.zeros((12, 8, 8), np.uint64)
    for x in range(12):                      # sliding window down y: acc - v[y-1] + v[y+4] (bytes never borrow)
        acc = z[x, 0] + z[x, 1] + z[x, 2] + z[x, 3] + z[x, 4]
        y[x, 0] = acc
        for yy in range(1, 8):
            acc = acc - z[x, yy - 1] + z[x, yy + 4]
            y[x, yy] = acc
    out = np.zeros((8, 8, 8), np.uint64)
    for x in range(8):
        out[x] = y[x] + y[x + 1] + y[x + 2] + y[x + 3] + y[x + 4]
    assert (out >> np.uint64(32)).max() == 0                                   # nothing left the 32-bit word
    got = np.stack([(out >> np.uint64(8 * q)) & np.uint64(0xff) for q in range(4)], axis=-1).reshape(8, 8, 32).astype(np.int64) - 125
    want = np.zeros((8, 8, 32), np.int64)
    for x in range(8):
        for yy in range(8):
            for zz in range(32):
                want[x, yy, zz] = s[x:x + 5, yy:yy + 5, zz + 2:zz + 7].sum()   # interior voxel zz <-> row byte zz + 4, taps -2..+2
    assert np.array_equal(got, want)


def test_hybrid_bitonic_network_sorts():
    # csrc/assemble.cu sort_candidates: two elements per thread in registers, strides 1 (in-thread) and 2..32 (warp shuffles),
    # strides >= 64 through shared memory -- the same compare-exchange network as the plain bitonic sort, in another order
    def sort_net(vals, P):
        a = list(vals)

        def ce_smem(size, stride):
            for t in range(P // 2):
                lo = 2 * t - (t & (stride - 1))
                hi = lo + stride
                if (a[lo] > a[hi]) == ((lo & size) == 0):
                    a[lo], a[hi] = a[hi], a[lo]

        def reg(size):
            stride = min(size >> 1, 32)
            while stride >= 2:
                m, new = stride >> 1, list(a)
                for t in range(P // 2):
                    keep_min = (((2 * t) & stride) == 0) == (((2 * t) & size) == 0)
                    for b in range(2):
                        k, ok = a[2 * t + b], a[2 * (t ^ m) + b]
                        if (ok < k) == keep_min:
                            new[2 * t + b] = ok
                a[:] = new
                stride >>= 1
            for t in range(P // 2):
                if (a[2 * t] > a[2 * t + 1]) == (((2 * t) & size) == 0):
                    a[2 * t], a[2 * t + 1] = a[2 * t + 1], a[2 * t]

        size = 2
        while size <= 64 and size <= P:
            reg(size)
            size <<= 1
        size = 128
        while size <= P:
            stride = size >> 1
            while stride >= 64:
                ce_smem(size, stride)
                stride >>= 1
            reg(size)
            size <<= 1
        return a

    rng = random.Random(3)
    for P in (64, 128, 256, 512):
        for _ in range(20):
            n = rng.randint(1, P)
            v = [(rng.randint(0, 40), i) for i in range(n)] + [(10 ** 9, 2 ** 31 - 1)] * (P - n)      # (key, id): strict order + padding
            rng.shuffle(v)
            assert sort_net(v, P) == sorted(v)


def test_cell_sampler_slot_arithmetic_gives_every_point_its_share():
    # csrc/assemble.cu subsample_cells_kernel: ONE integer x drawn from [0, sum_c count_c * wq_c) selects cell and point
    # (point = offset // wq_c); every point of cell c owns exactly wq_c slots, so P(point) = wq_c / total and the acceptance
    # w_i / (wq_c / 65535) makes P(propose and accept i) proportional to w_i.
    rng = np.random.RandomState(5)
    counts = rng.randint(0, 6, size=40)
    wq = np.where(counts > 0, rng.randint(3278, 65536, size=40), 0)
    prefix = np.cumsum(counts * wq)
    total = int(prefix[-1])
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    owned = np.zeros(int(counts.sum()), np.int64)
    xs = np.arange(total)
    cells = np.searchsorted(prefix, xs, side='right')                  # smallest c with prefix[c] > x (the kernel's search)
    off = xs - np.where(cells > 0, prefix[cells - 1], 0)
    pts = start[cells] + off // wq[cells]
    np.add.at(owned, pts, 1)
    want = np.repeat(wq, counts)
    assert np.array_equal(owned, want)
    # 40 random bits -> slot: floor(x40 * total / 2^40) is in range and monotone
    x40 = rng.randint(0, 2 ** 40, size=1000, dtype=np.int64)
    slot = (x40.astype(object) * total) >> 40
    assert all(0 <= int(v) < total for v in slot)
