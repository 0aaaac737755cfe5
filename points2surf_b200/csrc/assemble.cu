// K2 / K3: per-query patch assembly -- PointcloudPatchDataset.__getitem__ (source/data_loader.py:322-421).
//   K2  exact kNN (k smallest float64 distances on float32 coordinates, i.e. scipy cKDTree semantics,
//       source/base/point_cloud.py:174-175), patch radius and patch-space normalisation in float32
//       exactly like NumPy (source/base/utils.py:62-69,80-88).
//   K3  global sub-sample (source/base/utils.py:196-227): uniform with replacement, or distance-weighted
//       without replacement via exponential clocks (Efraimidis-Spirakis), which realises the same
//       successive-sampling law as RandomState.choice(replace=False, p).
// One CTA per query; the cloud (N*12 B, L1/L2-resident) is streamed twice per selection:
// a histogram pass over the top bits of the (monotone) key, then a collect pass.  All byte/compare work.
#include "common.cuh"

namespace p2s {

namespace {

constexpr int kThreads = 256;
constexpr int kBins = 2048;
constexpr int kCap = 1024;  // boundary-bin candidates that are sorted exactly (sub-sampler, kNN up to 512 neighbours)
constexpr int kCapBig = 2048;   // kNN with 513..1536 neighbours (large_kNN: 1200) and ball-query patches

// ---- key helpers: non-negative doubles order like their bit patterns ----
__device__ __forceinline__ unsigned long long dkey(double v) { return (unsigned long long)__double_as_longlong(v); }
// level-0 bin: sign+exponent+4 mantissa bits, rebased so that 2^-100 .. 2^27 maps to 0..2047
__device__ __forceinline__ int bin0(unsigned long long key) {
    long long b = (long long)(key >> 48) - ((1023 - 100) << 4);
    return (int)(b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b));
}

template <int CAP>
struct SelectSmemT {
    unsigned hist[kBins];
    unsigned long long cand_key[CAP];
    int cand_id[CAP];
    int bin_sel[4];       // selected bin per level
    unsigned below;       // number of keys strictly below the boundary bin
    unsigned n_direct;    // slots used by "surely in" members
    unsigned n_cand;      // boundary candidates collected
    int levels;           // refinement levels used (1..3)
};
using SelectSmem = SelectSmemT<kCap>;

// does `key` fall in the boundary bin chain selected so far (levels [0, upto))?
template <int CAP>
__device__ __forceinline__ int chain_cmp(const SelectSmemT<CAP>& s, unsigned long long key, int upto) {
    // returns -1 if key sorts below the chain, 0 if inside, +1 above
    int b = bin0(key);
    if (b != s.bin_sel[0]) return b < s.bin_sel[0] ? -1 : 1;
    for (int l = 1; l < upto; ++l) {
        int bl = (int)((key >> (48 - 11 * l)) & 0x7FF);
        if (bl != s.bin_sel[l]) return bl < s.bin_sel[l] ? -1 : 1;
    }
    return 0;
}

// Block-wide: find the bin chain that contains the k-th smallest key.  KeyFn(i) -> key of item i.
template <int CAP, class KeyFn>
__device__ void find_boundary(SelectSmemT<CAP>& s, int N, int k, int cand_cap, KeyFn keyfn) {
    const int tid = threadIdx.x;
    if (tid == 0) { s.below = 0; s.levels = 0; }
    for (int level = 0; level < 3; ++level) {
        for (int i = tid; i < kBins; i += kThreads) s.hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < N; i += kThreads) {
            unsigned long long key = keyfn(i);
            if (level == 0) atomicAdd(&s.hist[bin0(key)], 1u);
            else if (chain_cmp(s, key, level) == 0) atomicAdd(&s.hist[(int)((key >> (48 - 11 * level)) & 0x7FF)], 1u);
        }
        __syncthreads();
        if (tid < 32) {  // warp 0: locate the bin where the running count crosses k
            unsigned need = (unsigned)k - s.below;  // rank inside the current chain, 1-based
            unsigned run = 0;
            int found = -1;
            unsigned below_add = 0;
            for (int base = 0; base < kBins && found < 0; base += 32) {
                unsigned c = s.hist[base + tid];
                unsigned incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                unsigned excl = incl - c;
                unsigned hit = __ballot_sync(0xffffffffu, run + incl >= need);
                if (hit) {
                    int lane = __ffs(hit) - 1;
                    found = base + lane;
                    below_add = run + __shfl_sync(0xffffffffu, excl, lane);
                }
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (tid == 0) {
                s.bin_sel[level] = found;
                s.below += below_add;
                s.levels = level + 1;
            }
        }
        __syncthreads();
        if (s.hist[s.bin_sel[level]] <= (unsigned)cand_cap) break;  // uniform: same smem value for all threads
        __syncthreads();
    }
}

// bitonic sort of (key, id) ascending over the first P = pow2ceil(n) slots (slots >= n are padded with +inf keys).
// The total order (key, id) is strict, so the result does not depend on the network used.
// P <= 512: two elements per thread live in registers; compare-exchanges at element strides 1 (in-thread) and 2..32 (warp
// shuffles) need no barrier, only the strides >= 64 go through shared memory: 10 block barriers per sort instead of 45
// (the barriers, not the instructions, were what the kNN kernel spent its time on).
template <int CAP>
__device__ void sort_candidates(SelectSmemT<CAP>& s, int n) {
    const int tid = threadIdx.x;
    int P = 64;
    while (P < n) P <<= 1;                       // n <= CAP (checked by the callers), so P <= CAP
    for (int i = tid; i < P; i += kThreads)
        if (i >= n) { s.cand_key[i] = ~0ull; s.cand_id[i] = 0x7fffffff; }
    __syncthreads();
    auto smem_stage = [&](int size, int stride) {
        for (int t = tid; t < P / 2; t += kThreads) {
            int lo = 2 * t - (t & (stride - 1));
            int hi = lo + stride;
            bool up = ((lo & size) == 0);
            unsigned long long ka = s.cand_key[lo], kb = s.cand_key[hi];
            int ia = s.cand_id[lo], ib = s.cand_id[hi];
            bool gt = (ka > kb) || (ka == kb && ia > ib);
            if (gt == up) { s.cand_key[lo] = kb; s.cand_key[hi] = ka; s.cand_id[lo] = ib; s.cand_id[hi] = ia; }
        }
        __syncthreads();
    };
    if (P <= 2 * kThreads) {
        static_assert(kThreads == 256, "register sort assumes 256 threads (2 elements per thread at P = 512)");
        const bool active = 2 * tid < P;          // warp-uniform: P is a multiple of 64
        unsigned long long k0 = 0, k1 = 0;
        int i0 = 0, i1 = 0;
        auto load = [&] { if (active) { k0 = s.cand_key[2 * tid]; k1 = s.cand_key[2 * tid + 1]; i0 = s.cand_id[2 * tid]; i1 = s.cand_id[2 * tid + 1]; } };
        auto store = [&] { if (active) { s.cand_key[2 * tid] = k0; s.cand_key[2 * tid + 1] = k1; s.cand_id[2 * tid] = i0; s.cand_id[2 * tid + 1] = i1; } };
        // strides min(size / 2, 32) .. 1 of the merge of `size`, in registers
        auto reg_stages = [&](int size) {
            if (!active) return;
            for (int stride = (size >> 1) < 32 ? (size >> 1) : 32; stride >= 2; stride >>= 1) {
                const int m = stride >> 1;                              // lane mask of the partner thread
                const bool up = (((2 * tid) & size) == 0);
                const bool keep_min = (((2 * tid) & stride) == 0) == up;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    unsigned long long& k = b ? k1 : k0;
                    int& id = b ? i1 : i0;
                    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, m);
                    const int oi = __shfl_xor_sync(0xffffffffu, id, m);
                    const bool other_less = (ok < k) || (ok == k && oi < id);
                    if (other_less == keep_min) { k = ok; id = oi; }
                }
            }
            {   // stride 1: the thread's own pair
                const bool up = (((2 * tid) & size) == 0);
                const bool gt = (k0 > k1) || (k0 == k1 && i0 > i1);
                if (gt == up) { const unsigned long long tk = k0; k0 = k1; k1 = tk; const int ti = i0; i0 = i1; i1 = ti; }
            }
        };
        load();
        for (int size = 2; size <= 64 && size <= P; size <<= 1) reg_stages(size);
        store();
        __syncthreads();
        for (int size = 128; size <= P; size <<= 1) {
            for (int stride = size >> 1; stride >= 64; stride >>= 1) smem_stage(size, stride);
            load();
            reg_stages(size);
            store();
            __syncthreads();
        }
        return;
    }
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) smem_stage(size, stride);
}

__device__ __forceinline__ double dist2_f64(const float* __restrict__ pts, int i, double qx, double qy, double qz) {
    // cKDTree: sum over dimensions of (x-y)^2, float64, sequential, no FMA contraction
    double dx = (double)pts[i * 3 + 0] - qx, dy = (double)pts[i * 3 + 1] - qy, dz = (double)pts[i * 3 + 2] - qz;
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// NumPy float32: np.linalg.norm(q - p) = sqrt((dx*dx + dy*dy) + dz*dz), every op rounded to float32
__device__ __forceinline__ float norm_f32(float dx, float dy, float dz) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// ------------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------------
// One CTA walks kRun consecutive queries of the ordered list.  Consecutive voxel centres are close, so the exact
// k-th distance of query j gives a *guaranteed* bound for query j+1 (triangle inequality:
// kth(j+1) <= kth(j) + |q(j+1) - q(j)|): one pass collects every point within that bound (typically 1.1-1.2 k
// candidates) and an exact sort finishes the job -- no histogram.  The first query of a run, big jumps between
// queries and overflowing candidate lists fall back to the histogram selection.
constexpr int kRun = 8;        // shortest run; the launcher lengthens it so that the grid is one full wave (see knn_patch)

template <int CAP>
__global__ void __launch_bounds__(kThreads)
knn_patch_kernel(const float* __restrict__ pts, int N, const float* __restrict__ queries, int64_t Q, int k,
                 int32_t* __restrict__ ids_out, float* __restrict__ patch_out, float* __restrict__ radius_out,
                 int* __restrict__ err_flag, int run) {
    constexpr int kCap = CAP;                     // shadows the file-level constant inside this kernel
    __shared__ SelectSmemT<CAP> s;
    __shared__ float red[kThreads / 32];
    __shared__ float s_radius;
    __shared__ double s_kth;          // exact k-th squared distance of the previous query (0 = unknown)
    const int tid = threadIdx.x;
    const int64_t q_begin = (int64_t)blockIdx.x * run;
    const int64_t q_end = (q_begin + run < Q) ? (q_begin + run) : Q;
    if (tid == 0) s_kth = 0.0;
    double pqx = 0.0, pqy = 0.0, pqz = 0.0;
    for (int64_t q = q_begin; q < q_end; ++q) {
        __syncthreads();
        const float qxf = queries[q * 3 + 0], qyf = queries[q * 3 + 1], qzf = queries[q * 3 + 2];
        const double qx = qxf, qy = qyf, qz = qzf;
        auto keyfn = [&](int i) { return dkey(dist2_f64(pts, i, qx, qy, qz)); };
        // ---- fast path: bound from the previous query
        bool done = false;
        const double kth_prev = s_kth;
        if (kth_prev > 0.0) {
            const double ddx = qx - pqx, ddy = qy - pqy, ddz = qz - pqz;
            const double delta = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
            const double rprev = sqrt(kth_prev);
            if (delta < 0.25 * rprev) {
                const double rb = (rprev + delta) * (1.0 + 1e-12);
                const unsigned long long bound = dkey(rb * rb * (1.0 + 1e-12));
                if (tid == 0) s.n_cand = 0;
                __syncthreads();
                // float32 pre-filter (relative error of the fp32 squared distance < 1e-6): only points that can be
                // inside the bound pay for the exact float64 distance
                const float bound_f = (float)(rb * rb) * 1.00001f;
                for (int i = tid; i < N; i += kThreads) {
                    const float fx = pts[i * 3 + 0] - qxf, fy = pts[i * 3 + 1] - qyf, fz = pts[i * 3 + 2] - qzf;
                    if (fmaf(fx, fx, fmaf(fy, fy, fz * fz)) > bound_f) continue;
                    unsigned long long key = keyfn(i);
                    if (key <= bound) {
                        unsigned slot = atomicAdd(&s.n_cand, 1u);
                        if (slot < (unsigned)kCap) { s.cand_key[slot] = key; s.cand_id[slot] = i; }
                    }
                }
                __syncthreads();
                const unsigned n = s.n_cand;
                if (n >= (unsigned)k && n <= (unsigned)kCap) {
                    sort_candidates(s, (int)n);
                    done = true;
                }
            }
        }
        if (!done) {
            // ---- histogram selection
            __syncthreads();
            find_boundary(s, N, k, kCap - (k > 512 ? k : 512), keyfn);
            if (tid == 0) { s.n_direct = 0; s.n_cand = 0; }
            __syncthreads();
            const int levels = s.levels;
            // collect: keys below the boundary chain are members; keys inside it are candidates
            for (int i = tid; i < N; i += kThreads) {
                unsigned long long key = keyfn(i);
                int c = chain_cmp(s, key, levels);
                if (c < 0) {
                    // members go to the tail of the candidate arrays so one sort orders everything
                    unsigned slot = atomicAdd(&s.n_direct, 1u);
                    if (slot < (unsigned)kCap) { s.cand_key[kCap - 1 - slot] = key; s.cand_id[kCap - 1 - slot] = i; }
                } else if (c == 0) {
                    unsigned slot = atomicAdd(&s.n_cand, 1u);
                    if (slot < (unsigned)kCap) { s.cand_key[slot] = key; s.cand_id[slot] = i; }
                }
            }
            __syncthreads();
            const unsigned n_direct = s.n_direct, n_cand = s.n_cand;
            if (n_direct + n_cand > (unsigned)kCap || n_direct != s.below) {
                // more than kCap points tie into the boundary bin even after 3 refinement levels (degenerate cloud)
                if (tid == 0) atomicExch(err_flag, 1);
                return;
            }
            // compact: move the members right behind the candidates, then sort everything exactly
            unsigned long long mk[(kCap + kThreads - 1) / kThreads];
            int mi[(kCap + kThreads - 1) / kThreads];
            int cnt = 0;
            for (unsigned j = tid; j < n_direct; j += kThreads) { mk[cnt] = s.cand_key[kCap - 1 - j]; mi[cnt] = s.cand_id[kCap - 1 - j]; ++cnt; }
            __syncthreads();
            cnt = 0;
            for (unsigned j = tid; j < n_direct; j += kThreads) { s.cand_key[n_cand + j] = mk[cnt]; s.cand_id[n_cand + j] = mi[cnt]; ++cnt; }
            __syncthreads();
            sort_candidates(s, (int)(n_direct + n_cand));
        }
        // exact k-th squared distance -> bound for the next query of the run
        if (tid == 0) s_kth = __longlong_as_double((long long)s.cand_key[k - 1]);
        pqx = qx; pqy = qy; pqz = qz;

        // radius = max float32 norm over the k neighbours (utils.get_patch_radii)
        float r = 0.f;
        for (int j = tid; j < k; j += kThreads) {
            int id = s.cand_id[j];
            r = fmaxf(r, norm_f32(__fsub_rn(qxf, pts[id * 3 + 0]), __fsub_rn(qyf, pts[id * 3 + 1]), __fsub_rn(qzf, pts[id * 3 + 2])));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
        if ((tid & 31) == 0) red[tid >> 5] = r;
        __syncthreads();
        if (tid == 0) {
            float m = red[0];
            for (int w = 1; w < kThreads / 32; ++w) m = fmaxf(m, red[w]);
            s_radius = m;
            radius_out[q] = m;
        }
        __syncthreads();
        const float radius = s_radius;
        for (int j = tid; j < k; j += kThreads) {
            int id = s.cand_id[j];
            if (ids_out) ids_out[q * k + j] = id;
            float* o = patch_out + (q * k + j) * 3;
            // model_space_to_patch_space: (p - q) / r in float32
            o[0] = __fdiv_rn(__fsub_rn(pts[id * 3 + 0], qxf), radius);
            o[1] = __fdiv_rn(__fsub_rn(pts[id * 3 + 1], qyf), radius);
            o[2] = __fdiv_rn(__fsub_rn(pts[id * 3 + 2], qzf), radius);
        }
    }
}

__device__ __forceinline__ float u01_open(uint32_t x) {  // (0,1]
    return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// ------------------------------------------------------------------------------------------------
// K2b: ball-query patches (radius ablations) -- point_cloud.get_patch_kdtree with patch_radius > 0
// (source/base/point_cloud.py:176-192) + the padding rule of PointcloudPatchDataset.__getitem__
// (source/data_loader.py:340-350): all points with float64 distance <= r (cKDTree.query_ball_point semantics on
// float32 coordinates); more than k of them -> a uniformly random k-subset without replacement (rng.choice in the
// reference; here the k smallest Philox clocks, same law, different stream); fewer -> padded with the query point itself
// (patch-space origin, id 0).  Normalisation by the FIXED radius in float32.  One CTA per query.
// Output order: ascending point id when nothing is dropped (the reference's order is the kd-tree traversal order; the
// network max-pools over the patch, so order carries no information).
constexpr int kBallMaxK = 1536;

__global__ void __launch_bounds__(kThreads)
ball_patch_kernel(const float* __restrict__ pts, int N, const float* __restrict__ queries, int64_t qbase,
                  const int32_t* __restrict__ qidx, int k, double r2, float rf, uint64_t seed,
                  int32_t* __restrict__ ids_out, float* __restrict__ patch_out, float* __restrict__ radius_out,
                  int32_t* __restrict__ count_out, int* __restrict__ err_flag) {
    constexpr int CAP = kCapBig;
    __shared__ SelectSmemT<CAP> s;
    __shared__ int sel[kBallMaxK];
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const float qxf = queries[q * 3 + 0], qyf = queries[q * 3 + 1], qzf = queries[q * 3 + 2];
    const double qx = qxf, qy = qyf, qz = qzf;
    const uint64_t qi = (uint64_t)(qbase + (qidx ? (int64_t)qidx[q] : q));
    const float bound_f = (float)r2 * 1.00001f + 1e-30f;
    auto in_ball = [&](int i) {
        const float fx = pts[i * 3 + 0] - qxf, fy = pts[i * 3 + 1] - qyf, fz = pts[i * 3 + 2] - qzf;
        if (fmaf(fx, fx, fmaf(fy, fy, fz * fz)) > bound_f) return false;       // fp32 pre-filter, exact test below
        return dist2_f64(pts, i, qx, qy, qz) <= r2;
    };
    auto clock_key = [&](int i) {     // Exp(1) clock of point i for this query
        uint32_t r[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)(i >> 2), 0x3c6ef372u, r);
        return dkey((double)(-__logf(u01_open(r[i & 3]))));
    };
    if (tid == 0) { s.n_cand = 0; s.n_direct = 0; }
    __syncthreads();
    for (int i = tid; i < N; i += kThreads) {
        if (!in_ball(i)) continue;
        const unsigned slot = atomicAdd(&s.n_cand, 1u);
        if (slot < (unsigned)CAP) { s.cand_key[slot] = (unsigned long long)i; s.cand_id[slot] = i; }
    }
    __syncthreads();
    const int count = (int)s.n_cand;
    if (count_out && tid == 0) count_out[q] = count;
    int n_sel = count < k ? count : k;
    if (count <= k) {
        sort_candidates(s, count);                            // ascending id
        for (int j = tid; j < n_sel; j += kThreads) sel[j] = s.cand_id[j];
    } else if (count <= CAP) {
        for (int j = tid; j < count; j += kThreads) s.cand_key[j] = clock_key(s.cand_id[j]);
        __syncthreads();
        sort_candidates(s, count);
        for (int j = tid; j < k; j += kThreads) sel[j] = s.cand_id[j];
    } else {
        // more points in the ball than the candidate buffer holds: histogram selection of the k smallest clocks
        __syncthreads();
        auto keyfn = [&](int i) { return in_ball(i) ? clock_key(i) : ~0ull; };
        find_boundary(s, N, k, CAP, keyfn);
        if (tid == 0) { s.n_direct = 0; s.n_cand = 0; }
        __syncthreads();
        const int levels = s.levels;
        for (int i = tid; i < N; i += kThreads) {
            const unsigned long long key = keyfn(i);
            if (key == ~0ull) continue;
            const int c = chain_cmp(s, key, levels);
            if (c < 0) {
                const unsigned slot = atomicAdd(&s.n_direct, 1u);
                if (slot < (unsigned)k) sel[slot] = i;
            } else if (c == 0) {
                const unsigned slot = atomicAdd(&s.n_cand, 1u);
                if (slot < (unsigned)CAP) { s.cand_key[slot] = key; s.cand_id[slot] = i; }
            }
        }
        __syncthreads();
        const unsigned n_direct = s.n_direct, n_cand = s.n_cand;
        if (n_cand > (unsigned)CAP || n_direct != s.below || n_direct + n_cand < (unsigned)k) {
            if (tid == 0) atomicExch(err_flag, 3);
            return;
        }
        sort_candidates(s, (int)n_cand);
        for (unsigned j = tid; n_direct + j < (unsigned)k; j += kThreads) sel[n_direct + j] = s.cand_id[j];
    }
    __syncthreads();
    if (tid == 0) radius_out[q] = rf;
    for (int j = tid; j < k; j += kThreads) {
        float* o = patch_out + (q * k + j) * 3;
        if (j < n_sel) {
            const int id = sel[j];
            if (ids_out) ids_out[q * k + j] = id;
            o[0] = __fdiv_rn(__fsub_rn(pts[id * 3 + 0], qxf), rf);
            o[1] = __fdiv_rn(__fsub_rn(pts[id * 3 + 1], qyf), rf);
            o[2] = __fdiv_rn(__fsub_rn(pts[id * 3 + 2], qzf), rf);
        } else {
            if (ids_out) ids_out[q * k + j] = 0;              // -1 -> 0, coordinates <- query point (data_loader.py:341-345)
            o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3
// ------------------------------------------------------------------------------------------------

__global__ void subsample_uniform_kernel(int N, int64_t Q, int64_t qbase, const int32_t* __restrict__ qidx, int S, uint64_t seed, int32_t* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int quads = (S + 3) / 4;
    if (t >= Q * quads) return;
    int64_t q = t / quads;
    int j4 = (int)(t % quads);
    uint32_t r[4];
    uint64_t qi = (uint64_t)(qbase + (qidx ? (int64_t)qidx[q] : q));
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)j4, 0x5ab5a3e1u, r);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int j = j4 * 4 + e;
        if (j < S) out[q * S + j] = (int32_t)(((uint64_t)r[e] * (uint64_t)N) >> 32);
    }
}

// CACHE: per-point values live in dynamic shared memory (N floats) so that the distance, the Philox block (shared
// by 4 consecutive points) and the logarithm are evaluated once per point.  With the cache the selection needs no
// histogram: the clock of point i is Exp(1)/w_i, so the expected number of clocks below t is
//   C(t) = sum_i (1 - exp(-w_i t)) = t*S1 - t^2*S2/2 + t^3*S3/6 - ...        (w_i t <~ 0.2 for S/N ~ 0.1)
// Clocks <= t_lo (C = S - 5 sqrt(S)) are members, clocks in (t_lo, t_hi] (C = S + 5 sqrt(S)) are sorted exactly and
// the smallest S - n_members of them complete the draw -- the same S smallest clocks as the histogram selection,
// which takes over whenever the realised counts do not bracket S.
template <bool CACHE>
__global__ void __launch_bounds__(kThreads)
subsample_weighted_kernel(const float* __restrict__ pts, int N, const float* __restrict__ queries,
                          int64_t qbase, const int32_t* __restrict__ qidx, int S, uint64_t seed, int32_t* __restrict__ out, int* __restrict__ err_flag) {
    extern __shared__ float s_val[];     // [N] when CACHE: distance, then weight
    __shared__ SelectSmem s;
    __shared__ float red[3][kThreads / 32];
    __shared__ float s_dmax, s_tlo, s_thi;
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const float qx = queries[q * 3 + 0], qy = queries[q * 3 + 1], qz = queries[q * 3 + 2];
    // dist_prob (utils.py:200-208): float32 like NumPy
    float dmax = 0.f;
    for (int i = tid; i < N; i += kThreads) {
        float d = norm_f32(__fsub_rn(qx, pts[i * 3 + 0]), __fsub_rn(qy, pts[i * 3 + 1]), __fsub_rn(qz, pts[i * 3 + 2]));
        if (CACHE) s_val[i] = d;
        dmax = fmaxf(dmax, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if ((tid & 31) == 0) red[0][tid >> 5] = dmax;
    __syncthreads();
    if (tid == 0) {
        float m = red[0][0];
        for (int w = 1; w < kThreads / 32; ++w) m = fmaxf(m, red[0][w]);
        s_dmax = m;
    }
    __syncthreads();
    dmax = s_dmax;
    const uint64_t qi = (uint64_t)(qbase + (qidx ? (int64_t)qidx[q] : q));
    auto weight_of = [&](float d) {
        float dn = __fdiv_rn(d, dmax);
        float w = __fsub_rn(1.0f, __fmul_rn(1.5f, dn));
        return fminf(fmaxf(w, 0.05f), 1.0f);
    };
    // exponential clock with rate w: the S earliest arrivals are a draw without replacement with p ~ w
    auto clock_of = [&](float w, uint32_t rnd) { return __fdividef(-__logf(u01_open(rnd)), w); };
    auto philox_block = [&](int i4, uint32_t (&r)[4]) {
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)i4, 0x77f1e2d3u, r);
    };
    if (CACHE) {
        // ---- weights + their first three power sums
        float m1 = 0.f, m2 = 0.f, m3 = 0.f;
        for (int i = tid; i < N; i += kThreads) {
            const float w = weight_of(s_val[i]);
            s_val[i] = w;
            m1 += w; m2 += w * w; m3 += w * w * w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, o);
            m2 += __shfl_xor_sync(0xffffffffu, m2, o);
            m3 += __shfl_xor_sync(0xffffffffu, m3, o);
        }
        if ((tid & 31) == 0) { red[0][tid >> 5] = m1; red[1][tid >> 5] = m2; red[2][tid >> 5] = m3; }
        __syncthreads();
        if (tid == 0) {
            float S1 = 0.f, S2 = 0.f, S3 = 0.f;
            for (int w = 0; w < kThreads / 32; ++w) { S1 += red[0][w]; S2 += red[1][w]; S3 += red[2][w]; }
            auto solve = [&](float target) {
                if (target <= 0.f) return 0.f;
                float t = target / S1;
                for (int it = 0; it < 6; ++it) {
                    float f = t * (S1 - t * (0.5f * S2 - t * (S3 * (1.0f / 6.0f)))) - target;
                    float fp = S1 - t * (S2 - 0.5f * t * S3);
                    if (!(fp > 0.f)) break;
                    t -= f / fp;
                }
                return t > 0.f ? t : 0.f;
            };
            const float sig = 5.0f * sqrtf((float)S);
            s_tlo = solve((float)S - sig);
            s_thi = solve((float)S + sig);
            s.n_direct = 0; s.n_cand = 0;
        }
        __syncthreads();
        const float tlo = s_tlo, thi = s_thi;
        // ---- clocks (one Philox block per 4 consecutive points) and classification
        for (int i4 = tid; i4 * 4 < N; i4 += kThreads) {
            uint32_t r[4];
            philox_block(i4, r);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i4 * 4 + e;
                if (i >= N) break;
                const float c = clock_of(s_val[i], r[e]);
                if (c <= tlo) {
                    unsigned slot = atomicAdd(&s.n_direct, 1u);
                    if (slot < (unsigned)S) out[q * S + slot] = i;
                } else if (c <= thi) {
                    unsigned slot = atomicAdd(&s.n_cand, 1u);
                    if (slot < (unsigned)kCap) { s.cand_key[slot] = dkey((double)c); s.cand_id[slot] = i; }
                }
            }
        }
        __syncthreads();
        const unsigned n_direct = s.n_direct, n_cand = s.n_cand;
        if (n_direct <= (unsigned)S && n_direct + n_cand >= (unsigned)S && n_cand <= (unsigned)kCap) {
            sort_candidates(s, (int)n_cand);
            for (unsigned j = tid; n_direct + j < (unsigned)S; j += kThreads) out[q * S + n_direct + j] = s.cand_id[j];
            return;
        }
        __syncthreads();     // counts did not bracket S: histogram selection below (cache holds the weights)
    }
    auto keyfn = [&](int i) {
        uint32_t r[4];
        philox_block(i >> 2, r);
        const float w = CACHE ? s_val[i]
                              : weight_of(norm_f32(__fsub_rn(qx, pts[i * 3 + 0]), __fsub_rn(qy, pts[i * 3 + 1]), __fsub_rn(qz, pts[i * 3 + 2])));
        return dkey((double)clock_of(w, r[i & 3]));
    };
    find_boundary(s, N, S, kCap, keyfn);
    if (tid == 0) { s.n_direct = 0; s.n_cand = 0; }
    __syncthreads();
    const int levels = s.levels;
    for (int i = tid; i < N; i += kThreads) {
        unsigned long long key = keyfn(i);
        int c = chain_cmp(s, key, levels);
        if (c < 0) {
            unsigned slot = atomicAdd(&s.n_direct, 1u);
            if (slot < (unsigned)S) out[q * S + slot] = i;
        } else if (c == 0) {
            unsigned slot = atomicAdd(&s.n_cand, 1u);
            if (slot < (unsigned)kCap) { s.cand_key[slot] = key; s.cand_id[slot] = i; }
        }
    }
    __syncthreads();
    const unsigned n_direct = s.n_direct, n_cand = s.n_cand;
    if (n_cand > (unsigned)kCap || n_direct != s.below || n_direct + n_cand < (unsigned)S) {
        if (tid == 0) atomicExch(err_flag, 2);
        return;
    }
    sort_candidates(s, (int)n_cand);
    for (unsigned j = tid; n_direct + j < (unsigned)S; j += kThreads) out[q * S + n_direct + j] = s.cand_id[j];
}

// K3 by rejection: the same successive-sampling law with a quarter of the work when N >= 2 S.
// Drawing without replacement with probabilities ~ w_i is: propose a point uniformly, accept it with probability w_i
// (w_i <= 1), skip points that were already taken, repeat until S points are taken.  Proposal j of query q is a fixed
// function of (seed, q, j) (Philox block j / 2 -> two (index, uniform) pairs), so the accepted set -- the first S distinct
// accepted proposals in proposal order -- does not depend on how the proposals are spread over threads and rounds.
// Only ~S / mean(w) * 1.1 ~ 3 400 of the points are touched per query instead of all 10 000 (plus one cheap pass for the
// maximum distance), and there is no selection or sort.  s_first[i] = position of the first accepted proposal of point i.
constexpr int kRejPer = 8;        // proposals per thread and round (first round); even
constexpr int kRejRounds = 256;

__global__ void __launch_bounds__(kThreads)
subsample_reject_kernel(const float* __restrict__ pts, int N, const float* __restrict__ queries, int64_t qbase,
                        const int32_t* __restrict__ qidx, int S, uint64_t seed, int32_t* __restrict__ out_ids,
                        float* __restrict__ out_pts, int* __restrict__ err_flag) {
    extern __shared__ int s_first[];             // [N]
    __shared__ float redf[kThreads / 32];
    __shared__ int redi[kThreads / 32];
    __shared__ float s_dmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const float qx = queries[q * 3 + 0], qy = queries[q * 3 + 1], qz = queries[q * 3 + 2];
    const uint64_t qi = (uint64_t)(qbase + (qidx ? (int64_t)qidx[q] : q));
    // max_i ||q - p_i|| in float32 like NumPy: sqrt is monotone, so it is the sqrt of the largest float32 squared sum
    float m2 = 0.f;
    for (int i = tid; i < N; i += kThreads) {
        const float dx = __fsub_rn(qx, pts[i * 3 + 0]), dy = __fsub_rn(qy, pts[i * 3 + 1]), dz = __fsub_rn(qz, pts[i * 3 + 2]);
        m2 = fmaxf(m2, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        s_first[i] = 0x7fffffff;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    if (lane == 0) redf[warp] = m2;
    __syncthreads();
    if (tid == 0) {
        float m = redf[0];
        for (int w = 1; w < kThreads / 32; ++w) m = fmaxf(m, redf[w]);
        s_dmax = __fsqrt_rn(m);
    }
    __syncthreads();
    const float dmax = s_dmax;
    int base = 0;                  // distinct accepted points so far
    int pos0 = 0;                  // proposals consumed so far
    int per = kRejPer;             // proposals per thread in this round (even, <= kRejPer)
    for (int round = 0; round < kRejRounds; ++round) {
        const int mypos = pos0 + tid * per;
        int idx[kRejPer];
        unsigned acc = 0;
#pragma unroll
        for (int e2 = 0; e2 < kRejPer / 2; ++e2) {
            if (2 * e2 < per) {
                uint32_t r[4];
                philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)qi, (uint32_t)(qi >> 32), (uint32_t)((mypos >> 1) + e2), 0x9e3779b1u, r);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = (int)__umulhi(r[2 * h], (uint32_t)N);
                    // dist_prob (utils.py:200-208) in float32 like NumPy
                    const float d = norm_f32(__fsub_rn(qx, pts[i * 3 + 0]), __fsub_rn(qy, pts[i * 3 + 1]), __fsub_rn(qz, pts[i * 3 + 2]));
                    const float w = fminf(fmaxf(__fsub_rn(1.0f, __fmul_rn(1.5f, __fdiv_rn(d, dmax))), 0.05f), 1.0f);
                    const float u = (float)(r[2 * h + 1] >> 8) * (1.0f / 16777216.0f);     // [0, 1)
                    idx[2 * e2 + h] = i;
                    if (u < w) { acc |= 1u << (2 * e2 + h); atomicMin(&s_first[i], mypos + 2 * e2 + h); }
                }
            }
        }
        __syncthreads();
        unsigned fresh = 0;
#pragma unroll
        for (int e = 0; e < kRejPer; ++e)
            if (((acc >> e) & 1u) && s_first[idx[e]] == mypos + e) fresh |= 1u << e;
        // exclusive prefix of the fresh counts in thread (= proposal) order
        const int cnt = __popc(fresh);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) redi[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) { const int c = redi[w]; if (w < warp) wbase += c; total += c; }
        int slot = base + wbase + incl - cnt;
#pragma unroll
        for (int e = 0; e < kRejPer; ++e) {
            if ((fresh >> e) & 1u) {
                if (slot < S) {
                    const int i = idx[e];
                    out_ids[q * S + slot] = i;
                    if (out_pts) {
                        float* o = out_pts + (q * S + slot) * 3;
                        o[0] = pts[i * 3 + 0]; o[1] = pts[i * 3 + 1]; o[2] = pts[i * 3 + 2];
                    }
                }
                ++slot;
            }
        }
        base += total;
        if (base >= S) return;
        pos0 += kThreads * per;
        // size of the next round from this round's yield (deterministic: block-uniform integers only)
        const long long need = ((long long)(S - base) * (kThreads * per) * 23 / 20) / (total > 0 ? total : 1) + 1;
        const long long p2 = (need + 2 * kThreads - 1) / (2 * kThreads);
        per = (int)(p2 < 1 ? 1 : (p2 > kRejPer / 2 ? kRejPer / 2 : p2)) * 2;
        __syncthreads();       // redi is reused
    }
    if (tid == 0) atomicExch(err_flag, 2);
}

__global__ void gather_points_kernel(const float* __restrict__ pts, const int32_t* __restrict__ ids, int64_t count, float* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int id = ids[i];
    out[i * 3 + 0] = pts[id * 3 + 0];
    out[i * 3 + 1] = pts[id * 3 + 1];
    out[i * 3 + 2] = pts[id * 3 + 2];
}

}  // namespace

static int* err_flag_dev() {
    static thread_local int* flag = nullptr;
    if (!flag) {
        P2S_CUDA(cudaMalloc(&flag, sizeof(int)));
        P2S_CUDA(cudaMemset(flag, 0, sizeof(int)));
    }
    return flag;
}

void gather_points(const float* pts, const int32_t* ids, int64_t count, float* out, cudaStream_t st);

int assemble_error_check(cudaStream_t st) {  // sync; returns and clears the device error flag
    int h = 0;
    int* f = err_flag_dev();
    P2S_CUDA(cudaMemcpyAsync(&h, f, sizeof(int), cudaMemcpyDeviceToHost, st));
    P2S_CUDA(cudaStreamSynchronize(st));
    if (h) P2S_CUDA(cudaMemsetAsync(f, 0, sizeof(int), st));
    return h;
}

void knn_patch(const float* pts, int64_t N, const float* queries, int64_t Q, int k, int32_t* ids,
               float* patch, float* radius, cudaStream_t st) {
    P2S_CHECK(N >= k, "kNN needs N >= k (the reference returns out-of-range ids otherwise)");
    P2S_CHECK(k >= 1 && k <= kBallMaxK, "k must be in [1, 1536]");
    P2S_CHECK(N < (1 << 30), "cloud too large");
    if (Q <= 0) return;
    // Run length: with 8 queries per CTA a batch of 8 192 queries is 1 024 CTAs = 1.4 waves of the 740 co-resident CTAs, i.e.
    // two waves with the second 38 % full.  Lengthen the runs so that the whole batch is ONE wave (longer runs also amortise
    // the histogram selection of a run's first query better).
    static thread_local int slots_small = 0, slots_big = 0;
    if (!slots_small) {
        int dev = 0, sms = 0, a = 0, b = 0;
        P2S_CUDA(cudaGetDevice(&dev));
        P2S_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        P2S_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, knn_patch_kernel<kCap>, kThreads, 0));
        P2S_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, knn_patch_kernel<kCapBig>, kThreads, 0));
        slots_small = sms * (a > 0 ? a : 1); slots_big = sms * (b > 0 ? b : 1);
    }
    const int slots = k <= 512 ? slots_small : slots_big;
    int run = (int)cdiv(Q, slots);
    if (run < kRun) run = kRun;
    if (run > 64) run = 64;
    if (k <= 512) P2S_LAUNCH(knn_patch_kernel<kCap>, (unsigned)cdiv(Q, run), kThreads, 0, st, pts, (int)N, queries, Q, k, ids, patch, radius, err_flag_dev(), run);
    else P2S_LAUNCH(knn_patch_kernel<kCapBig>, (unsigned)cdiv(Q, run), kThreads, 0, st, pts, (int)N, queries, Q, k, ids, patch, radius, err_flag_dev(), run);
}

// the Philox stream of query q is keyed by qbase + (qidx ? qidx[q] : q), like the sub-sampler's
void ball_patch(const float* pts, int64_t N, const float* queries, int64_t Q, int64_t qbase, int k, double patch_radius,
                uint64_t seed, int32_t* ids, float* patch, float* radius, int32_t* counts, cudaStream_t st, const int32_t* qidx) {
    P2S_CHECK(patch_radius > 0.0, "ball query needs patch_radius > 0");
    P2S_CHECK(k >= 1 && k <= kBallMaxK, "points_per_patch must be in [1, 1536]");
    P2S_CHECK(N >= 1 && N < (1 << 30), "bad cloud size");
    if (Q <= 0) return;
    P2S_LAUNCH(ball_patch_kernel, (unsigned)Q, kThreads, 0, st, pts, (int)N, queries, qbase, qidx, k, patch_radius * patch_radius,
               (float)patch_radius, seed, ids, patch, radius, counts, err_flag_dev());
}

// the Philox stream of query q is keyed by qbase + (qidx ? qidx[q] : q).  pts_out (optional): [Q, S, 3] the selected
// points themselves (what gather_points would produce from `out`).
void subsample(const float* pts, int64_t N, const float* queries, int64_t Q, int64_t qbase, int S,
               int mode, uint64_t seed, int32_t* out, cudaStream_t st, const int32_t* qidx, float* pts_out) {
    P2S_CHECK(N >= S, "sub-sample needs N >= sub_sample_size (reference zero-pads after an in-place shuffle; unsupported)");
    P2S_CHECK(N < (1 << 30), "cloud too large");
    if (Q <= 0) return;
    bool gathered = false;
    if (mode == P2S_SUBSAMPLE_UNIFORM) {
        int64_t threads = Q * ((S + 3) / 4);
        P2S_LAUNCH(subsample_uniform_kernel, (unsigned)cdiv(threads, 256), 256, 0, st, (int)N, Q, qbase, qidx, S, seed, out);
    } else if (mode == P2S_SUBSAMPLE_WEIGHTED) {
        const size_t cache_bytes = (size_t)N * sizeof(float);
        static int no_reject = -1;
        if (no_reject < 0) { const char* e = getenv("P2S_SUBSAMPLE_CLOCKS"); no_reject = (e && e[0] == '1') ? 1 : 0; }
        if (cache_bytes <= 160 * 1024 && N >= 2 * (int64_t)S && !no_reject) {
            // rejection sampling: cheap when at most half of the cloud is drawn (the acceptance rate is >= 0.05 by construction)
            static thread_local bool attr_set = false;
            if (!attr_set) {
                P2S_CUDA(cudaFuncSetAttribute(subsample_reject_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr_set = true;
            }
            P2S_LAUNCH(subsample_reject_kernel, (unsigned)Q, kThreads, cache_bytes, st, pts, (int)N, queries, qbase, qidx, S, seed, out, pts_out, err_flag_dev());
            gathered = true;
        } else if (cache_bytes <= 160 * 1024) {
            static thread_local bool attr_set = false;
            if (!attr_set) {
                P2S_CUDA(cudaFuncSetAttribute(subsample_weighted_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr_set = true;
            }
            P2S_LAUNCH(subsample_weighted_kernel<true>, (unsigned)Q, kThreads, cache_bytes, st, pts, (int)N, queries, qbase, qidx, S, seed, out, err_flag_dev());
        } else {
            P2S_LAUNCH(subsample_weighted_kernel<false>, (unsigned)Q, kThreads, 0, st, pts, (int)N, queries, qbase, qidx, S, seed, out, err_flag_dev());
        }
    } else {
        throw Error("unknown sub-sample mode");
    }
    if (pts_out && !gathered) gather_points(pts, out, Q * S, pts_out, st);
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, int32_t* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void scatter_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int64_t n, float* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}
void gather_i32(const int32_t* src, const int32_t* idx, int64_t n, int32_t* dst, cudaStream_t st) {
    if (n > 0) P2S_LAUNCH(gather_i32_kernel, (unsigned)cdiv(n, 256), 256, 0, st, src, idx, n, dst);
}
void scatter_f32(const float* src, const int32_t* idx, int64_t n, float* dst, cudaStream_t st) {
    if (n > 0) P2S_LAUNCH(scatter_f32_kernel, (unsigned)cdiv(n, 256), 256, 0, st, src, idx, n, dst);
}

void gather_points(const float* pts, const int32_t* ids, int64_t count, float* out, cudaStream_t st) {
    if (count <= 0) return;
    P2S_LAUNCH(gather_points_kernel, (unsigned)cdiv(count, 256), 256, 0, st, pts, ids, count, out);
}

}  // namespace p2s
