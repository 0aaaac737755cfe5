// K6 / K7: SDF epilogue, scatter into the dense volume, iterative sign propagation.
//   post_process + combine       source/sdf_nn.py:11-21, source/points_to_surf_eval.py:184-196,263-271,205-207
//   add_samples_to_volume        source/sdf.py:82-111 (reconstruction case: one sample per voxel -> scatter)
//   propagate_sign               source/sdf.py:114-178
//   clamp                        source/sdf.py:200-202
// Signs are int8; the box sums are exact integers (the reference's float sums of {-1,0,1} are too), so the
// result is bit-identical to the reference.  L2/HBM-bound byte work: per iteration res^3 * ~10 B.
#include "common.cuh"

namespace p2s {

namespace {

__global__ void sdf_from_logits_kernel(const float* __restrict__ logits, const float* __restrict__ radius,
                                       int64_t B, float* __restrict__ sdf) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float t = tanhf(logits[i * 2 + 0]);
    float mag = __fmul_rn(__fmul_rn(t, t), radius[i]);
    float v = logits[i * 2 + 1] >= 0.0f ? mag : -mag;
    if (isnan(v)) v = 1.0f;  // points_to_surf_eval.py:205-207
    sdf[i] = v;
}

__global__ void scatter_kernel(const int32_t* __restrict__ lin, const float* __restrict__ sdf, int64_t Q,
                               float* __restrict__ vol) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Q) vol[lin[i]] = sdf[i];
}

struct Ctrl {
    unsigned long long cnt[2];  // zero count of S, double-buffered by iteration parity
    unsigned long long cntN;    // zero count of the thresholded neighbourhood vote
    int done;
    int iters;                  // applied iterations
    int all_zero;               // every sample is exactly 0 (sdf.py:187-189)
    unsigned nonzero_seen;
};

__global__ void any_nonzero_kernel(const float* __restrict__ sdf, int64_t Q, Ctrl* c) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = (i < Q) && (sdf[i] != 0.0f);
    if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(&c->nonzero_seen, 1u);
}

// block-wide sum of v, ONE atomic per block (a same-address atomic per warp serialises: res^3 / 32 of them cost
// ~0.25 ms per kernel at 256^3).  Every thread of the block must call it.
__device__ __forceinline__ void block_count_add(unsigned v, unsigned long long* dst) {
    __shared__ unsigned warp_sum[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned t = threadIdx.x < (blockDim.x >> 5) ? warp_sum[threadIdx.x] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0 && t) atomicAdd(dst, (unsigned long long)t);
    }
}

// S = sign(vol), U0 = (S == 0), then the six border faces of vol (not of S) are set to -1  (sdf.py:144-154)
__global__ void init_sign_kernel(float* __restrict__ vol, int res, int8_t* __restrict__ S, uint8_t* __restrict__ U0, Ctrl* c) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t V = (int64_t)res * res * res;
    unsigned z = 0;
    if (v < V) {
        float x = vol[v];
        int8_t s = x > 0.f ? 1 : (x < 0.f ? -1 : 0);
        S[v] = s;
        U0[v] = (s == 0);
        z = (s == 0);
        int iz = (int)(v % res), iy = (int)((v / res) % res), ix = (int)(v / ((int64_t)res * res));
        if (ix == 0 || iy == 0 || iz == 0 || ix == res - 1 || iy == res - 1 || iz == res - 1) vol[v] = -1.0f;
    }
    block_count_add(z, &c->cnt[0]);
}

// 1-D box sum along one axis with replicated edges: out[o] = sum_{t=lo..hi} in[clamp(o+t)]
template <int AXIS>
__global__ void box_axis_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out, int res, int lo, int hi,
                                Ctrl* c, int iter) {
    if (c->done) return;
    if (AXIS == 2 && c->cnt[iter & 1] == 0) {  // `if unknown_before.sum() == 0: break`  (sdf.py:157-159)
        if (blockIdx.x == 0 && threadIdx.x == 0) c->done = 1;
        return;
    }
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t V = (int64_t)res * res * res;
    if (AXIS == 2 && v == 0) { c->cntN = 0; c->cnt[(iter + 1) & 1] = 0; }
    if (v >= V) return;
    int iz = (int)(v % res), iy = (int)((v / res) % res), ix = (int)(v / ((int64_t)res * res));
    int pos = AXIS == 2 ? iz : (AXIS == 1 ? iy : ix);
    int64_t stride = AXIS == 2 ? 1 : (AXIS == 1 ? res : (int64_t)res * res);
    const int8_t* base = in + v - (int64_t)pos * stride;
    int acc = 0;
    for (int t = lo; t <= hi; ++t) {
        int p = min(max(pos + t, 0), res - 1);
        acc += base[(int64_t)p * stride];
    }
    out[v] = (int8_t)acc;
}

// last axis (x) fused with the certainty threshold, sign and the zero count of the vote (sdf.py:162-172)
__global__ void box_x_vote_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ vote, int res, int lo, int hi,
                                  float thr, Ctrl* c) {
    if (c->done) return;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t V = (int64_t)res * res * res;
    unsigned z = 0;
    if (v < V) {
        int ix = (int)(v / ((int64_t)res * res));
        int64_t stride = (int64_t)res * res;
        const int8_t* base = in + v - (int64_t)ix * stride;
        int acc = 0;
        for (int t = lo; t <= hi; ++t) {
            int p = min(max(ix + t, 0), res - 1);
            acc += base[(int64_t)p * stride];
        }
        int8_t s = 0;
        if (!(fabsf((float)acc) < thr)) s = acc > 0 ? 1 : (acc < 0 ? -1 : 0);
        vote[v] = s;
        z = (s == 0);
    }
    block_count_add(z, &c->cntN);
}

// `if unknown_after.sum() >= unknown_before.sum(): break` else S[U0] = vote[U0]   (sdf.py:175-177)
__global__ void apply_vote_kernel(int8_t* __restrict__ S, const uint8_t* __restrict__ U0, const int8_t* __restrict__ vote,
                                  int64_t V, Ctrl* c, int iter) {
    if (c->done) return;
    // cntN / cnt[iter&1] are final here (previous kernels completed); every thread takes the same branch
    if (c->cntN >= c->cnt[iter & 1]) {
        if (blockIdx.x == 0 && threadIdx.x == 0) c->done = 1;
        return;
    }
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned z = 0;
    if (v < V) {
        int8_t s = S[v];
        if (U0[v]) { s = vote[v]; S[v] = s; }
        z = (s == 0);
    }
    block_count_add(z, &c->cnt[(iter + 1) & 1]);
    if (v == 0) atomicAdd(&c->iters, 1);
}

// ---- word-wide variants (res % 4 == 0): one thread = 4 consecutive z voxels packed in a 32-bit word, per-byte SIMD adds
// (sums stay within int8: |S| <= 1, three axes of at most 11 taps each only when sigma <= 5: 5^3 = 125; larger sigma use
// the scalar kernels) and grid-stride loops (a few thousand counter atomics per kernel instead of one per 256 voxels).
template <int AXIS>   // 2: z (inside the word and its neighbours), 1: y
__global__ void __launch_bounds__(256)
box_axis4_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int res, int lo, int hi, Ctrl* c, int iter) {
    if (c->done) return;
    if (AXIS == 2 && c->cnt[iter & 1] == 0) {  // `if unknown_before.sum() == 0: break`  (sdf.py:157-159)
        if (blockIdx.x == 0 && threadIdx.x == 0) c->done = 1;
        return;
    }
    if (AXIS == 2 && blockIdx.x == 0 && threadIdx.x == 0) { c->cntN = 0; c->cnt[(iter + 1) & 1] = 0; }
    const unsigned rw = (unsigned)res >> 2;
    const unsigned W = (unsigned)res * (unsigned)res * rw;
    for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
        const unsigned wz = w % rw, row = w / rw;
        if (AXIS == 2) {
            const uint32_t* r = in + (size_t)row * rw;
            // bytes b[8 + j] = position 4*wz + j, j in [-8, 11], replicated at the row ends
            int b[20];
            const int dmin = lo < -4 ? -2 : (lo < 0 ? -1 : 0), dmax = hi > 4 ? 2 : (hi > 0 ? 1 : 0);
            const uint32_t first = r[0] & 0xffu, last = r[rw - 1] >> 24;
#pragma unroll
            for (int d = -2; d <= 2; ++d) {
                uint32_t x = 0;
                if (d >= dmin && d <= dmax) {
                    const int wi = (int)wz + d;
                    x = wi < 0 ? first * 0x01010101u : (wi >= (int)rw ? last * 0x01010101u : r[wi]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) b[4 * (d + 2) + i] = (int)(int8_t)(x >> (8 * i));
            }
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int acc = 0;
#pragma unroll
                for (int t = -5; t <= 5; ++t)           // static indices: b[] stays in registers
                    if (t >= lo && t <= hi) acc += b[8 + k + t];
                o |= ((uint32_t)acc & 0xffu) << (8 * k);
            }
            out[w] = o;
        } else {
            const unsigned iy = row % (unsigned)res, ix = row / (unsigned)res;
            uint32_t acc = 0;
            for (int t = lo; t <= hi; ++t) {
                const int p = min(max((int)iy + t, 0), res - 1);
                acc = __vadd4(acc, in[((size_t)ix * res + p) * rw + wz]);
            }
            out[w] = acc;
        }
    }
}

__global__ void __launch_bounds__(256)
box_x_vote4_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ vote, int res, int lo, int hi, float thr, Ctrl* c) {
    if (c->done) return;
    const unsigned rw = (unsigned)res >> 2;
    const unsigned plane = (unsigned)res * rw;         // words per x plane
    const unsigned W = (unsigned)res * plane;
    unsigned z = 0;
    for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
        const unsigned ix = w / plane, rem = w - ix * plane;
        uint32_t acc = 0;
        for (int t = lo; t <= hi; ++t) {
            const int p = min(max((int)ix + t, 0), res - 1);
            acc = __vadd4(acc, in[(size_t)p * plane + rem]);
        }
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int a = (int)(int8_t)(acc >> (8 * k));
            int s = 0;
            if (!(fabsf((float)a) < thr)) s = a > 0 ? 1 : (a < 0 ? -1 : 0);
            o |= ((uint32_t)s & 0xffu) << (8 * k);
            z += (s == 0);
        }
        vote[w] = o;
    }
    block_count_add(z, &c->cntN);
}

__global__ void __launch_bounds__(256)
apply_vote4_kernel(uint32_t* __restrict__ S, const uint32_t* __restrict__ U0, const uint32_t* __restrict__ vote, unsigned W,
                   Ctrl* c, int iter) {
    if (c->done) return;
    if (c->cntN >= c->cnt[iter & 1]) {       // `if unknown_after.sum() >= unknown_before.sum(): break`
        if (blockIdx.x == 0 && threadIdx.x == 0) c->done = 1;
        return;
    }
    unsigned z = 0;
    for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < W; w += gridDim.x * blockDim.x) {
        const uint32_t m = __vcmpne4(U0[w], 0u);          // 0xff where the voxel was unknown at the start
        const uint32_t s = S[w];
        const uint32_t n = (vote[w] & m) | (s & ~m);
        if (n != s) S[w] = n;
        z += (unsigned)__popc(__vcmpeq4(n, 0u)) >> 3;
    }
    block_count_add(z, &c->cnt[(iter + 1) & 1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&c->iters, 1);
}

// vol[vol == 0] = S[vol == 0]; clamp to [-1, 1]   (sdf.py:179,200-202)
__global__ void finalize_kernel(float* __restrict__ vol, const int8_t* __restrict__ S, int64_t V) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float x = vol[v];
    if (x == 0.0f) x = (float)S[v];
    x = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x);
    vol[v] = x;
}

thread_local DevBuf t_vol_ws;

}  // namespace

void sdf_from_logits(const float* logits, const float* radius, int64_t B, float* sdf, cudaStream_t st) {
    if (B <= 0) return;
    P2S_LAUNCH(sdf_from_logits_kernel, (unsigned)cdiv(B, 256), 256, 0, st, logits, radius, B, sdf);
}

void sdf_to_volume(const int32_t* lin_idx, const float* sdf, int64_t Q, int res, int sigma, float thr,
                   float* vol, int* iterations_host, cudaStream_t st) {
    P2S_CHECK(res >= 2 && res <= 1024, "grid resolution out of range");
    P2S_CHECK(sigma >= 1 && sigma <= 11, "sigma out of range [1, 11]");
    const int64_t V = (int64_t)res * res * res;
    const unsigned blocks = (unsigned)cdiv(V, 256);
    size_t off_S = 256, off_U0 = off_S + V, off_t1 = off_U0 + V, off_t2 = off_t1 + V, off_vote = off_t2 + V;
    uint8_t* base = (uint8_t*)t_vol_ws.get(off_vote + V);
    Ctrl* ctrl = (Ctrl*)base;
    int8_t* S = (int8_t*)(base + off_S);
    uint8_t* U0 = base + off_U0;
    int8_t* t1 = (int8_t*)(base + off_t1);
    int8_t* t2 = (int8_t*)(base + off_t2);
    int8_t* vote = (int8_t*)(base + off_vote);

    P2S_CUDA(cudaMemsetAsync(ctrl, 0, sizeof(Ctrl), st));
    P2S_CUDA(cudaMemsetAsync(vol, 0, (size_t)V * sizeof(float), st));
    if (Q > 0) {
        P2S_LAUNCH(any_nonzero_kernel, (unsigned)cdiv(Q, 256), 256, 0, st, sdf, Q, ctrl);
        P2S_LAUNCH(scatter_kernel, (unsigned)cdiv(Q, 256), 256, 0, st, lin_idx, sdf, Q, vol);
    }
    P2S_LAUNCH(init_sign_kernel, blocks, 256, 0, st, vol, res, S, U0, ctrl);
    // convolve(ones(sigma^3), mode='nearest'): output o sums inputs o-ceil(s/2)+1 .. o+floor(s/2)
    const int lo = -((sigma + 1) / 2) + 1, hi = sigma / 2;
    // word-wide kernels: 4 z voxels per thread, needs word-aligned rows and box sums that fit int8 (sigma^3 <= 127)
    const bool words = (res % 4 == 0) && sigma <= 5;
    int dev_id = 0, sms = 148;
    P2S_CUDA(cudaGetDevice(&dev_id));
    P2S_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id));
    const unsigned wblocks = (unsigned)std::min<int64_t>(cdiv(V / 4, 256), (int64_t)sms * 8);
    Ctrl h{};
    int iter = 0;
    const int kBatch = 8;
    for (;;) {
        for (int b = 0; b < kBatch; ++b, ++iter) {
            if (words) {
                P2S_LAUNCH(box_axis4_kernel<2>, wblocks, 256, 0, st, (const uint32_t*)S, (uint32_t*)t1, res, lo, hi, ctrl, iter);
                P2S_LAUNCH(box_axis4_kernel<1>, wblocks, 256, 0, st, (const uint32_t*)t1, (uint32_t*)t2, res, lo, hi, ctrl, iter);
                P2S_LAUNCH(box_x_vote4_kernel, wblocks, 256, 0, st, (const uint32_t*)t2, (uint32_t*)vote, res, lo, hi, thr, ctrl);
                P2S_LAUNCH(apply_vote4_kernel, wblocks, 256, 0, st, (uint32_t*)S, (const uint32_t*)U0, (const uint32_t*)vote,
                           (unsigned)(V / 4), ctrl, iter);
                continue;
            }
            P2S_LAUNCH(box_axis_kernel<2>, blocks, 256, 0, st, S, t1, res, lo, hi, ctrl, iter);
            P2S_LAUNCH(box_axis_kernel<1>, blocks, 256, 0, st, t1, t2, res, lo, hi, ctrl, iter);
            P2S_LAUNCH(box_x_vote_kernel, blocks, 256, 0, st, t2, vote, res, lo, hi, thr, ctrl);
            P2S_LAUNCH(apply_vote_kernel, blocks, 256, 0, st, S, U0, vote, V, ctrl, iter);
        }
        P2S_CUDA(cudaMemcpyAsync(&h, ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
        P2S_CUDA(cudaStreamSynchronize(st));
        if (h.done) break;
        P2S_CHECK(iter < 64 * res, "sign propagation did not converge");
    }
    P2S_LAUNCH(finalize_kernel, blocks, 256, 0, st, vol, S, V);
    if (iterations_host) *iterations_host = h.iters;
    if (Q > 0 && !h.nonzero_seen) {
        // the reference prints a warning and returns without writing anything (sdf.py:187-189)
        if (iterations_host) *iterations_host = -1;
    }
}

}  // namespace p2s
