// Standalone probe for the tcgen05 building blocks used by net_tc.cu.  Run on a B200:
//   tools/bin/umma_probe
// 1. correctness of the no-swizzle K-major smem descriptors (which LBO/SBO convention is right),
// 2. A-operand-from-TMEM layout, 3. bulk copy + mbarrier, 4. MMA issue rate (SS vs TS, N=128/256),
// 5. TMEM->register epilogue cost with a 3-input max reduction.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../points2surf_b200/csrc/tc_ptx.cuh"

using namespace p2s::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// bounded mbarrier wait: a wrong barrier protocol must not hang the box
__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("TIMEOUT waiting on mbarrier (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
#define mbar_wait wait_or_trap

__device__ __forceinline__ uint32_t kmajor_off(int r, int k, uint32_t lbo, uint32_t sbo) {
    return (uint32_t)(r >> 3) * sbo + (uint32_t)(k >> 3) * lbo + (uint32_t)(r & 7) * 16 + (uint32_t)(k & 7) * 2;
}

struct GemmCfg {
    int M, N, K;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // physical layout strides (bytes)
    int swap_fields;                        // put LBO in the SBO field and vice versa
    int a_from_tmem;
    int use_bulk;                           // stage B through a global buffer + cp.async.bulk
};

// one CTA, 128 threads: D[M=128][N] = A[128][K] * B[N][K]^T
__global__ void __launch_bounds__(128) gemm_probe(GemmCfg c, const __half* __restrict__ A, const __half* __restrict__ B,
                                                  const uint8_t* __restrict__ B_packed, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma, bar_tx;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 65536;
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_tx, 1); fence_mbar_init(); }
    for (int e = tid; e < c.M * c.K; e += 128) {
        int r = e / c.K, k = e % c.K;
        *reinterpret_cast<__half*>(sA + kmajor_off(r, k, c.a_lbo, c.a_sbo)) = A[e];
    }
    if (!c.use_bulk) {
        for (int e = tid; e < c.N * c.K; e += 128) {
            int r = e / c.K, k = e % c.K;
            *reinterpret_cast<__half*>(sB + kmajor_off(r, k, c.b_lbo, c.b_sbo)) = B[e];
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (c.use_bulk) {
        if (tid == 0) {
            uint32_t bytes = (uint32_t)(c.N * c.K * 2);
            mbar_arrive_expect_tx(&bar_tx, bytes);
            bulk_g2s(sB, B_packed, bytes, &bar_tx);
        }
        mbar_wait(&bar_tx, 0);
    }
    const uint32_t a_col = 256;  // TMEM columns for the A operand (TS mode)
    if (c.a_from_tmem) {
        // lane = row; column j holds K elements 2j, 2j+1
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + a_col;
        for (int j0 = 0; j0 < c.K / 2; j0 += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) {
                __half2 h = __halves2half2(A[tid * c.K + 2 * (j0 + j)], A[tid * c.K + 2 * (j0 + j) + 1]);
                v[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            tmem_st_x8(taddr + j0, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(c.M, c.N);
        for (int ks = 0; ks < c.K / 16; ++ks) {
            uint32_t a_addr = smem_u32(sA) + ks * 2 * c.a_lbo;
            uint32_t b_addr = smem_u32(sB) + ks * 2 * c.b_lbo;
            uint64_t da = c.swap_fields ? make_smem_desc(a_addr, c.a_sbo, c.a_lbo) : make_smem_desc(a_addr, c.a_lbo, c.a_sbo);
            uint64_t db = c.swap_fields ? make_smem_desc(b_addr, c.b_sbo, c.b_lbo) : make_smem_desc(b_addr, c.b_lbo, c.b_sbo);
            if (c.a_from_tmem) mma_ts(tmem, tmem + a_col + ks * 8, db, idesc, ks > 0);
            else mma_ss(tmem, da, db, idesc, ks > 0);
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < c.N; n0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16) + n0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[tid * c.N + n0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------
// throughput probe: every CTA issues reps x (K=128 -> 8 MMAs of M=128 x N) back to back
struct RateCfg { int N; int ts; int reps; int epi; };

__global__ void __launch_bounds__(192) rate_probe(RateCfg c, long long* cycles_out, long long* epi_out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma[2], bar_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    // A: 128 x 128 fp16 (32 KB) at 0, B: 256 x 128 fp16 (64 KB) at 32 KB ; LBO=128, SBO=2048 (K=128)
    for (int e = tid; e < (32768 + 65536) / 4; e += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u ^ ((e * 2654435761u) & 0x03ff03ffu);  // halves near 1.0
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar_mma[0], 1); mbar_init(&bar_mma[1], 1); mbar_init(&bar_empty[0], 128); mbar_init(&bar_empty[1], 128); fence_mbar_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (c.ts && warp < 4) {
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = 0x3c003800u + j;
        tmem_st_x32(tmem + ((uint32_t)(warp * 32) << 16) + 448, v);
        tmem_st_x32(tmem + ((uint32_t)(warp * 32) << 16) + 480, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, c.N);
    const uint32_t stage_cols = (c.N <= 128) ? 128 : 256;   // two accumulator stages
    long long t0 = 0, t1 = 0;
    if (warp == 4) {
        if ((tid & 31) == 0) {
            t0 = clock64();
            for (int rep = 0; rep < c.reps; ++rep) {
                const int s = rep & 1;
                if (c.epi) mbar_wait(&bar_empty[s], ((rep >> 1) & 1) ^ 1);   // epilogue drained this stage
                tc_fence_after();
                uint32_t d = tmem + ((c.ts && c.N > 128) ? 0 : s * stage_cols);
                for (int ks = 0; ks < 8; ++ks) {
                    uint64_t da = make_smem_desc(smem_u32(smem) + ks * 256, 128, 2048);
                    uint64_t db = make_smem_desc(smem_u32(smem + 32768) + ks * 256, 128, 2048);
                    if (c.ts) mma_ts(d, tmem + 448 + ks * 8, db, idesc, 1);
                    else mma_ss(d, da, db, idesc, 1);
                }
                if (c.epi) mma_commit(&bar_mma[s]);
            }
            if (!c.epi) mma_commit(&bar_mma[0]);
            else {
                mbar_wait(&bar_mma[(c.reps - 1) & 1], ((c.reps - 1) >> 1) & 1);
            }
            if (c.epi) { t1 = clock64(); cycles_out[blockIdx.x] = t1 - t0; }
        }
    }
    if (!c.epi) {
        mbar_wait(&bar_mma[0], 0);
        if (tid == 128) { t1 = clock64(); cycles_out[blockIdx.x] = t1 - t0; }
    } else if (warp < 4) {
        float acc = -INFINITY;
        long long e0 = clock64();
        for (int rep = 0; rep < c.reps; ++rep) {
            const int s = rep & 1;
            mbar_wait(&bar_mma[s], (rep >> 1) & 1);
            tc_fence_after();
            uint32_t d = tmem + ((uint32_t)(warp * 32) << 16) + s * stage_cols;
            for (int n0 = 0; n0 < c.N; n0 += 32) {
                uint32_t r[32];
                tmem_ld_x32(d + n0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) acc = fmax3(acc, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
            }
            tc_fence_before();
            mbar_arrive(&bar_empty[s]);
        }
        long long e1 = clock64();
        if (tid == 0) epi_out[blockIdx.x] = e1 - e0;
        sink[blockIdx.x * 128 + tid] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static void run_gemm(const char* name, GemmCfg c) {
    std::vector<__half> hA(c.M * c.K), hB(c.N * c.K);
    std::vector<float> fA(c.M * c.K), fB(c.N * c.K);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 5 - 2); hA[i] = __float2half(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 7 - 3); hB[i] = __float2half(fB[i]); }
    // packed image of B in the physical layout (for the bulk-copy variant)
    std::vector<uint8_t> packed(c.N * c.K * 2, 0);
    for (int r = 0; r < c.N; ++r)
        for (int k = 0; k < c.K; ++k) {
            uint32_t off = (uint32_t)(r >> 3) * c.b_sbo + (uint32_t)(k >> 3) * c.b_lbo + (r & 7) * 16 + (k & 7) * 2;
            if (off + 2 <= packed.size()) *reinterpret_cast<__half*>(&packed[off]) = hB[r * c.K + k];
        }
    __half *dA, *dB; uint8_t* dP; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dP, packed.size()));
    CK(cudaMalloc(&dD, c.M * c.N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dP, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, c.M * c.N * 4));
    CK(cudaFuncSetAttribute(gemm_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    gemm_probe<<<1, 128, 131072>>>(c, dA, dB, dP, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s : KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<float> hD(c.M * c.N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < c.M; ++m)
        for (int n = 0; n < c.N; ++n) {
            float ref = 0;
            for (int k = 0; k < c.K; ++k) ref += fA[m * c.K + k] * fB[n * c.K + k];
            double err = fabs((double)hD[m * c.N + n] - ref);
            if (!(err <= 1e-3)) ++bad;
            if (err > maxerr || err != err) maxerr = err;
        }
    printf("%-44s : max_err %.4g  bad %d / %d  %s\n", name, maxerr, bad, c.M * c.N, bad == 0 ? "OK" : "WRONG");
    cudaFree(dA); cudaFree(dB); cudaFree(dP); cudaFree(dD);
}

static void run_rate(const char* name, RateCfg c, int grid) {
    long long *dC, *dE; float* dS;
    CK(cudaMalloc(&dC, grid * 8)); CK(cudaMalloc(&dE, grid * 8)); CK(cudaMalloc(&dS, grid * 128 * 4));
    CK(cudaMemset(dC, 0, grid * 8)); CK(cudaMemset(dE, 0, grid * 8));
    CK(cudaFuncSetAttribute(rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 65536));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_probe<<<grid, 192, 32768 + 65536>>>(c, dC, dE, dS);  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    rate_probe<<<grid, 192, 32768 + 65536>>>(c, dC, dE, dS);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s : KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> hC(grid), hE(grid);
    CK(cudaMemcpy(hC.data(), dC, grid * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hE.data(), dE, grid * 8, cudaMemcpyDeviceToHost));
    long long mn = hC[0], mx = hC[0];
    for (int i = 0; i < grid; ++i) { if (hC[i] < mn) mn = hC[i]; if (hC[i] > mx) mx = hC[i]; }
    double per_mma = (double)mx / (c.reps * 8.0);
    double flops = 2.0 * 128 * c.N * 128 * (double)c.reps * grid;
    printf("%-44s : cycles/MMA(K16) min %.1f max %.1f | epi cycles/tile %.0f | kernel %.3f ms -> %.1f TFLOP/s\n", name,
           (double)mn / (c.reps * 8.0), per_mma, c.epi ? (double)hE[0] / c.reps : 0.0, ms, flops / (ms * 1e-3) / 1e12);
    cudaFree(dC); cudaFree(dE); cudaFree(dS);
}


// ---------------------------------------------------------------------------------------------------
// SWIZZLE_128B K-major correctness: D[128][N] = A[128][K] B[N][K]^T with K = 64 * katoms
__device__ __forceinline__ uint32_t sw128_off(int r, int k, int rows) {
    int atom = k >> 6, kk = k & 63, chunk = kk >> 3;
    return (uint32_t)atom * (uint32_t)rows * 128u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
           (uint32_t)((chunk ^ (r & 7)) * 16) + (uint32_t)(kk & 7) * 2u;
}
__global__ void __launch_bounds__(128) gemm_sw128_probe(int N, int K, const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 65536;
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
    for (int e = tid; e < 128 * K; e += 128) *reinterpret_cast<__half*>(sA + sw128_off(e / K, e % K, 128)) = A[e];
    for (int e = tid; e < N * K; e += 128) *reinterpret_cast<__half*>(sB + sw128_off(e / K, e % K, N)) = B[e];
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint32_t a_addr = smem_u32(sA) + (ks >> 2) * 128 * 128 + (ks & 3) * 32;
            uint32_t b_addr = smem_u32(sB) + (ks >> 2) * N * 128 + (ks & 3) * 32;
            mma_ss(tmem, make_smem_desc_sw128(a_addr), make_smem_desc_sw128(b_addr), idesc, ks > 0);
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < N; n0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16) + n0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[tid * N + n0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static void run_gemm_sw128(const char* name, int N, int K) {
    std::vector<__half> hA(128 * K), hB(N * K);
    std::vector<float> fA(128 * K), fB(N * K);
    srand(2);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 5 - 2); hA[i] = __float2half(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 7 - 3); hB[i] = __float2half(fB[i]); }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(gemm_sw128_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    gemm_sw128_probe<<<1, 128, 131072>>>(N, K, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s : KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<float> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < K; ++k) ref += fA[m * K + k] * fB[n * K + k];
            double err = fabs((double)hD[m * N + n] - ref);
            if (!(err <= 1e-3)) ++bad;
            if (err > maxerr || err != err) maxerr = err;
        }
    printf("%-44s : max_err %.4g  bad %d / %d  %s\n", name, maxerr, bad, 128 * N, bad == 0 ? "OK" : "WRONG");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// generalised issue-rate probe: K=128 (8 MMAs) per rep, A 128x128, B Nx128
struct Rate2 { int N; int ts; int layout; int alt; int reps; };   // layout 0: LBO128/SBO2048, 1: LBO rows*16/SBO128, 2: SW128
__global__ void __launch_bounds__(160) rate2_probe(Rate2 c, long long* cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (32768 + 65536) / 4; e += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u ^ ((e * 2654435761u) & 0x03ff03ffu);
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (c.ts && warp < 4) {
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = 0x3c003800u + j;
        tmem_st_x32(tmem + ((uint32_t)(warp * 32) << 16) + 448, v);
        tmem_st_x32(tmem + ((uint32_t)(warp * 32) << 16) + 480, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, c.N);
    if (tid == 128) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
        long long t0 = clock64();
        for (int rep = 0; rep < c.reps; ++rep) {
            for (int ks = 0; ks < 8; ++ks) {
                uint64_t da, db;
                if (c.layout == 0) { da = make_smem_desc(sa + ks * 256, 128, 2048); db = make_smem_desc(sb + ks * 256, 128, 2048); }
                else if (c.layout == 1) { da = make_smem_desc(sa + ks * 4096, 2048, 128); db = make_smem_desc(sb + ks * 2 * c.N * 16, c.N * 16, 128); }
                else { da = make_smem_desc_sw128(sa + (ks >> 2) * 16384 + (ks & 3) * 32); db = make_smem_desc_sw128(sb + (ks >> 2) * c.N * 128 + (ks & 3) * 32); }
                uint32_t d = tmem;
                if (c.alt && c.N <= 192) d += ((rep * 8 + ks) & 1) * 192;
                if (c.ts) mma_ts(d, tmem + 448 + ks * 8, db, idesc, 1);
                else mma_ss(d, da, db, idesc, 1);
            }
        }
        mma_commit(&bar_mma);
        mbar_wait(&bar_mma, 0);
        cycles_out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static void run_rate2(const char* name, Rate2 c, int grid) {
    long long* dC;
    CK(cudaMalloc(&dC, grid * 8));
    CK(cudaFuncSetAttribute(rate2_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 65536));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate2_probe<<<grid, 160, 32768 + 65536>>>(c, dC);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    rate2_probe<<<grid, 160, 32768 + 65536>>>(c, dC);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s : KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> hC(grid);
    CK(cudaMemcpy(hC.data(), dC, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) if (hC[i] > mx) mx = hC[i];
    double ideal = 128.0 * c.N / 256.0;
    printf("%-44s : cycles/MMA %.1f (ideal %.0f, eff %.0f%%) kernel %.3f ms -> %.0f TFLOP/s\n", name, (double)mx / (c.reps * 8.0), ideal,
           100.0 * ideal / ((double)mx / (c.reps * 8.0)), ms, 2.0 * 128 * c.N * 128 * (double)c.reps * grid / (ms * 1e-3) / 1e12);
    cudaFree(dC);
}

// ---------------------------------------------------------------------------------------------------
// probe 3: tight issue loop (descriptors precomputed, compile-time shape) -- the real tensor-pipe rate
template <int N, bool TS, bool EPI>
__global__ void __launch_bounds__(192) rate3_probe(int reps, long long* cycles_out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full[2], bar_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (32768 + 65536) / 4; e += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u ^ ((e * 2654435761u) & 0x03ff03ffu);
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1); mbar_init(&bar_empty[0], 128); mbar_init(&bar_empty[1], 128); fence_mbar_init(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t kStage = (N <= 128) ? 128 : 256;
    constexpr uint32_t kACol = 2 * kStage;    // A operand columns (TS); only valid when N <= 128 or !EPI
    if (TS && warp < 4) {
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = 0x3c003800u + j;
        const uint32_t col = (N <= 128) ? kACol : 448;
        tmem_st_x32(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
        tmem_st_x32(tmem + ((uint32_t)(warp * 32) << 16) + col + 32, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    if (warp == 4) {
        if ((tid & 31) == 0) {
            uint64_t da[8], db[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                da[ks] = make_smem_desc(smem_u32(smem) + ks * 256, 128, 2048);
                db[ks] = make_smem_desc(smem_u32(smem + 32768) + ks * 256, 128, 2048);
            }
            const uint32_t acol = tmem + ((N <= 128) ? kACol : 448);
            long long t0 = clock64();
            for (int rep = 0; rep < reps; ++rep) {
                const int s = rep & 1;
                if (EPI) { mbar_wait(&bar_empty[s], ((rep >> 1) & 1) ^ 1); tc_fence_after(); }
                const uint32_t d = tmem + ((TS && N > 128) ? 0 : s * kStage);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    if (TS) mma_ts(d, acol + ks * 8, db[ks], idesc, 1);
                    else mma_ss(d, da[ks], db[ks], idesc, 1);
                }
                if (EPI) mma_commit(&bar_full[s]);
            }
            if (!EPI) { mma_commit(&bar_full[0]); mbar_wait(&bar_full[0], 0); }
            else mbar_wait(&bar_full[(reps - 1) & 1], ((reps - 1) >> 1) & 1);
            cycles_out[blockIdx.x] = clock64() - t0;
        }
    } else if (EPI && warp < 4) {
        float acc = -INFINITY;
        for (int rep = 0; rep < reps; ++rep) {
            const int s = rep & 1;
            mbar_wait(&bar_full[s], (rep >> 1) & 1);
            tc_fence_after();
            const uint32_t d = tmem + ((uint32_t)(warp * 32) << 16) + s * kStage;
#pragma unroll
            for (int n0 = 0; n0 < N; n0 += 32) {
                uint32_t r[32];
                tmem_ld_x32(d + n0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) acc = fmax3(acc, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
            }
            tc_fence_before();
            mbar_arrive(&bar_empty[s]);
        }
        sink[blockIdx.x * 128 + tid] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, bool TS, bool EPI>
static void run_rate3(const char* name, int grid, int reps) {
    long long* dC; float* dS;
    CK(cudaMalloc(&dC, grid * 8)); CK(cudaMalloc(&dS, grid * 128 * 4));
    auto k = rate3_probe<N, TS, EPI>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 65536));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, 192, 32768 + 65536>>>(reps, dC, dS);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<<<grid, 192, 32768 + 65536>>>(reps, dC, dS);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s : KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> hC(grid);
    CK(cudaMemcpy(hC.data(), dC, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) if (hC[i] > mx) mx = hC[i];
    double ideal = 128.0 * N / 256.0, got = (double)mx / (reps * 8.0);
    printf("%-44s : cycles/MMA %.1f (ideal %.0f, eff %.0f%%) kernel %.3f ms -> %.0f TFLOP/s\n", name, got, ideal, 100.0 * ideal / got, ms,
           2.0 * 128 * N * 128 * (double)reps * grid / (ms * 1e-3) / 1e12);
    cudaFree(dC); cudaFree(dS);
}

// ---------------------------------------------------------------------------------------------------
// probe 4: TMEM -> register read bandwidth (tcgen05.ld 32x32b.x32), optionally with a concurrent MMA stream
template <int NWARPS, bool WITH_MMA, int BATCH>
__global__ void __launch_bounds__(NWARPS * 32 + 32) tmem_bw_probe(int reps, long long* cycles_out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (32768 + 32768) / 4; e += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u ^ ((e * 2654435761u) & 0x03ff03ffu);
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp < NWARPS) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) * 128);
        float acc = 0.f;
        __syncwarp();
        long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            uint32_t r[BATCH][32];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) tmem_ld_x32(base + ((rep * BATCH + b) & 3) * 32, r[b]);
            tmem_ld_wait();
#pragma unroll
            for (int b = 0; b < BATCH; ++b) acc = fmax3(acc, __uint_as_float(r[b][0]), __uint_as_float(r[b][31]));
        }
        long long t1 = clock64();
        if ((tid & 31) == 0) cycles_out[blockIdx.x * 8 + warp] = t1 - t0;
        sink[blockIdx.x * 256 + tid] = acc;
        if (WITH_MMA) { __syncwarp(); if (tid == 0) stop = 1; }
    } else if (WITH_MMA && (tid & 31) == 0) {
        // background MMA stream into columns 256..511 (N = 128, two accumulators), until the readers finish
        const uint32_t idesc = make_idesc_f16(128, 128);
        uint64_t da[8], db[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) { da[ks] = make_smem_desc(smem_u32(smem) + ks * 256, 128, 2048); db[ks] = make_smem_desc(smem_u32(smem + 32768) + ks * 256, 128, 2048); }
        int it = 0;
        while (!stop && it < 200000) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) mma_ss(tmem + 256 + (it & 1) * 128, da[ks], db[ks], idesc, 1);
            ++it;
            if ((it & 15) == 0) { mma_commit(&bar); wait_or_trap(&bar, ((it >> 4) - 1) & 1); }   // bound the queue depth
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int NWARPS, bool WITH_MMA, int BATCH>
static void run_tmem_bw(const char* name, int grid, int reps) {
    long long* dC; float* dS;
    CK(cudaMalloc(&dC, grid * 8 * 8)); CK(cudaMalloc(&dS, grid * 256 * 4));
    CK(cudaMemset(dC, 0, grid * 64));
    auto k = tmem_bw_probe<NWARPS, WITH_MMA, BATCH>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    k<<<grid, NWARPS * 32 + 32, 65536>>>(reps, dC, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s : KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<long long> hC(grid * 8);
    CK(cudaMemcpy(hC.data(), dC, grid * 64, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid * 8; ++i) if (hC[i] > mx) mx = hC[i];
    double bytes = (double)reps * BATCH * 4096.0 * NWARPS;
    printf("%-44s : %.1f cycles per LDTM.x32 per warp, %.1f B/cycle/SM\n", name, (double)mx / (reps * BATCH), bytes / (double)mx);
    cudaFree(dC); cudaFree(dS);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs %d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    if (argc > 1 && atoi(argv[1]) == 4) {
        const int g = p.multiProcessorCount;
        run_tmem_bw<4, false, 1>("tmem read 4 warps batch 1", g, 4096);
        run_tmem_bw<4, false, 2>("tmem read 4 warps batch 2", g, 2048);
        run_tmem_bw<4, false, 4>("tmem read 4 warps batch 4", g, 1024);
        run_tmem_bw<8, false, 2>("tmem read 8 warps batch 2", g, 2048);
        run_tmem_bw<8, false, 4>("tmem read 8 warps batch 4", g, 1024);
        run_tmem_bw<1, false, 4>("tmem read 1 warp  batch 4", g, 1024);
        run_tmem_bw<4, true, 4>("tmem read 4 warps batch 4 + MMA stream", g, 1024);
        run_tmem_bw<8, true, 4>("tmem read 8 warps batch 4 + MMA stream", g, 1024);
        return 0;
    }
    if (argc > 1 && atoi(argv[1]) == 3) {
        const int g = p.multiProcessorCount;
        run_rate3<256, false, false>("r3 SS N=256", g, 512);
        run_rate3<128, false, false>("r3 SS N=128", g, 512);
        run_rate3<64, false, false>("r3 SS N=64", g, 512);
        run_rate3<256, true, false>("r3 TS N=256", g, 512);
        run_rate3<128, true, false>("r3 TS N=128", g, 512);
        run_rate3<64, true, false>("r3 TS N=64", g, 512);
        run_rate3<256, false, true>("r3 SS N=256 + max epilogue", g, 512);
        run_rate3<128, false, true>("r3 SS N=128 + max epilogue", g, 512);
        run_rate3<128, true, true>("r3 TS N=128 + max epilogue", g, 512);
        run_rate3<64, true, true>("r3 TS N=64 + max epilogue", g, 512);
        run_rate3<256, false, true>("r3 SS N=256 + max epilogue, 20k reps (sustained)", g, 20000);
        run_rate3<128, true, true>("r3 TS N=128 + max epilogue, 40k reps (sustained)", g, 40000);
        return 0;
    }
    const int M = 128, N = 64, K = 64;
    // layout A: [r/8][k/8] (SBO = K/8*128, LBO = 128); layout B: [k/8][r/8] (LBO = R/8*128, SBO = 128)
    GemmCfg a{M, N, K, 128, (uint32_t)(K / 8 * 128), 128, (uint32_t)(K / 8 * 128), 0, 0, 0};
    run_gemm("SS layoutA fields(lbo,sbo)", a);
    a.swap_fields = 1; run_gemm("SS layoutA fields swapped", a);
    GemmCfg b{M, N, K, (uint32_t)(M / 8 * 128), 128, (uint32_t)(N / 8 * 128), 128, 0, 0, 0};
    run_gemm("SS layoutB fields(lbo,sbo)", b);
    b.swap_fields = 1; run_gemm("SS layoutB fields swapped", b);
    GemmCfg t{M, N, K, 128, (uint32_t)(K / 8 * 128), 128, (uint32_t)(K / 8 * 128), 0, 1, 0};
    run_gemm("TS (A in TMEM) layoutA", t);
    t.swap_fields = 1; run_gemm("TS (A in TMEM) layoutA swapped", t);
    GemmCfg u{M, N, K, 128, (uint32_t)(K / 8 * 128), 128, (uint32_t)(K / 8 * 128), 0, 0, 1};
    run_gemm("SS layoutA, B via cp.async.bulk", u);
    u.swap_fields = 1; run_gemm("SS layoutA swapped, B via cp.async.bulk", u);
    GemmCfg w{128, 256, 128, 128, 2048, 128, 2048, 0, 0, 0};
    run_gemm("SS layoutA M128 N256 K128", w);
    w.swap_fields = 1; run_gemm("SS layoutA swapped M128 N256 K128", w);
    GemmCfg x{128, 256, 128, 128, 2048, 128, 2048, 0, 1, 0};
    run_gemm("TS layoutA M128 N256 K128", x);

    const int grid = p.multiProcessorCount;
    run_rate("rate SS N=256 (1 CTA)", RateCfg{256, 0, 256, 0}, 1);
    run_rate("rate SS N=256 (all SMs)", RateCfg{256, 0, 256, 0}, grid);
    run_rate("rate SS N=128 (all SMs)", RateCfg{128, 0, 256, 0}, grid);
    run_rate("rate TS N=256 (all SMs)", RateCfg{256, 1, 256, 0}, grid);
    run_rate("rate TS N=128 (all SMs)", RateCfg{128, 1, 256, 0}, grid);
    run_rate("rate SS N=128 + max epilogue (all SMs)", RateCfg{128, 0, 256, 1}, grid);
    run_rate("rate TS N=128 + max epilogue (all SMs)", RateCfg{128, 1, 256, 1}, grid);
    run_rate("rate SS N=256 + max epilogue (all SMs)", RateCfg{256, 0, 256, 1}, grid);
    run_gemm_sw128("SS SW128 N=64 K=64", 64, 64);
    run_gemm_sw128("SS SW128 N=256 K=128", 256, 128);
    run_rate2("r2 SS nosw layoutA N=256", Rate2{256, 0, 0, 0, 256}, grid);
    run_rate2("r2 SS nosw layoutB N=256", Rate2{256, 0, 1, 0, 256}, grid);
    run_rate2("r2 SS SW128 N=256", Rate2{256, 0, 2, 0, 256}, grid);
    run_rate2("r2 SS SW128 N=128", Rate2{128, 0, 2, 0, 256}, grid);
    run_rate2("r2 SS SW128 N=128 alt-acc", Rate2{128, 0, 2, 1, 256}, grid);
    run_rate2("r2 SS SW128 N=192", Rate2{192, 0, 2, 0, 256}, grid);
    run_rate2("r2 SS SW128 N=64", Rate2{64, 0, 2, 0, 256}, grid);
    run_rate2("r2 TS SW128 N=256", Rate2{256, 1, 2, 0, 256}, grid);
    run_rate2("r2 TS SW128 N=128", Rate2{128, 1, 2, 0, 256}, grid);
    run_rate2("r2 TS SW128 N=128 alt-acc", Rate2{128, 1, 2, 1, 256}, grid);
    run_rate2("r2 TS SW128 N=64", Rate2{64, 1, 2, 0, 256}, grid);
    run_rate2("r2 TS nosw N=128 alt-acc", Rate2{128, 1, 0, 1, 256}, grid);
    run_rate2("r2 TS nosw N=192", Rate2{192, 1, 0, 0, 256}, grid);
    run_rate2("r2 SS nosw N=64", Rate2{64, 0, 0, 0, 256}, grid);
    return 0;
}
