// Building-block checks for the CTA-pair (tcgen05 cta_group::2) kernels:
//   cnerf_umma_selftest_pair   d[256,n] = a[256,k] b[n,k]^T with ONE M=256 instruction stream issued by the leader CTA; each
//                              CTA of the pair holds its 128 rows of A/D and n/2 rows of B (fp16 hi/lo split, 3 MMAs per MAC)
//   cnerf_debug_umma_rate_pair cycles per M=256 x N=256 x K=16 instruction, back to back, on all 74 pairs
#include "umma.cuh"

namespace cnerf {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_selftest_pair_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int k, float* __restrict__ d) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = cluster_ctarank();
    const int kgs = k / 8, nh = n / 2;
    const uint32_t a_hi = sbase, a_lo = a_hi + kgs * 2048;
    const uint32_t lbo_b = (uint32_t)nh * 16;
    const uint32_t b_hi = a_lo + kgs * 2048, b_lo = b_hi + kgs * lbo_b;
    const uint32_t bar = b_lo + kgs * lbo_b;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + (bar - sbase) + 8);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc2(bar + 8, 256);
    // A: this CTA's 128 rows; B: this CTA's n/2 rows
    for (int kg = 0; kg < kgs; ++kg) {
        float v[8];
        for (int e = 0; e < 8; ++e) v[e] = a[(size_t)(rank * 128 + threadIdx.x) * k + kg * 8 + e];
        store_split8(a_hi, a_lo, threadIdx.x, kg, v);
    }
    for (int r = threadIdx.x; r < nh; r += 128)
        for (int kg = 0; kg < kgs; ++kg) {
            float v[8];
            for (int e = 0; e < 8; ++e) v[e] = b[(size_t)(rank * nh + r) * k + kg * 8 + e];
            uint32_t h[4], l[4];
            for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
            uint32_t off = kg * lbo_b + r * 16;
            st_shared_v4(b_hi + off, h[0], h[1], h[2], h[3]);
            st_shared_v4(b_lo + off, l[0], l[1], l[2], l[3]);
        }
    fence_proxy_async();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(256, (uint32_t)n);
        for (int ks = 0; ks < k / 16; ++ks) {
            uint64_t ah = smem_desc(a_hi + ks * 2 * kLBO), al = smem_desc(a_lo + ks * 2 * kLBO);
            uint64_t bh = smem_desc_any(b_hi + ks * 2 * lbo_b, lbo_b, kSBO), bl = smem_desc_any(b_lo + ks * 2 * lbo_b, lbo_b, kSBO);
            umma2_f16(tmem, ah, bh, idesc, ks == 0 ? 0u : 1u);
            umma2_f16(tmem, ah, bl, idesc, 1u);
            umma2_f16(tmem, al, bh, idesc, 1u);
        }
        umma2_commit(bar, 3);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < n; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32 && c + j < n; ++j) d[(size_t)(rank * 128 + threadIdx.x) * n + c + j] = v[j];
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

// Layout probe: M=128 cta_group::2 (64 rows of A per CTA, n/2 rows of B per CTA); every CTA dumps its whole TMEM window
// [128 lanes][256 columns] so the host can locate where D(row, col) lands.  The accumulator is zero-filled first.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_pair_layout_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int k, float* __restrict__ dump) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = cluster_ctarank();
    const int kgs = k / 8, nh = n / 2;
    const uint32_t lbo_a = 64 * 16, lbo_b = (uint32_t)nh * 16;
    const uint32_t a_hi = sbase, b_hi = a_hi + kgs * lbo_a, bar = b_hi + kgs * lbo_b;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + (bar - sbase) + 8);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc2(bar + 8, 256);
    auto put = [&](uint32_t base, uint32_t lbo, int r, int kg, const float* src) {
        float v[8];
        for (int e = 0; e < 8; ++e) v[e] = src[kg * 8 + e];
        uint32_t h[4], l[4];
        for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
        st_shared_v4(base + kg * lbo + r * 16, h[0], h[1], h[2], h[3]);
    };
    for (int r = threadIdx.x; r < 64; r += 128)
        for (int kg = 0; kg < kgs; ++kg) put(a_hi, lbo_a, r, kg, a + (size_t)(rank * 64 + r) * k);
    for (int r = threadIdx.x; r < nh; r += 128)
        for (int kg = 0; kg < kgs; ++kg) put(b_hi, lbo_b, r, kg, b + (size_t)(rank * nh + r) * k);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    {   // zero the window
        uint32_t z[16];
        for (int i = 0; i < 16; ++i) z[i] = 0;
        for (int c = 0; c < 256; c += 16) tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + c, z);
        tmem_st_wait();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t idesc = instr_desc(128, (uint32_t)n);
        for (int ks = 0; ks < k / 16; ++ks)
            umma2_f16(tmem, smem_desc_any(a_hi + ks * 2 * lbo_a, lbo_a, kSBO), smem_desc_any(b_hi + ks * 2 * lbo_b, lbo_b, kSBO), idesc, ks == 0 ? 0u : 1u);
        umma2_commit(bar, 3);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < 256; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) dump[((size_t)rank * 128 + threadIdx.x) * 256 + c + j] = v[j];
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

// `iters` back-to-back MMAs (N=256, K=16; kPair: M=256 cta_group::2 on 74 pairs, else M=128 cta_group::1 on 148 CTAs), groups
// of 4 alternating between two accumulators.  traffic bit 0: warps 2-3 keep writing 16-byte shared-memory stores (an
// epilogue-like load on the SMEM port); bit 1: warp 1 streams 16 KB (8 KB per CTA in pair mode) bulk copies global ->
// shared through a 4-slot ring, like the weight loader.
template <bool kPair>
__global__ void __launch_bounds__(128, 1)
umma_bench_pair_kernel(int iters, int traffic, const uint8_t* __restrict__ src, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = kPair ? cluster_ctarank() : 0;
    const uint32_t a_s = sbase, b_s = sbase + 8192, scratch = sbase + 8192 + 32768, ring = scratch + 16384, bar = ring + 65536;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + (bar - sbase) + 8);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + (bar - sbase) + 16);
    const uint32_t rbar = bar + 32;                                       // 4 ring barriers
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (8192 + 32768) / 4; i += 128) {
        uint32_t x = (i + 77u * blockIdx.x) * 2654435761u;            // traffic bit 4: pseudo-random fp16 pairs in (-2, 2) instead of zeros
        reinterpret_cast<uint32_t*>(smem)[i] = (traffic & 16) ? ((x & 0x83ff83ffu) | 0x3c003c00u) : 0u;
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 24, 1); mbar_init(bar + 64, 1); for (int s = 0; s < 4; ++s) mbar_init(rbar + 8 * s, 1); *stop = 0; fence_barrier_init(); }
    if (warp == 0) { if (kPair) tmem_alloc2(bar + 8, 512); else tmem_alloc(bar + 8, 512); }
    fence_proxy_async();
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 0) {
        if (rank == 0) {
            const uint32_t idesc = instr_desc(kPair ? ((traffic & 4) ? 128 : 256) : 128, 256);      // traffic bit 2: pair M=128 (64 rows per CTA)
            const uint32_t lbo_b = (kPair ? 128 : 256) * 16;
            uint64_t bd[4], ad[4];
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) {
                bd[ks] = smem_desc_any(b_s + ks * 2 * lbo_b, lbo_b, kSBO);
                ad[ks] = smem_desc(a_s + ks * 2 * kLBO);
            }
            long long t0 = clock64();
            if (threadIdx.x == 0) mbar_arrive(bar + 64);          // a completed barrier for the per-group wait experiment
            __syncwarp();
            for (int i = 0; i < iters; i += 4) {
                if (traffic & 32) mbar_wait_cluster(bar + 64, 0);            // the fused kernel's per-block waits and fence
                if (traffic & 64) tc_fence_after();
                if (elect_one()) {
                    if (kPair && (traffic & 8)) {           // the fused kernel's pattern: 3 MMAs on one accumulator + a multicast commit
                        umma2_f16(tmem, ad[0], bd[0], idesc, 1u); umma2_f16(tmem, ad[0], bd[1], idesc, 1u);
                        umma2_f16(tmem, ad[1], bd[0], idesc, 1u); umma2_f16(tmem, ad[2], bd[2], idesc, 1u);
                        umma2_commit(bar + 24, 3);
                    } else if (kPair) {
                        umma2_f16(tmem, ad[0], bd[0], idesc, 1u); umma2_f16(tmem + 256, ad[1], bd[1], idesc, 1u);
                        umma2_f16(tmem, ad[2], bd[2], idesc, 1u); umma2_f16(tmem + 256, ad[3], bd[3], idesc, 1u);
                    } else {
                        umma_f16(tmem, ad[0], bd[0], idesc, 1u); umma_f16(tmem + 256, ad[1], bd[1], idesc, 1u);
                        umma_f16(tmem, ad[2], bd[2], idesc, 1u); umma_f16(tmem + 256, ad[3], bd[3], idesc, 1u);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) { if (kPair) umma2_commit(bar, 3); else umma_commit(bar); }
            __syncwarp();
            mbar_wait(bar, 0);
            long long t1 = clock64();
            if (threadIdx.x == 0) out[kPair ? blockIdx.x >> 1 : blockIdx.x] = (float)(t1 - t0) / (float)iters;
        } else {
            mbar_wait(bar, 0);
        }
        *stop = 1;
    } else if (warp == 1) {
        if ((traffic & 2) && (threadIdx.x & 31) == 0) {
            const uint32_t bytes = kPair ? 8192 : 16384;
            uint32_t it = 0;
            while (!*stop) {
                const uint32_t s = it & 3;
                if (it >= 4) mbar_wait(rbar + 8 * s, ((it >> 2) - 1) & 1);
                mbar_arrive_expect_tx(rbar + 8 * s, bytes);
                bulk_g2s(ring + s * 16384, src + (size_t)((it * 7 + blockIdx.x) & 63) * 16384, bytes, rbar + 8 * s);
                ++it;
            }
            for (uint32_t k = it > 4 ? it - 4 : 0; k < it; ++k) mbar_wait(rbar + 8 * (k & 3), (k >> 2) & 1);
        }
    } else if (traffic & 1) {
        uint32_t addr = scratch + (threadIdx.x - 64) * 16, x = threadIdx.x;
        while (!*stop) {
#pragma unroll
            for (int r = 0; r < 8; ++r) st_shared_v4(addr + (r & 7) * 1024, x, x + 1, x + 2, x + 3);
        }
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();
    if (warp == 0) { if (kPair) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}

}  // namespace cnerf

using namespace cnerf;

extern "C" int cnerf_umma_selftest_pair(const float* a, const float* b, int n, int k, float* d, void* stream) {
    CNERF_REQUIRE(a && b && d, "cnerf_umma_selftest_pair: null pointer");
    CNERF_REQUIRE(n >= 32 && n <= 256 && n % 32 == 0 && k >= 16 && k <= 128 && k % 16 == 0, "cnerf_umma_selftest_pair: bad n/k");
    size_t smem = (size_t)(k / 8) * 2048 * 2 + (size_t)(k / 8) * (n / 2) * 16 * 2 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(umma_selftest_pair_kernel)");
    umma_selftest_pair_kernel<<<2, 128, smem, as_stream(stream)>>>(a, b, n, k, d);
    CNERF_LAUNCH_CHECK("umma_selftest_pair_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_debug_umma_rate_pair(int pair, int iters, int traffic, const void* src, float* out, void* stream) {
    CNERF_REQUIRE(out && src && iters > 0, "cnerf_debug_umma_rate_pair: bad arguments");
    size_t smem = 8192 + 32768 + 16384 + 65536 + 128;
    cudaError_t e = cudaFuncSetAttribute(umma_bench_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_bench_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(umma_bench_pair_kernel)");
    if (pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kNumSMs); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = as_stream(stream);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, umma_bench_pair_kernel<true>, iters, traffic, (const uint8_t*)src, out);
        if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(umma_bench_pair_kernel)");
    } else {
        umma_bench_pair_kernel<false><<<kNumSMs, 128, smem, as_stream(stream)>>>(iters, traffic, (const uint8_t*)src, out);
    }
    CNERF_LAUNCH_CHECK("umma_bench_pair_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_debug_pair_layout(const float* a, const float* b, int n, int k, float* dump, void* stream) {
    CNERF_REQUIRE(a && b && dump && n >= 32 && n <= 256 && n % 32 == 0 && k >= 16 && k <= 64 && k % 16 == 0, "cnerf_debug_pair_layout: bad arguments");
    size_t smem = (size_t)(k / 8) * 1024 + (size_t)(k / 8) * (n / 2) * 16 + 64;
    umma_pair_layout_kernel<<<2, 128, smem, as_stream(stream)>>>(a, b, n, k, dump);
    CNERF_LAUNCH_CHECK("umma_pair_layout_kernel");
    return CNERF_OK;
}
