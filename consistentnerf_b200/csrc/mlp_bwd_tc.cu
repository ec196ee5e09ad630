// K3b: backward of the canonical 8x256 NeRF MLP on tcgen05 tensor cores (sm_100a).
//
// Inputs: d_raw [P,4] (from the compositing backward), the activation records written by the training
// forward (mlp_fwd3.cu, kSave), the live weights.  Three stages, all with the forward's fp16 hi/lo split
// (3 MMAs per MAC, fp32 accumulation in TMEM):
//
//   1. mlp_bwd_data_kernel   the data-gradient chain, fused over the layers like the forward: per 128-point
//      tile G9 = (d_rgb W_rgb) * [hv>0]  ->  G8 = G9 W9[:, :256]  ->  G7 = (G8 W8 + d_sigma w_alpha) * [h7>0]
//      ->  G_{l-1} = (G_l W_l) * [h_{l-1}>0] ... down to G0.  G_l is the gradient w.r.t. the pre-activation of
//      layer l; every G_l tile is streamed to HBM as the next stage's operand.  Transposed weight blocks come
//      through the same cp.async.bulk ring as in the forward.
//   2. mlp_bwd_weight_kernel dW_l = G_l^T X_l as a GEMM whose reduction dimension is the points: both operands
//      are read back as MN-major UMMA operands (the record layout is valid for both majors), accumulators stay
//      in TMEM over all tiles of a CTA, one partial per CTA, deterministic reduction.  Spare warps sum the bias
//      gradients from the same shared-memory tiles.
//   3. mlp_heads_grad_kernel the two narrow heads (alpha_linear 256->1, rgb_linear 128->3) in fp32 on CUDA cores.
//
// Gradients are tiny (mean losses over thousands of rays), so G is carried times a power-of-two scale derived
// on the device from max|d_raw| (no host sync); the final reductions multiply by its exact inverse.
#include "mlp_layout.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace cnerf {

// gradient record per 128-point tile: G9 (K=128: [hi 32K | lo 32K]) then G8, G7, ..., G0 (128 KB each)
__host__ __device__ constexpr size_t g_slot(int l) { return l == 9 ? 0 : 65536 + (size_t)(8 - l) * 131072; }
constexpr size_t kGTileBytes = 65536 + 9 * 131072;     // 1 245 184 B

// workspace map (bytes)
constexpr size_t kWsAmax = 0;                                           // uint32: bits of max|d_raw|
constexpr int kDwPasses = 10;
constexpr size_t kDwPassFloats = (size_t)kNumSMs * 128 * 512;           // one pass: [148][128 x 512] fp32
constexpr size_t kDbPassFloats = (size_t)kNumSMs * 2 * 256;             // one pass: [148][2][256] fp32
constexpr size_t kWsDwPart = 256;                                       // [10 passes][148][128 x 512] fp32
constexpr size_t kWsDbPart = kWsDwPart + kDwPasses * kDwPassFloats * 4; // [10 passes][148][2][256] fp32
constexpr int kHeadCtas = 2 * kNumSMs;                                  // the narrow-head kernel is a pure HBM stream: two CTAs per SM
constexpr size_t kWsHeadPart = kWsDbPart + kDwPasses * kDbPassFloats * 4;   // [kHeadCtas][kHeadFloats] fp32
constexpr int kHeadFloats = 384 + 256 + 4;                              // dW_rgb, dW_alpha, db_rgb(3)+db_alpha
constexpr size_t kWsBytes = kWsHeadPart + (size_t)kHeadCtas * kHeadFloats * 4 + 256;

constexpr float kGradTarget = 256.f;                                    // max|d_raw| is scaled to [128, 256]

__device__ __forceinline__ float grad_scale(const uint32_t* amax_bits) {
    float amax = __uint_as_float(*amax_bits);
    if (!(amax > 0.f) || !(amax < 3.0e38f)) return 1.f;
    return exp2f(floorf(log2f(kGradTarget / amax)));
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = fabsf(x[i]);
        if (v < 3.0e38f) m = fmaxf(m, v);              // ignore inf / nan
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));   // non-negative floats order like uints
}

__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }

// ------------------------------------------------------------------------------------
// 1. data-gradient chain (N=256 instructions, double-buffered accumulators, k-block pipelining: the forward's structure)
// ------------------------------------------------------------------------------------
constexpr int kC3Threads = 576;
constexpr uint32_t kC3ActHi = 0, kC3ActLo = 65536;
constexpr uint32_t kC3Ring = 131072;
constexpr int kC3Stages = 6;
constexpr uint32_t kC3Bars = kC3Ring + kC3Stages * kBlockBytes;       // 229376
constexpr uint32_t kC3TmemSlot = kC3Bars + 192;
constexpr uint32_t kC3Smem = kC3Bars + 256;
constexpr int kC3NumBlocks = 8 + 8 * 16;                              // step 0: K = 128, steps 1-8: K = 256

// block b of the chain stream: [256 rows (input feature j) x 16 k (output feature n)] of layer L = 9 - step
__global__ void __launch_bounds__(256)
pack_bwd_weights3_kernel(RawParams p, uint8_t* __restrict__ stream) {
    const int b = blockIdx.x;
    const int step = b < 8 ? 0 : 1 + (b - 8) / 16, kb = b < 8 ? b : (b - 8) % 16;
    const int L = 9 - step;
    const float* W = p.w[L];
    const int ld = p.ld[L], col0 = (L == 5) ? 63 : 0;
    uint8_t* dst = stream + ((size_t)blockIdx.y * kC3NumBlocks + b) * kBlockBytes;      // blockIdx.y: replica
    for (int u = threadIdx.x; u < 512; u += 256) {
        const int r = u & 255, kg = u >> 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = W[(size_t)(kb * 16 + kg * 8 + e) * ld + col0 + r];
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pack2(v[2 * e], v[2 * e + 1], h[e], l[e]);
        size_t off = (size_t)kg * 4096 + (size_t)r * 16;
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + kBlockHalfBytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

__device__ __forceinline__ void st_global_v4g(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld8g(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// one k-group of G: hi/lo words into the SMEM operand tile (the MMA warp streams finished k-blocks to the record)
template <bool kLo>
__device__ __forceinline__ void emit_g(uint32_t hi_base, uint32_t lo_base, uint32_t row, uint32_t kg, const float* v) {
    uint32_t h[4], l[4];
    const uint32_t off = kg * kLBO + row * 16;
    if (kLo) {
#pragma unroll
        for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
        st_shared_v4(hi_base + off, h[0], h[1], h[2], h[3]);
        st_shared_v4(lo_base + off, l[0], l[1], l[2], l[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { __half2 t = __floats2half2_rn(v[2 * i], v[2 * i + 1]); h[i] = *reinterpret_cast<uint32_t*>(&t); }
        st_shared_v4(hi_base + off, h[0], h[1], h[2], h[3]);
    }
}

__device__ unsigned long long g_profc[16];
#define PROFC_T0() long long pt0__ = kProf ? clock64() : 0
#define PROFC_ADD(var) do { if (kProf) { long long t__ = clock64(); var += t__ - pt0__; } } while (0)

// kProf: instrumented instantiation for the phase profile (cnerf_debug_profile_chain); the production kernel carries none of it
// kTerms: 3 = G W with the full hi/lo split (3 MMAs per MAC); 1 = fp16 operands, one MMA per MAC (only the hi halves of the
//         weight blocks are fetched).  kSaveLo: the gradient record also carries the lo halves of G (three-term dW).
template <bool kProf, int kTerms, bool kSaveLo>
__global__ void __launch_bounds__(kC3Threads, 1)
mlp_bwd_data3_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc, const float* __restrict__ d_raw,
                     const uint8_t* __restrict__ acts, const uint32_t* __restrict__ amax_bits, int n_points,
                     uint8_t* __restrict__ grads) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + kC3Bars, bar_empty = bar_full + 8 * kC3Stages;
    const uint32_t bar_dfull = bar_empty + 8 * kC3Stages;       // [2]
    const uint32_t bar_aready = bar_dfull + 16;                  // [8]  16 arrivals each (one per epilogue warp)
    const uint32_t bar_sdone = bar_aready + 64;                  //      the tile's last G stores have left shared memory
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kC3TmemSlot);
    const int num_tiles = (n_points + (int)kRows - 1) / (int)kRows;
    wstream += (size_t)(blockIdx.x % kWeightReplicas) * kC3NumBlocks * kBlockBytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kC3Stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_dfull, 1); mbar_init(bar_dfull + 8, 1);
        for (int k = 0; k < 8; ++k) mbar_init(bar_aready + 8 * k, 16);
        mbar_init(bar_sdone, 1);
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc(sbase + kC3TmemSlot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        // ===== weight loader =====
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                for (int b = 0; b < kC3NumBlocks; ++b, ++it) {
                    const uint32_t s = it % kC3Stages, ph = (it / kC3Stages) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    constexpr uint32_t kFetch = kTerms == 1 ? kBlockHalfBytes : kBlockBytes;
                    mbar_arrive_expect_tx(bar_full + 8 * s, kFetch);
                    bulk_g2s_hint(sbase + kC3Ring + s * kBlockBytes, wstream + (size_t)b * kBlockBytes, kFetch, bar_full + 8 * s, keep);
                }
        }
    } else if (warp == 17) {
        // ===== MMA issuer (warp-uniform walk, one elected lane issues) =====
        constexpr uint32_t idesc = instr_desc(128, 256);
        const uint64_t b256 = smem_desc_any(sbase + kC3Ring, 4096, 128);
        const uint64_t act_hi = smem_desc(sbase + kC3ActHi), act_lo = smem_desc(sbase + kC3ActLo);
        constexpr uint32_t kStep = 2 * (kLBO >> 4);
        const uint64_t stream_pol = l2_policy_evict_first();
        uint32_t it = 0;
        int tl = 0;
        long long pw_a = 0, pw_full = 0, pw_issue = 0, p_start = kProf ? clock64() : 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            uint8_t* grec = grads + (size_t)tile * kGTileBytes;
            // every k-block of an operand tile (G of some layer) is streamed to the gradient record as soon as it is final
            auto store_kblock = [&](int layer, int kb) {
                uint8_t* slot = grec + g_slot(layer);
                const size_t lo_off = layer == 9 ? 32768 : 65536;
                bulk_s2g_hint(slot + (size_t)kb * 8192, sbase + kC3ActHi + kb * 8192, 8192, stream_pol);
                if (kSaveLo) bulk_s2g_hint(slot + lo_off + (size_t)kb * 8192, sbase + kC3ActLo + kb * 8192, 8192, stream_pol);
                bulk_commit();
            };
#pragma unroll 1
            for (int step = 0; step < 9; ++step) {
                const uint32_t d = tmem + (uint32_t)(step & 1) * 256;
                const int nb = step == 0 ? 8 : 16;
#pragma unroll 1
                for (int j = 0; j < nb; ++j, ++it) {
                    if (!(j & 1)) {
                        const int kb = j >> 1;
                        const uint32_t aph = kb < 4 ? (uint32_t)(tl * 10 + step) & 1 : (uint32_t)(tl * 9 + step - 1) & 1;
                        { PROFC_T0(); mbar_wait(bar_aready + 8 * kb, aph); PROFC_ADD(pw_a); }
                        if (elect_one()) store_kblock(9 - step, kb);
                        __syncwarp();
                    }
                    const uint32_t s = it % kC3Stages, ph = (it / kC3Stages) & 1;
                    { PROFC_T0(); mbar_wait(bar_full + 8 * s, ph); PROFC_ADD(pw_full); }
                    tc_fence_after();
                    PROFC_T0();
                    if (elect_one()) {
                        const uint64_t bh = b256 + (uint64_t)(s * (kBlockBytes >> 4)), bl = bh + (kBlockHalfBytes >> 4);
                        const uint64_t ah = act_hi + (uint64_t)(j * kStep), al = act_lo + (uint64_t)(j * kStep);
                        umma_f16(d, ah, bh, idesc, j == 0 ? 0u : 1u);
                        if (kTerms == 3) { umma_f16(d, ah, bl, idesc, 1u); umma_f16(d, al, bh, idesc, 1u); }
                        umma_commit(bar_empty + 8 * s);
                        if (j + 1 == nb) {
                            bulk_wait_read0();                  // the epilogue overwrites the operand tile once it sees this step done
                            umma_commit(bar_dfull + 8 * (step & 1));
                        }
                    }
                    __syncwarp();
                    PROFC_ADD(pw_issue);
                }
            }
            // G0, written by the last epilogue, has no consumer here: store it and release the tile to the next prologue
#pragma unroll 1
            for (int kb = 0; kb < 8; ++kb) {
                mbar_wait(bar_aready + 8 * kb, kb < 4 ? (uint32_t)(tl * 10 + 9) & 1 : (uint32_t)(tl * 9 + 8) & 1);
                if (elect_one()) store_kblock(0, kb);
                __syncwarp();
            }
            if (elect_one()) { bulk_wait_read0(); mbar_arrive(bar_sdone); }
            __syncwarp();
        }
        if (elect_one()) bulk_wait0();
        __syncwarp();
        if (kProf && lane == 0) {
            atomicAdd(&g_profc[0], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_profc[1], (unsigned long long)pw_a); atomicAdd(&g_profc[3], (unsigned long long)pw_full);
            atomicAdd(&g_profc[4], (unsigned long long)pw_issue);
        }
    } else {
        // ===== prologue + epilogue warps: thread = (row, p); per 32-column k-block it owns columns 8p..8p+7 =====
        const int q = warp & 3, p = warp >> 2;
        const uint32_t row = (uint32_t)(q * 32 + lane);
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t ah = sbase + kC3ActHi, al = sbase + kC3ActLo;
        const float scale = grad_scale(amax_bits);
        int tl = 0;
        long long pw_d = 0, pw_s = 0, p_start = kProf ? clock64() : 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            const int grow = tile * (int)kRows + (int)row;
            const bool valid = grow < n_points;
            const uint8_t* arec = acts + (size_t)tile * kTileBytes;
            float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) dr = __ldg(reinterpret_cast<const float4*>(d_raw) + grow);
            dr.x *= scale; dr.y *= scale; dr.z *= scale; dr.w *= scale;
            if (tl > 0) { PROFC_T0(); mbar_wait(bar_sdone, (uint32_t)(tl - 1) & 1); PROFC_ADD(pw_s); }      // the previous tile's G0 has left the operand tile
            // G9 = (d_rgb W_rgb) * [hv > 0]: four k-blocks of 32 columns
            // ReLU sign bits of the views layer: the row's sixteen k-group bytes in one load
            const uint4 mrow = __ldg(reinterpret_cast<const uint4*>(arec + kSlotM + 32768 + row * 16));
            const uint32_t mword[4] = {mrow.x, mrow.y, mrow.z, mrow.w};
#pragma unroll
            for (uint32_t kb = 0; kb < 4; ++kb) {
                const uint32_t kg = kb * 4 + (uint32_t)p, c = kg * 8;
                const uint32_t mbits = mword[kb] >> (8 * (uint32_t)p);          // byte kg = 4 kb + p of the row
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float gsum = dr.x * __ldg(misc + kMiscRgbW + c + j) + dr.y * __ldg(misc + kMiscRgbW + 128 + c + j) +
                                 dr.z * __ldg(misc + kMiscRgbW + 256 + c + j);
                    v[j] = ((mbits >> j) & 1u) ? clamp_h(gsum) : 0.f;
                }
                emit_g<kTerms == 3 || kSaveLo>(ah, al, row, kg, v);
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_aready + 8 * kb);
            }
#pragma unroll 1
            for (int step = 0; step < 9; ++step) {
                const int L = 9 - step;                           // D = gradient w.r.t. the input of layer L = G_{L-1} before masking
                const uint8_t* msk = arec + kSlotM + (size_t)(L - 1) * 4096;         // ReLU sign bits of h_{L-1} (L <= 8)
                // the sign bits do not depend on this step's MMAs: the load is issued BEFORE waiting for the accumulator, so its HBM
                // latency (the record is not L2 resident) hides behind the tensor-core work of the step
                uint2 mk8 = make_uint2(0u, 0u);                                        // this thread's eight bytes: k-blocks 0-3 | 4-7
                if (L != 9) mk8 = __ldg(reinterpret_cast<const uint2*>(msk + (size_t)p * 1024 + row * 8));
                { PROFC_T0(); mbar_wait(bar_dfull + 8 * (step & 1), (uint32_t)(tl * ((step & 1) ? 4 : 5) + (step >> 1)) & 1); PROFC_ADD(pw_d); }
                tc_fence_after();
                const uint32_t dcol = t_lane + (uint32_t)(step & 1) * 256 + (uint32_t)p * 8;
#pragma unroll
                for (uint32_t kb = 0; kb < 8; kb += 2) {
                    float v[16];
                    const uint32_t mw = kb < 4 ? mk8.x : mk8.y;
                    const uint32_t m[2] = {mw >> (8 * (kb & 3)), mw >> (8 * ((kb + 1) & 3))};
                    tmem_ld8g(dcol + kb * 32, v);
                    tmem_ld8g(dcol + kb * 32 + 32, v + 8);
                    tmem_ld_wait();
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t kg = (kb + u) * 4 + (uint32_t)p, c = kg * 8;
                        float* w = v + 8 * u;
                        if (L == 8) {                             // + d_sigma * alpha_linear.weight
                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c)), a1 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c + 4));
                            w[0] = fmaf(dr.w, a0.x, w[0]); w[1] = fmaf(dr.w, a0.y, w[1]); w[2] = fmaf(dr.w, a0.z, w[2]); w[3] = fmaf(dr.w, a0.w, w[3]);
                            w[4] = fmaf(dr.w, a1.x, w[4]); w[5] = fmaf(dr.w, a1.y, w[5]); w[6] = fmaf(dr.w, a1.z, w[6]); w[7] = fmaf(dr.w, a1.w, w[7]);
                        }
                        if (L != 9) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) w[j] = ((m[u] >> j) & 1u) ? clamp_h(w[j]) : 0.f;
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) w[j] = clamp_h(w[j]);
                        }
                        emit_g<kTerms == 3 || kSaveLo>(ah, al, row, kg, w);
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_aready + 8 * (kb + u));
                    }
                }
            }
            tc_fence_before();
        }
        if (kProf && lane == 0 && warp == 0) {
            atomicAdd(&g_profc[8], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_profc[9], (unsigned long long)pw_d); atomicAdd(&g_profc[10], (unsigned long long)pw_s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
// 1b. data-gradient chain with fp16 operands (chain_terms == 1): one MMA per MAC and TWO tiles in flight per SM -- the structure of
//     the fp16 forward (mlp_fwd5.cu).  With hi halves only a G tile is 64 KB, two fit, and the MMA warp alternates between them step by
//     step: while the epilogue warps turn tile X's accumulator into G_{L-1} (mask, clamp, fp16), the tensor pipe runs tile Y's step.
//     SMEM: 2 x 64 KB G tiles + 8-stage ring of 8 KB weight units (the hi halves of the chain stream's blocks); TMEM: one 256-column
//     accumulator per slot.  Per slot: a_ready (16 warp arrivals: G tile complete, accumulator read), d_full (tcgen05.commit), s_done
//     (the tile's G0 has left shared memory).  The gradient record receives the hi halves only (dw_terms == 1).
// ------------------------------------------------------------------------------------
constexpr int kC5Threads = 576;
constexpr uint32_t kC5Act = 0;                                        // + slot * 65536
constexpr uint32_t kC5Ring = 131072;
constexpr int kC5Stages = 8;
constexpr uint32_t kC5Unit = kBlockHalfBytes;
constexpr uint32_t kC5Bars = kC5Ring + kC5Stages * kC5Unit;           // 196608
constexpr uint32_t kC5TmemSlot = kC5Bars + 192;
constexpr uint32_t kC5Smem = kC5Bars + 256;

__device__ __forceinline__ void emit_g_hi(uint32_t base, uint32_t row, uint32_t kg, const float* v) {
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __half2 t = __floats2half2_rn(v[2 * i], v[2 * i + 1]); h[i] = *reinterpret_cast<uint32_t*>(&t); }
    st_shared_v4(base + kg * kLBO + row * 16, h[0], h[1], h[2], h[3]);
}

__global__ void __launch_bounds__(kC5Threads, 1)
mlp_bwd_data5_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc, const float* __restrict__ d_raw,
                     const uint8_t* __restrict__ acts, const uint32_t* __restrict__ amax_bits, int n_points,
                     uint8_t* __restrict__ grads) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + kC5Bars, bar_empty = bar_full + 8 * kC5Stages;
    const uint32_t bar_dfull = bar_empty + 8 * kC5Stages;       // [2]
    const uint32_t bar_aready = bar_dfull + 16;                  // [2]  16 warp arrivals
    const uint32_t bar_sdone = bar_aready + 16;                  // [2]
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kC5TmemSlot);
    const int num_tiles = (n_points + (int)kRows - 1) / (int)kRows;
    const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_iter = (my_tiles + 1) / 2;                       // slot s works on this CTA's tiles 2 it + s

    if (threadIdx.x == 0) {
        for (int s = 0; s < kC5Stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_aready + 8 * s, 16); mbar_init(bar_sdone + 8 * s, 1); }
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc(sbase + kC5TmemSlot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        // ===== weight loader: the hi half of every block, in the order the MMA warp consumes them =====
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            uint32_t st = 0, ph = 0;
            for (int it = 0; it < n_iter; ++it)
                for (int m = 0; m < 9; ++m)
                    for (int s = 0; s < 2; ++s) {
                        if (2 * it + s >= my_tiles) continue;
                        const uint8_t* src = wstream + (size_t)(m == 0 ? 0 : 8 + (m - 1) * 16) * kBlockBytes;
                        for (int nb = m == 0 ? 8 : 16; nb > 0; --nb, src += kBlockBytes) {
                            mbar_wait(bar_empty + 8 * st, ph ^ 1);
                            mbar_arrive_expect_tx(bar_full + 8 * st, kC5Unit);
                            bulk_g2s_hint(sbase + kC5Ring + st * kC5Unit, src, kC5Unit, bar_full + 8 * st, keep);
                            st = (st + 1) & (kC5Stages - 1);
                            ph ^= (st == 0);
                        }
                    }
        }
    } else if (warp == 17) {
        // ===== MMA issuer: alternates between the two slots step by step =====
        constexpr uint32_t idesc = instr_desc(128, 256);
        constexpr uint32_t kStep = 2 * (kLBO >> 4);
        const uint64_t b256 = smem_desc_any(sbase + kC5Ring, 4096, 128);
        const uint64_t stream_pol = l2_policy_evict_first();
        uint32_t st = 0, ph = 0;
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll 1
            for (int p = 0; p < 10; ++p) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const int n = 2 * it + s;
                    if (n >= my_tiles) continue;
                    const int tile = (int)blockIdx.x + n * (int)gridDim.x;
                    const uint32_t act = sbase + kC5Act + (uint32_t)s * 65536;
                    const uint32_t d = tmem + (uint32_t)s * 256;
                    mbar_wait(bar_aready + 8 * s, (uint32_t)(it * 10 + p) & 1);
                    tc_fence_after();
                    if (elect_one()) {      // G_{9-p} of this tile is final: stream it to the gradient record (hi halves)
                        bulk_s2g_hint(grads + (size_t)tile * kGTileBytes + g_slot(9 - p), act, p == 0 ? 32768u : 65536u, stream_pol);
                        bulk_commit();
                    }
                    __syncwarp();
                    if (p < 9) {
                        uint64_t a = smem_desc(act);
                        uint32_t acc = 0u;
#pragma unroll 1
                        for (int j = p == 0 ? 8 : 16; j > 0; --j, a += kStep) {
                            mbar_wait(bar_full + 8 * st, ph);
                            tc_fence_after();
                            if (elect_one()) {
                                umma_f16(d, a, b256 + (uint64_t)(st * (kC5Unit >> 4)), idesc, acc);
                                umma_commit(bar_empty + 8 * st);
                            }
                            __syncwarp();
                            acc = 1u;
                            st = (st + 1) & (kC5Stages - 1);
                            ph ^= (st == 0);
                        }
                        if (elect_one()) {
                            bulk_wait_read0();                      // the epilogue overwrites the G tile once it sees this step done
                            umma_commit(bar_dfull + 8 * s);
                        }
                    } else if (elect_one()) {                       // G0 has no consumer here: release the slot to the next tile's prologue
                        bulk_wait_read0();
                        mbar_arrive(bar_sdone + 8 * s);
                    }
                    __syncwarp();
                }
            }
        }
        if (elect_one()) bulk_wait0();
        __syncwarp();
    } else {
        // ===== prologue + epilogue warps: thread = (row, q-lane, p); per 32-column k-block it owns columns 8p..8p+7; both slots in turn =====
        const int q = warp & 3, p = warp >> 2;
        const uint32_t row = (uint32_t)(q * 32 + lane);
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const float scale = grad_scale(amax_bits);
        float4 dr0 = make_float4(0.f, 0.f, 0.f, 0.f), dr1 = dr0;      // scaled d_raw row of the tile in slot 0 / 1
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll 1
            for (int ph10 = 0; ph10 < 10; ++ph10) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const int n = 2 * it + s;
                    if (n >= my_tiles) continue;
                    const int tile = (int)blockIdx.x + n * (int)gridDim.x;
                    const int grow = tile * (int)kRows + (int)row;
                    const uint8_t* arec = acts + (size_t)tile * kTileBytes;
                    const uint32_t ab = sbase + kC5Act + (uint32_t)s * 65536;
                    if (ph10 == 0) {
                        // G9 = (d_rgb W_rgb) * [hv > 0]: K = 128, four k-blocks of 32 columns
                        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (grow < n_points) dr = __ldg(reinterpret_cast<const float4*>(d_raw) + grow);
                        dr.x *= scale; dr.y *= scale; dr.z *= scale; dr.w *= scale;
                        if (s) dr1 = dr; else dr0 = dr;
                        const uint4 mrow = __ldg(reinterpret_cast<const uint4*>(arec + kSlotM + 32768 + row * 16));
                        const uint32_t mword[4] = {mrow.x, mrow.y, mrow.z, mrow.w};
                        if (it > 0) mbar_wait(bar_sdone + 8 * s, (uint32_t)(it - 1) & 1);      // the slot's previous G0 has left the tile
#pragma unroll
                        for (uint32_t kb = 0; kb < 4; ++kb) {
                            const uint32_t kg = kb * 4 + (uint32_t)p, c = kg * 8;
                            const uint32_t mbits = mword[kb] >> (8 * (uint32_t)p);
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float gsum = dr.x * __ldg(misc + kMiscRgbW + c + j) + dr.y * __ldg(misc + kMiscRgbW + 128 + c + j) +
                                                   dr.z * __ldg(misc + kMiscRgbW + 256 + c + j);
                                v[j] = ((mbits >> j) & 1u) ? clamp_h(gsum) : 0.f;
                            }
                            emit_g_hi(ab, row, kg, v);
                        }
                    } else {
                        const int m = ph10 - 1, L = 9 - m;               // accumulator = gradient w.r.t. the input of layer L = G_{L-1} before masking
                        const float drw = s ? dr1.w : dr0.w;
                        uint2 mk8 = make_uint2(0u, 0u);                    // ReLU sign bits of h_{L-1}: this thread's eight bytes (k-blocks 0-3 | 4-7)
                        if (L != 9) mk8 = __ldg(reinterpret_cast<const uint2*>(arec + kSlotM + (size_t)(L - 1) * 4096 + (size_t)p * 1024 + row * 8));
                        mbar_wait(bar_dfull + 8 * s, (uint32_t)(it * 9 + m) & 1);
                        tc_fence_after();
                        const uint32_t dcol = t_lane + (uint32_t)s * 256 + (uint32_t)p * 8;
#pragma unroll
                        for (uint32_t half = 0; half < 2; ++half) {
                            float v[32];
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4) tmem_ld8g(dcol + (half * 4 + k4) * 32, v + 8 * k4);
                            tmem_ld_wait();
                            const uint32_t mw = half == 0 ? mk8.x : mk8.y;
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4) {
                                const uint32_t kg = (half * 4 + k4) * 4 + (uint32_t)p, c = kg * 8;
                                float* w = v + 8 * k4;
                                if (L == 8) {                             // + d_sigma * alpha_linear.weight
                                    const float4 a0 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c)), a1 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c + 4));
                                    w[0] = fmaf(drw, a0.x, w[0]); w[1] = fmaf(drw, a0.y, w[1]); w[2] = fmaf(drw, a0.z, w[2]); w[3] = fmaf(drw, a0.w, w[3]);
                                    w[4] = fmaf(drw, a1.x, w[4]); w[5] = fmaf(drw, a1.y, w[5]); w[6] = fmaf(drw, a1.z, w[6]); w[7] = fmaf(drw, a1.w, w[7]);
                                }
                                const uint32_t mb = L != 9 ? (mw >> (8 * k4)) : 0xffu;
#pragma unroll
                                for (int j = 0; j < 8; ++j) w[j] = ((mb >> j) & 1u) ? clamp_h(w[j]) : 0.f;
                                emit_g_hi(ab, row, kg, w);
                            }
                        }
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_aready + 8 * s);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------
// 2. weight gradients: D[n_out, k_in] = sum_p G[p, n_out] X[p, k_in]
// ------------------------------------------------------------------------------------
struct DwSrc { uint32_t slot_off, kgroups, lo_off; };        // hi k-groups at slot_off, lo k-groups at slot_off + lo_off
// The records are 2-D arrays of 2048-byte k-group rows (128 points x 16 B); a stage takes the 512-byte column
// window of one 32-point quarter from all k-group rows of a source with ONE TMA tensor copy (box = rows x 512 B)
// -- 128 separate 512-byte bulk copies per stage were TMA-issue bound (ncu: 8.4k cycles per stage).
struct DwPass {
    int n_a, n_x;
    DwSrc a[2];            // G sources in the gradient record (kgroups 32 -> two 128-row halves, 16 -> one)
    DwSrc x[2];            // X sources in the activation record (N = 8 * kgroups)
    int db_mask;           // bit i: sum bias gradient of a[i]
    int terms;             // 3: G and X as fp16 hi + lo, three MMAs per MAC; 1: hi halves only (fp16 operands), one MMA per MAC
};

constexpr int kDwStages = 3;                                  // three-term mode; the fp16 mode runs 2 x kDwStages half-size stages
constexpr uint32_t kDwStageBytes = 73728;                     // 72 KB: two 256-wide G quarters + the 64-wide encoding (hi + lo)
constexpr uint32_t kDwBars = kDwStages * kDwStageBytes;       // 221184
constexpr uint32_t kDwSmem = kDwBars + 128;
constexpr int kDwThreads = 320;                               // 8 reduction/epilogue warps, loader warp, MMA warp
constexpr uint32_t kQuarter = 512;                            // bytes of one k-group for 32 points

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t x, uint32_t y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar) : "memory");
}

// All passes of one network in ONE persistent launch: the TMA ring keeps streaming across pass boundaries (while the epilogue warps
// drain a pass's accumulators the loader already fetches the next pass's first stages), and the per-launch prologue (TMEM
// allocation, barrier init, tensor-map fetch, pipeline fill) is paid once instead of ten times -- the coarse network's passes
// move only ~130 MB each in fp16 mode, where those fixed costs were a third of the pass.
struct alignas(64) DwAll {
    CUtensorMap map[kDwPasses][3];      // per pass: G record, X source 0, X source 1
    DwPass P[kDwPasses];
    int n_pass;
};

struct DwLayout { uint32_t a_off[2], x_off[2], stage_bytes, nhalf, qb, ksteps, per_tile; };
__device__ __forceinline__ DwLayout dw_layout(const DwPass& P) {
    // three-term mode: a stage holds 32 points of every source as [hi | lo] k-group rows of 512 B; fp16 mode: 64 points of the
    // hi halves as rows of 1024 B (the same stage size: half the barrier hand-shakes per byte)
    DwLayout L;
    L.nhalf = P.terms == 1 ? 1u : 2u;
    L.qb = P.terms == 1 ? 2 * kQuarter : kQuarter;
    L.ksteps = L.qb / 256;
    L.per_tile = 128 * 16 / L.qb;
    uint32_t off = 0;
    for (int i = 0; i < 2; ++i) { L.a_off[i] = off; if (i < P.n_a) off += L.nhalf * P.a[i].kgroups * L.qb; }
    for (int j = 0; j < 2; ++j) { L.x_off[j] = off; if (j < P.n_x) off += L.nhalf * P.x[j].kgroups * L.qb; }
    L.stage_bytes = off;
    return L;
}

__global__ void __launch_bounds__(kDwThreads, 1)
mlp_bwd_weight_kernel(const __grid_constant__ DwAll A, int num_tiles, float* __restrict__ dw_part_all, float* __restrict__ db_part_all) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + kDwBars, bar_empty = bar_full + 8 * kDwStages, bar_acc = bar_empty + 8 * kDwStages;
    const uint32_t bar_drained = bar_acc + 8;                    // 8 warp arrivals: the pass's accumulators have left tensor memory
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kDwBars + 112);

    // contiguous tile range of this CTA
    const int per = num_tiles / gridDim.x, rem = num_tiles % gridDim.x;
    const int t0 = blockIdx.x * per + min((int)blockIdx.x, rem), t1 = t0 + per + ((int)blockIdx.x < rem ? 1 : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDwStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 9); }
        mbar_init(bar_acc, 1);
        mbar_init(bar_drained, 8);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(sbase + kDwBars + 112, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ===== loader: per (tile, point window) one stage; one TMA box copy per source (two when hi / lo rows are apart) =====
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int pass = 0; pass < A.n_pass; ++pass) {
                const DwPass& P = A.P[pass];
                const DwLayout L = dw_layout(P);
                const int n_stage_iters = (t1 - t0) * (int)L.per_tile;
                for (int it = 0; it < n_stage_iters; ++it) {
                    const int tile = t0 + it / (int)L.per_tile, q = it % (int)L.per_tile;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(bar_full + 8 * s, L.stage_bytes);
                    const uint32_t dst0 = sbase + s * kDwStageBytes;
                    for (int i = 0; i < P.n_a + P.n_x; ++i) {
                        const bool is_a = i < P.n_a;
                        const DwSrc src = is_a ? P.a[i] : P.x[i - P.n_a];
                        const CUtensorMap* map = &A.map[pass][is_a ? 0 : 1 + (i - P.n_a)];
                        const uint32_t row0 = (uint32_t)tile * (uint32_t)((is_a ? kGTileBytes : kTileBytes) / 2048) + src.slot_off / 2048;
                        const uint32_t d = dst0 + (is_a ? L.a_off[i] : L.x_off[i - P.n_a]);
                        // box = rows x 256 elements: u16 elements (512 B, 32 points) in three-term mode, u32 (1024 B, 64 points) in fp16 mode
                        tma_load_2d(d, map, q * 256, row0, bar_full + 8 * s);                 // box rows = 2*kgroups or kgroups
                        if (L.nhalf == 2 && src.lo_off != src.kgroups * 2048)
                            tma_load_2d(d + src.kgroups * kQuarter, map, q * 256, row0 + src.lo_off / 2048, bar_full + 8 * s);
                    }
                    if (++s == kDwStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int pass = 0; pass < A.n_pass; ++pass) {
                const DwPass& P = A.P[pass];
                const DwLayout L = dw_layout(P);
                const int n_stage_iters = (t1 - t0) * (int)L.per_tile;
                if (pass > 0) { mbar_wait(bar_drained, (uint32_t)(pass - 1) & 1); tc_fence_after(); }      // tensor memory is free again
                for (int it = 0; it < n_stage_iters; ++it) {
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t st = sbase + s * kDwStageBytes;
#pragma unroll 1
                    for (uint32_t ks = 0; ks < L.ksteps; ++ks) {
                        uint32_t col = 0;
                        for (int i = 0; i < P.n_a; ++i) {
                            const uint32_t halves = P.a[i].kgroups / 16, a_lo_off = P.a[i].kgroups * L.qb;
                            for (uint32_t h = 0; h < halves; ++h) {
                                const uint32_t a_hi = st + L.a_off[i] + h * 16 * L.qb + ks * 256;
                                const uint64_t ah = smem_desc_any(a_hi, 128, L.qb), al = smem_desc_any(a_hi + a_lo_off, 128, L.qb);
                                for (int j = 0; j < P.n_x; ++j) {
                                    const uint32_t N = P.x[j].kgroups * 8;
                                    const uint32_t x_hi = st + L.x_off[j] + ks * 256;
                                    const uint64_t xh = smem_desc_any(x_hi, 128, L.qb), xl = smem_desc_any(x_hi + P.x[j].kgroups * L.qb, 128, L.qb);
                                    const uint32_t idesc = instr_desc_mn(128, N);
                                    umma_f16(tmem + col, ah, xh, idesc, (it == 0 && ks == 0) ? 0u : 1u);
                                    if (L.nhalf == 2) { umma_f16(tmem + col, ah, xl, idesc, 1u); umma_f16(tmem + col, al, xh, idesc, 1u); }
                                    col += N;
                                }
                            }
                        }
                    }
                    umma_commit(bar_empty + 8 * s);
                    if (++s == kDwStages) { s = 0; ph ^= 1; }
                }
                umma_commit(bar_acc);
            }
        }
    } else {
        // ===== bias-gradient column sums (from the same SMEM tiles), then the TMEM -> partial epilogue, pass by pass =====
        uint32_t s = 0, ph = 0;
        for (int pass = 0; pass < A.n_pass; ++pass) {
            const DwPass& P = A.P[pass];
            const DwLayout L = dw_layout(P);
            const int n_stage_iters = (t1 - t0) * (int)L.per_tile;
            float* dw_part = dw_part_all + (size_t)pass * kDwPassFloats;
            float* db_part = db_part_all + (size_t)pass * kDbPassFloats;
            float acc[2][32];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int k = 0; k < 32; ++k) acc[i][k] = 0.f;
            for (int it = 0; it < n_stage_iters; ++it) {
                mbar_wait(bar_full + 8 * s, ph);
                const uint32_t st = sbase + s * kDwStageBytes;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i < P.n_a && ((P.db_mask >> i) & 1)) {
                        const uint32_t per_warp = P.a[i].kgroups / 8;       // 4 (256 wide) or 2 (128 wide)
#pragma unroll
                        for (uint32_t g = 0; g < 4; ++g) {
                            if (g < per_warp) {
                                const uint32_t kg = warp * per_warp + g;
                                const uint32_t addr = st + L.a_off[i] + kg * L.qb + lane * 16;
                                uint4 hi, lo = make_uint4(0u, 0u, 0u, 0u);
                                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(addr));
                                if (L.nhalf == 2)      // three-term: the lo halves of the same 32 points; fp16: the second 32 points of the row
                                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(addr + P.a[i].kgroups * L.qb));
                                else
                                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(addr + 512));
                                const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                                    float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                                    acc[i][g * 8 + 2 * e] += a.x + b.x;
                                    acc[i][g * 8 + 2 * e + 1] += a.y + b.y;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * s);
                if (++s == kDwStages) { s = 0; ph ^= 1; }
            }
            // bias partials: reduce over the 32 point-lanes
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (i < P.n_a && ((P.db_mask >> i) & 1)) {
                    const uint32_t per_warp = P.a[i].kgroups / 8;
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        float v = warp_sum(acc[i][k]);
                        if (lane == 0 && (uint32_t)(k >> 3) < per_warp)
                            db_part[((size_t)blockIdx.x * 2 + i) * 256 + (warp * per_warp + (k >> 3)) * 8 + (k & 7)] = v;
                    }
                }
            }
            // accumulators -> per-CTA partial, block by block: [block][128 rows][N]
            mbar_wait(bar_acc, (uint32_t)pass & 1);
            tc_fence_after();
            float* part = dw_part + (size_t)blockIdx.x * 128 * 512;
            const uint32_t rowq = (uint32_t)(warp & 3) * 32, row = rowq + lane;
            uint32_t col = 0;
            for (int i = 0; i < P.n_a; ++i)
                for (uint32_t h = 0; h < P.a[i].kgroups / 16; ++h)
                    for (int j = 0; j < P.n_x; ++j) {
                        const uint32_t N = P.x[j].kgroups * 8;
                        float* blk = part + (size_t)col * 128;                 // block base: 128 x N floats
                        for (uint32_t c = (uint32_t)(warp >> 2) * 32; c < N; c += 64) {
                            float v[32];
                            tmem_ld32(tmem + (rowq << 16) + col + c, v);
                            tmem_ld_wait();
                            float4* o = reinterpret_cast<float4*>(blk + (size_t)row * N + c);
                            if (n_stage_iters == 0) {
#pragma unroll
                                for (int k = 0; k < 32; ++k) v[k] = 0.f;
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) o[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        }
                        col += N;
                    }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_drained);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, 512);
}

// dst[(row0 + r) * ld + col0 + c] (+)= inv_scale * sum_cta part[pass][cta][blk_off + r * blk_n + src_col0 + c]
// All passes write disjoint partial regions, so ONE launch reduces every segment of every pass (and one the biases).
struct DwSeg { float* dst; int ld, row0, col0, ncols, blk_off, blk_n, src_col0, pass; };
struct DwSegs { int n; DwSeg s[24]; };
struct DbSeg { float* dst; int n, pass, src; };
struct DbSegs { int n; DbSeg s[12]; };

__global__ void dw_reduce_kernel(DwSegs S, const float* __restrict__ part, int n_cta, const uint32_t* __restrict__ amax_bits,
                                 int accumulate) {
    const DwSeg sg = S.s[blockIdx.y];
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 128 * sg.ncols) return;
    int r = idx / sg.ncols, c = idx - r * sg.ncols;
    const float* p = part + (size_t)sg.pass * kDwPassFloats + sg.blk_off + (size_t)r * sg.blk_n + sg.src_col0 + c;
    float acc = 0.f;
    for (int k = 0; k < n_cta; ++k) acc += p[(size_t)k * 128 * 512];
    acc *= 1.f / grad_scale(amax_bits);
    float* d = sg.dst + (size_t)(sg.row0 + r) * sg.ld + sg.col0 + c;
    *d = accumulate ? *d + acc : acc;
}

__global__ void db_reduce_kernel(DbSegs S, const float* __restrict__ db_part, int n_cta, const uint32_t* __restrict__ amax_bits,
                                 int accumulate) {
    const DbSeg sg = S.s[blockIdx.y];
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= sg.n) return;
    const float* p = db_part + (size_t)sg.pass * kDbPassFloats + (size_t)sg.src * 256 + c;
    float acc = 0.f;
    for (int k = 0; k < n_cta; ++k) acc += p[(size_t)k * 512];
    acc *= 1.f / grad_scale(amax_bits);
    sg.dst[c] = accumulate ? sg.dst[c] + acc : acc;
}

// ------------------------------------------------------------------------------------
// 3. narrow heads in fp32: dW_rgb[ch][n] = sum d_rgb[p][ch] hv[p][n], dW_alpha[k] = sum d_sigma[p] h7[p][k]
// ------------------------------------------------------------------------------------
template <bool kLo>
__device__ __forceinline__ void load_hilo8(const uint8_t* hi_ptr, size_t lo_off, float* out) {
    uint4 hi = __ldg(reinterpret_cast<const uint4*>(hi_ptr));
    uint4 lo = make_uint4(0u, 0u, 0u, 0u);
    if (kLo) lo = __ldg(reinterpret_cast<const uint4*>(hi_ptr + lo_off));
    const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
        float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
        out[2 * e] = a.x + b.x; out[2 * e + 1] = a.y + b.y;
    }
}

template <bool kLo>      // kLo: the activation record carries the lo halves
__global__ void __launch_bounds__(256)
mlp_heads_grad_kernel(const float* __restrict__ d_raw, const uint8_t* __restrict__ acts, int n_points, float* __restrict__ part) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (n_points + 127) / 128;
    float rgb[3][16], al[32], bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; ++k) rgb[0][k] = rgb[1][k] = rgb[2][k] = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) al[k] = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint8_t* rec = acts + (size_t)tile * kTileBytes;
        for (int rg = 0; rg < 4; ++rg) {
            const int row = rg * 32 + lane, grow = tile * 128 + row;
            float4 dr = grow < n_points ? __ldg(reinterpret_cast<const float4*>(d_raw) + grow) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (warp == 0) { bsum[0] += dr.x; bsum[1] += dr.y; bsum[2] += dr.z; bsum[3] += dr.w; }
            float v[8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {                     // hv: 16 k-groups, two per warp
                load_hilo8<kLo>(rec + kSlotHV + (size_t)(warp * 2 + g) * 2048 + row * 16, 32768, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    rgb[0][g * 8 + e] = fmaf(dr.x, v[e], rgb[0][g * 8 + e]);
                    rgb[1][g * 8 + e] = fmaf(dr.y, v[e], rgb[1][g * 8 + e]);
                    rgb[2][g * 8 + e] = fmaf(dr.z, v[e], rgb[2][g * 8 + e]);
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {                     // h7: 32 k-groups, four per warp
                load_hilo8<kLo>(rec + kSlotH0 + 7 * 131072 + (size_t)(warp * 4 + g) * 2048 + row * 16, 65536, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) al[g * 8 + e] = fmaf(dr.w, v[e], al[g * 8 + e]);
            }
        }
    }
    float* out = part + (size_t)blockIdx.x * kHeadFloats;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float v = warp_sum(rgb[ch][k]);
            if (lane == 0) out[ch * 128 + warp * 16 + k] = v;
        }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        float v = warp_sum(al[k]);
        if (lane == 0) out[384 + warp * 32 + k] = v;
    }
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = warp_sum(bsum[k]);
            if (lane == 0) out[640 + k] = v;
        }
    }
}

__global__ void heads_reduce_kernel(const float* __restrict__ part, int n_cta, float* __restrict__ d_rgb_w,
                                    float* __restrict__ d_rgb_b, float* __restrict__ d_alpha_w, float* __restrict__ d_alpha_b,
                                    int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kHeadFloats) return;
    float acc = 0.f;
    for (int k = 0; k < n_cta; ++k) acc += part[(size_t)k * kHeadFloats + i];
    float* d = i < 384 ? d_rgb_w + i : i < 640 ? d_alpha_w + (i - 384) : i < 643 ? d_rgb_b + (i - 640) : d_alpha_b;
    *d = accumulate ? *d + acc : acc;
}

}  // namespace cnerf

using namespace cnerf;

// called from cnerf_weights_refresh (mlp_tc.cu)
namespace cnerf {
int bwd_stream3_blocks() { return kC3NumBlocks; }
int pack_bwd_stream3(const RawParams& p, uint8_t* stream, cudaStream_t st) {
    pack_bwd_weights3_kernel<<<dim3(kC3NumBlocks, kWeightReplicas), 256, 0, st>>>(p, stream);
    CNERF_LAUNCH_CHECK("pack_bwd_weights3_kernel");
    return CNERF_OK;
}
}  // namespace cnerf

extern "C" int64_t cnerf_mlp_grads_bytes(int64_t n_points) { return ceil_div64(n_points, kRows) * (int64_t)kGTileBytes; }
extern "C" int64_t cnerf_mlp_bwd_workspace_bytes(void) { return (int64_t)kWsBytes; }

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// record buffer as [rows][1024 x u16] (2048-byte k-group rows); box = box_rows x 256 u16 (one 32-point quarter)
// wide: the same rows as [512 x u32], box = box_rows x 256 u32 (two quarters = 64 points per copy; a box dimension is at most 256 elements)
int make_record_map(CUtensorMap* m, const void* base, uint64_t rows, uint32_t box_rows, bool wide) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return set_error(CNERF_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {wide ? 512u : 1024u, rows};
    cuuint64_t gstride[1] = {2048};
    cuuint32_t box[2] = {256, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, wide ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(CNERF_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CNERF_OK;
}
struct BwdCtx {
    cudaStream_t st; uint32_t* amax; float *dw_part, *db_part, *head_part; const uint8_t* a; uint8_t* g; int tiles, grid;
};
int bwd_ctx(const void* acts, void* grads_rec, int n_points, void* workspace, void* stream, BwdCtx* c) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_bwd_data3_kernel<false, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_bwd_data3_kernel<false, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_bwd_data3_kernel<false, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_bwd_data3_kernel<true, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_bwd_data3_kernel<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC3Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_bwd_data5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC5Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDwSmem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(mlp_bwd kernels)");
        attr_set = true;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    c->st = as_stream(stream);
    c->amax = reinterpret_cast<uint32_t*>(ws + kWsAmax);
    c->dw_part = reinterpret_cast<float*>(ws + kWsDwPart);
    c->db_part = reinterpret_cast<float*>(ws + kWsDbPart);
    c->head_part = reinterpret_cast<float*>(ws + kWsHeadPart);
    c->a = reinterpret_cast<const uint8_t*>(acts);
    c->g = reinterpret_cast<uint8_t*>(grads_rec);
    c->tiles = ceil_div(n_points, (int)kRows);
    c->grid = c->tiles < kNumSMs ? c->tiles : kNumSMs;
    return CNERF_OK;
}
}  // namespace

static int g_profc_host = 0;
// Debug: in-kernel phase profile of mlp_bwd_data3_kernel (cycles summed over CTAs):
//  [0] MMA warp total  [1] wait operand k-blocks  [3] wait weights  [4] MMA issue + commit  [8] epilogue total  [9] wait D  [10] wait G0 stored
extern "C" int cnerf_debug_profile_chain(int enable, unsigned long long* out16) {
    unsigned long long zero[16] = {0};
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out16) e = cudaMemcpyFromSymbol(out16, g_profc, sizeof(zero));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_profc, zero, sizeof(zero));
    if (e != cudaSuccess) return check_cuda(e, "cnerf_debug_profile_chain");
    g_profc_host = enable & 1;
    return CNERF_OK;
}

// Stage 1: gradient scale + data-gradient chain -> grads_rec (G tiles of every layer).
static int check_terms(const char* who, int chain_terms, int dw_terms) {
    CNERF_REQUIRE(chain_terms == 1 || chain_terms == 3, "%s: chain_terms must be 1 (fp16 operands) or 3 (hi/lo split)", who);
    CNERF_REQUIRE(dw_terms == 1 || dw_terms == 3, "%s: dw_terms must be 1 (fp16 operands) or 3 (hi/lo split)", who);
    CNERF_REQUIRE(!(chain_terms == 1 && dw_terms == 3), "%s: a one-term chain does not produce the lo halves a three-term dW reads", who);
    return CNERF_OK;
}

extern "C" int cnerf_mlp_bwd_data(const cnerf_weights* w, const float* d_raw, const void* acts, void* grads_rec,
                                  int n_points, int chain_terms, int dw_terms, void* workspace, void* stream) {
    CNERF_REQUIRE(w && w->packed && w->stream_bwd3, "cnerf_mlp_bwd_data: weights handle not packed");
    if (int rc = check_terms("cnerf_mlp_bwd_data", chain_terms, dw_terms)) return rc;
    CNERF_REQUIRE(d_raw && acts && grads_rec && workspace, "cnerf_mlp_bwd_data: null pointer");
    CNERF_REQUIRE(n_points >= 0, "cnerf_mlp_bwd_data: negative n_points");
    if (n_points == 0) return CNERF_OK;
    BwdCtx c;
    int rc = bwd_ctx(acts, grads_rec, n_points, workspace, stream, &c);
    if (rc != CNERF_OK) return rc;
    cudaError_t e = cudaMemsetAsync(c.amax, 0, 4, c.st);
    if (e != cudaSuccess) return check_cuda(e, "cudaMemsetAsync(amax)");
    absmax_kernel<<<kNumSMs, 256, 0, c.st>>>(d_raw, (int64_t)n_points * 4, c.amax);
    CNERF_LAUNCH_CHECK("absmax_kernel");
#define CNERF_CHAIN(P, T, L) mlp_bwd_data3_kernel<P, T, L><<<c.grid, kC3Threads, kC3Smem, c.st>>>(w->stream_bwd3, w->misc, d_raw, c.a, c.amax, n_points, c.g)
    // fp16 chain: the two-tile kernel; CNERF_CHAIN1=single (or the phase profile) keeps the single-tile one-term instantiation
    static const bool single1 = [] { const char* ev = getenv("CNERF_CHAIN1"); return ev && ev[0] == 's'; }();
    if (chain_terms == 1 && !g_profc_host && !single1)
        mlp_bwd_data5_kernel<<<c.grid, kC5Threads, kC5Smem, c.st>>>(w->stream_bwd3, w->misc, d_raw, c.a, c.amax, n_points, c.g);
    else if (chain_terms == 1) { if (g_profc_host) CNERF_CHAIN(true, 1, false); else CNERF_CHAIN(false, 1, false); }
    else if (dw_terms == 1) CNERF_CHAIN(false, 3, false);
    else if (g_profc_host) CNERF_CHAIN(true, 3, true);
    else CNERF_CHAIN(false, 3, true);
#undef CNERF_CHAIN
    CNERF_LAUNCH_CHECK("mlp_bwd_data3_kernel");
    return CNERF_OK;
}

// Stage 3: the two narrow heads (needs only d_raw and the activation record).
extern "C" int cnerf_mlp_bwd_heads(const float* d_raw, const void* acts, int n_points, float* d_alpha_w, float* d_alpha_b,
                                   float* d_rgb_w, float* d_rgb_b, int accumulate, int dw_terms, void* workspace, void* stream) {
    if (int rc0 = check_terms("cnerf_mlp_bwd_heads", 3, dw_terms)) return rc0;
    CNERF_REQUIRE(d_raw && acts && workspace && d_alpha_w && d_alpha_b && d_rgb_w && d_rgb_b, "cnerf_mlp_bwd_heads: null pointer");
    CNERF_REQUIRE(n_points >= 0, "cnerf_mlp_bwd_heads: negative n_points");
    if (n_points == 0) return CNERF_OK;
    BwdCtx c;
    int rc = bwd_ctx(acts, nullptr, n_points, workspace, stream, &c);
    if (rc != CNERF_OK) return rc;
    const int head_grid = c.tiles < kHeadCtas ? c.tiles : kHeadCtas;
    if (dw_terms == 3) mlp_heads_grad_kernel<true><<<head_grid, 256, 0, c.st>>>(d_raw, c.a, n_points, c.head_part);
    else mlp_heads_grad_kernel<false><<<head_grid, 256, 0, c.st>>>(d_raw, c.a, n_points, c.head_part);
    CNERF_LAUNCH_CHECK("mlp_heads_grad_kernel");
    heads_reduce_kernel<<<ceil_div(kHeadFloats, 256), 256, 0, c.st>>>(c.head_part, head_grid, d_rgb_w, d_rgb_b, d_alpha_w, d_alpha_b, accumulate);
    CNERF_LAUNCH_CHECK("heads_reduce_kernel");
    return CNERF_OK;
}

// Stage 2: weight / bias gradients of the ten GEMM layers from grads_rec (after cnerf_mlp_bwd_data on the same workspace).
extern "C" int cnerf_mlp_bwd_weights(const void* acts, const void* grads_rec, int n_points, float* const* d_pts_w,
                                     float* const* d_pts_b, float* d_feature_w, float* d_feature_b, float* d_views_w,
                                     float* d_views_b, int accumulate, int dw_terms, void* workspace, void* stream) {
    CNERF_REQUIRE(acts && grads_rec && workspace && d_pts_w && d_pts_b && d_feature_w && d_feature_b && d_views_w && d_views_b,
                  "cnerf_mlp_bwd_weights: null pointer");
    if (int rc0 = check_terms("cnerf_mlp_bwd_weights", 3, dw_terms)) return rc0;
    CNERF_REQUIRE(n_points >= 0, "cnerf_mlp_bwd_weights: negative n_points");
    for (int i = 0; i < 8; ++i) CNERF_REQUIRE(d_pts_w[i] && d_pts_b[i], "cnerf_mlp_bwd_weights: null pts_linears.%d gradient", i);
    if (n_points == 0) return CNERF_OK;
    BwdCtx c;
    int rc = bwd_ctx(acts, const_cast<void*>(grads_rec), n_points, workspace, stream, &c);
    if (rc != CNERF_OK) return rc;
    cudaStream_t st = c.st;
    const int grid = c.grid, tiles = c.tiles;
    auto H = [](int l) { return (uint32_t)(kSlotH0 + (size_t)l * 131072); };
    const uint64_t a_rows = (uint64_t)tiles * (kTileBytes / 2048), g_rows = (uint64_t)tiles * (kGTileBytes / 2048);
    auto box_rows = [dw_terms](const DwSrc& s) { return (dw_terms == 3 && s.lo_off == s.kgroups * 2048) ? 2 * s.kgroups : s.kgroups; };
    DwSegs all = {};
    DbSegs alldb = {};
    static thread_local DwAll A;      // ~4.6 KB of kernel parameters: tensor maps + pass descriptors of all ten passes
    int pass = 0;
    auto run_pass = [&](DwPass& P, const DwSeg* segs, int nseg, float* db0, int n0, float* db1, int n1) -> int {
        P.terms = dw_terms;
        const bool wide = dw_terms == 1;
        int r = make_record_map(&A.map[pass][0], c.g, g_rows, box_rows(P.a[0]), wide);
        if (r == CNERF_OK) r = make_record_map(&A.map[pass][1], c.a, a_rows, box_rows(P.x[0]), wide);
        if (r == CNERF_OK) r = make_record_map(&A.map[pass][2], c.a, a_rows, box_rows(P.x[P.n_x - 1]), wide);
        if (r != CNERF_OK) return r;
        A.P[pass] = P;
        for (int i = 0; i < nseg; ++i) { all.s[all.n] = segs[i]; all.s[all.n].pass = pass; ++all.n; }
        if (db0) alldb.s[alldb.n++] = {db0, n0, pass, 0};
        if (db1) alldb.s[alldb.n++] = {db1, n1, pass, 1};
        ++pass;
        return CNERF_OK;
    };
    {   // encoding pass: dW0 = G0^T E, dW5[:, :63] = G5^T E
        DwPass P = {}; P.n_a = 2; P.n_x = 1; P.db_mask = 3;
        P.a[0] = {(uint32_t)g_slot(0), 32, 65536}; P.a[1] = {(uint32_t)g_slot(5), 32, 65536}; P.x[0] = {(uint32_t)kSlotE, 8, 16384};
        DwSeg S[4] = {{d_pts_w[0], 63, 0, 0, 63, 0 * 128 * 64, 64, 0, 0}, {d_pts_w[0], 63, 128, 0, 63, 1 * 128 * 64, 64, 0, 0},
                      {d_pts_w[5], 319, 0, 0, 63, 2 * 128 * 64, 64, 0, 0}, {d_pts_w[5], 319, 128, 0, 63, 3 * 128 * 64, 64, 0, 0}};
        if ((rc = run_pass(P, S, 4, d_pts_b[0], 256, d_pts_b[5], 256)) != CNERF_OK) return rc;
    }
    for (int l = 1; l <= 8; ++l) {   // 256-wide hidden inputs: layers 1..7 (layer 5: the h4 columns) and feature_linear (l == 8)
        DwPass P = {}; P.n_a = 1; P.n_x = 1; P.db_mask = (l == 5) ? 0 : 1;
        P.a[0] = {(uint32_t)g_slot(l), 32, 65536}; P.x[0] = {H(l - 1), 32, 65536};
        float* dst = l == 8 ? d_feature_w : d_pts_w[l];
        const int ld = l == 5 ? 319 : 256, c0 = l == 5 ? 63 : 0;
        DwSeg S[2] = {{dst, ld, 0, c0, 256, 0, 256, 0, 0}, {dst, ld, 128, c0, 256, 128 * 256, 256, 0, 0}};
        float* db = l == 5 ? nullptr : (l == 8 ? d_feature_b : d_pts_b[l]);
        if ((rc = run_pass(P, S, 2, db, 256, nullptr, 0)) != CNERF_OK) return rc;
    }
    {   // views_linears.0: X = [feature (256), direction encoding (27 of 32)]
        DwPass P = {}; P.n_a = 1; P.n_x = 2; P.db_mask = 1;
        P.a[0] = {(uint32_t)g_slot(9), 16, 32768}; P.x[0] = {(uint32_t)kSlotF, 32, 65536}; P.x[1] = {(uint32_t)kSlotV, 4, 16384};
        DwSeg S[2] = {{d_views_w, 283, 0, 0, 256, 0, 256, 0, 0}, {d_views_w, 283, 0, 256, 27, 128 * 256, 32, 0, 0}};
        if ((rc = run_pass(P, S, 2, d_views_b, 128, nullptr, 0)) != CNERF_OK) return rc;
    }
    A.n_pass = pass;
    mlp_bwd_weight_kernel<<<grid, kDwThreads, kDwSmem, st>>>(A, tiles, c.dw_part, c.db_part);
    CNERF_LAUNCH_CHECK("mlp_bwd_weight_kernel");
    dw_reduce_kernel<<<dim3(ceil_div(128 * 256, 256), all.n), 256, 0, st>>>(all, c.dw_part, grid, c.amax, accumulate);
    CNERF_LAUNCH_CHECK("dw_reduce_kernel");
    db_reduce_kernel<<<dim3(1, alldb.n), 256, 0, st>>>(alldb, c.db_part, grid, c.amax, accumulate);
    CNERF_LAUNCH_CHECK("db_reduce_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_mlp_bwd(const cnerf_weights* w, const float* d_raw, const void* acts, void* grads_rec, int n_points,
                             float* const* d_pts_w, float* const* d_pts_b, float* d_feature_w, float* d_feature_b,
                             float* d_alpha_w, float* d_alpha_b, float* d_views_w, float* d_views_b, float* d_rgb_w,
                             float* d_rgb_b, int accumulate, int chain_terms, int dw_terms, void* workspace, void* stream) {
    int rc = cnerf_mlp_bwd_data(w, d_raw, acts, grads_rec, n_points, chain_terms, dw_terms, workspace, stream);
    if (rc == CNERF_OK) rc = cnerf_mlp_bwd_heads(d_raw, acts, n_points, d_alpha_w, d_alpha_b, d_rgb_w, d_rgb_b, accumulate, dw_terms, workspace, stream);
    if (rc == CNERF_OK) rc = cnerf_mlp_bwd_weights(acts, grads_rec, n_points, d_pts_w, d_pts_b, d_feature_w, d_feature_b, d_views_w,
                                                   d_views_b, accumulate, dw_terms, workspace, stream);
    return rc;
}
