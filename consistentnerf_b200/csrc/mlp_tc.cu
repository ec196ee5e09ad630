// K2+K3 fused: positional encoding + the 8x256 NeRF MLP on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM walks 128-point tiles.  Per tile the ten GEMM layers
// (pts_linears.0-7, feature_linear, views_linears.0; alpha_linear and rgb_linear are folded
// into the epilogues as fp32 dot products) run back to back without the activations ever
// leaving the SM:
//
//   warp 8  (1 lane)  streams pre-packed 16 KB weight blocks HBM/L2 -> SMEM ring with
//                     cp.async.bulk (TMA bulk copy, mbarrier complete_tx);
//   warp 9  (1 lane)  issues tcgen05.mma (M=128, N=128, K=16, fp16 x fp16 -> fp32 in TMEM) and
//                     owns the TMEM allocation;
//   warps 0-7         prologue (points -> sin/cos encoding -> SMEM A operand) and the epilogues
//                     (TMEM -> registers, +bias, ReLU, fp16 hi/lo split, -> SMEM A operand of the
//                     next layer).
//
// Precision ("fp16x3"): every fp32 operand x is carried as hi = fp16(x), lo = fp16(x - hi) and a
// product is evaluated as a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation, i.e. the
// dropped a_lo*w_lo term is ~2^-22 relative.  That is what keeps the renderer inside the 1e-4
// parity bar of the fp32 reference (hard part 1) at the price of 3 MMAs per MAC.
//
// SMEM operand layout (both A and B): K-major, no swizzle ("interleaved"): 8x8 fp16 core matrices of
// 128 contiguous bytes; element (row r, k) lives at  (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2  with
// SBO = 128 B and LBO = rows*16 B = 2048 B for 128 rows.  A thread that owns a row writes 8
// consecutive k as one 16-byte store, and a warp's 32 rows are 512 contiguous bytes: conflict free.
#include "mlp_layout.cuh"
#include <stdlib.h>
#include <vector>

namespace cnerf {

// ------------------------------------------------------------------------------------
// The weight-block program: the order in which 16 KB blocks (128 out-rows x 32 k, hi+lo) are
// streamed and multiplied.  Built once on the host, mirrored in constant memory.
// ------------------------------------------------------------------------------------
struct BlkInfo {
    uint8_t layer, half, a_src, a_kg;           // a_src: 0 = encoding buffer, 1 = activation buffer
    uint8_t first, last_of_layer, wait_a, kvalid;
    uint16_t src_k0, pad;
};
__constant__ BlkInfo c_blocks[kMaxBlocks];
__constant__ int c_num_blocks;

static const int kLayerLd[kNumLayers] = {63, 256, 256, 256, 256, 319, 256, 256, 256, 283};

static std::vector<BlkInfo> build_program() {
    std::vector<BlkInfo> prog;
    for (int layer = 0; layer < kNumLayers; ++layer) {
        struct KB { uint8_t src, kg; uint16_t k0; uint8_t kvalid; };
        std::vector<KB> kbs;
        auto act_blocks = [&](int col0) { for (int j = 0; j < 8; ++j) kbs.push_back({1, (uint8_t)(4 * j), (uint16_t)(col0 + 32 * j), 32}); };
        if (layer == 0) { kbs.push_back({0, 0, 0, 32}); kbs.push_back({0, 4, 32, 31}); }
        else if (layer == 5) { kbs.push_back({0, 0, 0, 32}); kbs.push_back({0, 4, 32, 31}); act_blocks(63); }   // cat([pts, h])
        else if (layer == 9) { act_blocks(0); kbs.push_back({0, 0, 256, 27}); }                                // cat([feature, dirs])
        else act_blocks(0);
        int halves = layer == 9 ? 1 : 2;
        for (int h = 0; h < halves; ++h)
            for (size_t j = 0; j < kbs.size(); ++j) {
                BlkInfo b = {};
                b.layer = (uint8_t)layer; b.half = (uint8_t)h; b.a_src = kbs[j].src; b.a_kg = kbs[j].kg;
                b.first = j == 0; b.last_of_layer = (h == halves - 1) && (j + 1 == kbs.size());
                b.wait_a = (h == 0 && j == 0); b.kvalid = kbs[j].kvalid; b.src_k0 = kbs[j].k0;
                prog.push_back(b);
            }
    }
    return prog;
}

__global__ void __launch_bounds__(256)
pack_weights_kernel(RawParams p, uint8_t* __restrict__ stream) {
    BlkInfo bi = c_blocks[blockIdx.x];
    const float* W = p.w[bi.layer];
    int ld = p.ld[bi.layer];
    uint8_t* dst = stream + (size_t)blockIdx.x * kBlockBytes;
    for (int u = threadIdx.x; u < 512; u += 256) {
        int n = u & 127, kg = u >> 7;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int k = kg * 8 + e;
            v[e] = (k < bi.kvalid) ? W[(size_t)(bi.half * 128 + n) * ld + bi.src_k0 + k] : 0.f;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pack2(v[2 * e], v[2 * e + 1], h[e], l[e]);
        size_t off = (size_t)kg * kLBO + (size_t)n * 16;
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + kBlockHalfBytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

__global__ void pack_misc_kernel(RawParams p, float* __restrict__ misc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kMiscFloats) return;
    float v = 0.f;
    if (i < kMiscAlphaW) { int l = i >> 8, c = i & 255; v = (l == 9 && c >= 128) ? 0.f : p.b[l][c]; }
    else if (i < kMiscAlphaB) v = p.alpha_w[i - kMiscAlphaW];
    else if (i == kMiscAlphaB) v = p.alpha_b[0];
    else if (i >= kMiscRgbW && i < kMiscRgbB) v = p.rgb_w[i - kMiscRgbW];
    else if (i >= kMiscRgbB && i < kMiscRgbB + 3) v = p.rgb_b[i - kMiscRgbB];
    misc[i] = v;
}

// column j of the 63-wide point encoding / 27-wide direction encoding of (x0,x1,x2)
template <int J>
__device__ __forceinline__ float enc_col(const float (&x)[3], int width) {
    if (J >= width) return 0.f;
    if (J < 3) return x[J];
    constexpr int b = (J - 3) / 3, c = (J - 3) % 3, oct = b / 2;
    float arg = x[c] * (float)(1 << oct);
    return (b & 1) ? cosf(arg) : sinf(arg);
}
template <int J0>
__device__ __forceinline__ void enc_group8(const float (&x)[3], int width, float* v) {
    v[0] = enc_col<J0 + 0>(x, width); v[1] = enc_col<J0 + 1>(x, width); v[2] = enc_col<J0 + 2>(x, width);
    v[3] = enc_col<J0 + 3>(x, width); v[4] = enc_col<J0 + 4>(x, width); v[5] = enc_col<J0 + 5>(x, width);
    v[6] = enc_col<J0 + 6>(x, width); v[7] = enc_col<J0 + 7>(x, width);
}
// 32 encoding columns [32*part, 32*part+32) of row `row` -> k-groups 4*part .. 4*part+3
__device__ __forceinline__ void write_encoding(uint32_t hi_base, uint32_t lo_base, uint32_t row, int part,
                                               const float (&x)[3], int width) {
    float v[8];
    if (part == 0) {
        enc_group8<0>(x, width, v);  store_split8(hi_base, lo_base, row, 0, v);
        enc_group8<8>(x, width, v);  store_split8(hi_base, lo_base, row, 1, v);
        enc_group8<16>(x, width, v); store_split8(hi_base, lo_base, row, 2, v);
        enc_group8<24>(x, width, v); store_split8(hi_base, lo_base, row, 3, v);
    } else {
        enc_group8<32>(x, width, v); store_split8(hi_base, lo_base, row, 4, v);
        enc_group8<40>(x, width, v); store_split8(hi_base, lo_base, row, 5, v);
        enc_group8<48>(x, width, v); store_split8(hi_base, lo_base, row, 6, v);
        enc_group8<56>(x, width, v); store_split8(hi_base, lo_base, row, 7, v);
    }
}

template <bool kSave>
__global__ void __launch_bounds__(kThreads, 1)
mlp_fused_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc, const float* __restrict__ pts,
                 const float* __restrict__ viewdirs, int n_points, int n_samples, int n_rays, float* __restrict__ raw,
                 uint8_t* __restrict__ acts) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + kBars, bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_a = bar_empty + 8 * kStages, bar_d = bar_a + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kTmemSlot);
    float* s_alpha = reinterpret_cast<float*>(smem + kSAlpha);
    float* s_rgb = reinterpret_cast<float*>(smem + kSRgb);
    const int num_tiles = (n_points + (int)kRows - 1) / (int)kRows;
    const int nblk = c_num_blocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_a, kEpiThreads);
        mbar_init(bar_d, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(sbase + kTmemSlot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ===== weight loader =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int b = 0; b < nblk; ++b, ++it) {
                    uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(bar_full + 8 * s, kBlockBytes);
                    bulk_g2s(sbase + kRing + s * kBlockBytes, wstream + (size_t)b * kBlockBytes, kBlockBytes, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(128, 128);
            uint32_t it = 0, a_cnt = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int b = 0; b < nblk; ++b, ++it) {
                    BlkInfo bi = c_blocks[b];
                    if (bi.wait_a) {
                        mbar_wait(bar_a, a_cnt & 1); ++a_cnt; tc_fence_after();
                        if (kSave) {      // the A operand of this layer is complete in SMEM: stream it out
                            uint8_t* rec = acts + (size_t)tile * kTileBytes;
                            const int L = bi.layer;
                            if (L == 0) bulk_s2g(rec + kSlotE, sbase + kEmbHi, 32768);
                            else if (L <= 8) bulk_s2g(rec + kSlotH0 + (size_t)(L - 1) * 131072, sbase + kActHi, 131072);
                            else bulk_s2g(rec + kSlotF, sbase + kActHi, 131072);
                            if (L == 6) bulk_s2g(rec + kSlotV, sbase + kEmbHi, 32768);
                            bulk_commit();
                        }
                    }
                    uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    uint32_t a_hi = sbase + (bi.a_src ? kActHi : kEmbHi) + bi.a_kg * kLBO;
                    uint32_t a_lo = sbase + (bi.a_src ? kActLo : kEmbLo) + bi.a_kg * kLBO;
                    uint32_t b_hi = sbase + kRing + s * kBlockBytes, b_lo = b_hi + kBlockHalfBytes;
                    uint32_t d = tmem + bi.half * 128;
#pragma unroll
                    for (uint32_t ks = 0; ks < 2; ++ks) {
                        uint64_t ah = smem_desc(a_hi + ks * 2 * kLBO), al = smem_desc(a_lo + ks * 2 * kLBO);
                        uint64_t bh = smem_desc(b_hi + ks * 2 * kLBO), bl = smem_desc(b_lo + ks * 2 * kLBO);
                        umma_f16(d, ah, bh, idesc, (bi.first && ks == 0) ? 0u : 1u);
                        umma_f16(d, ah, bl, idesc, 1u);
                        umma_f16(d, al, bh, idesc, 1u);
                    }
                    umma_commit(bar_empty + 8 * s);            // frees the ring slot when these MMAs retire
                    if (bi.last_of_layer) {
                        if (kSave) bulk_wait_read0();           // the epilogue may now overwrite the stored buffers
                        umma_commit(bar_d);                     // accumulator of the layer is complete
                    }
                }
                if (kSave) {              // views_linears output (post-ReLU) written by the last epilogue
                    mbar_wait(bar_a, a_cnt & 1); ++a_cnt;
                    uint8_t* rec = acts + (size_t)tile * kTileBytes;
                    bulk_s2g(rec + kSlotHV, sbase + kActHi, 32768);
                    bulk_s2g(rec + kSlotHV + 32768, sbase + kActLo, 32768);
                    bulk_commit();
                    bulk_wait_read0();    // the next tile's first epilogue overwrites the buffer only after its MMAs anyway
                }
            }
            if (kSave) bulk_wait0();
        }
    } else {
        // ===== prologue + epilogue warps =====
        const int e = warp;                                  // 0..7
        const uint32_t row = (uint32_t)((e & 3) * 32 + lane);
        const int part = e >> 2;                             // which half of the columns this thread owns
        const uint32_t t_lane = tmem + ((uint32_t)((e & 3) * 32) << 16);
        uint32_t d_cnt = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int grow = tile * (int)kRows + (int)row;
            const bool valid = grow < n_points;
            {   // point encoding -> A operand of layer 0 (and of the skip at layer 5)
                float x[3] = {0.f, 0.f, 0.f};
                if (valid) { x[0] = pts[3 * (size_t)grow]; x[1] = pts[3 * (size_t)grow + 1]; x[2] = pts[3 * (size_t)grow + 2]; }
                write_encoding(sbase + kEmbHi, sbase + kEmbLo, row, part, x, 63);
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(bar_a);
            }
            float alpha_acc = 0.f;
            for (int layer = 0; layer < kNumLayers; ++layer) {
                mbar_wait(bar_d, d_cnt & 1); ++d_cnt;
                tc_fence_after();
                const float* bias = misc + kMiscBias + layer * 256;
                if (layer < 9) {
                    const bool relu = layer != 8;
                    const uint32_t col0 = (uint32_t)part * 128;
#pragma unroll 1
                    for (uint32_t ch = 0; ch < 4; ++ch) {
                        float v[32];
                        tmem_ld32(t_lane + col0 + ch * 32, v);
                        tmem_ld_wait();
                        const uint32_t c = col0 + ch * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c + j));
                            v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float t = relu ? fmaxf(v[j], 0.f) : fmaxf(v[j], -65504.f);
                            v[j] = fminf(t, 65504.f);                     // keep the fp16 split finite
                        }
                        if (layer == 7) {                                 // alpha_linear on the fp32 activations
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 aw = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c + j));
                                alpha_acc = fmaf(v[j], aw.x, alpha_acc); alpha_acc = fmaf(v[j + 1], aw.y, alpha_acc);
                                alpha_acc = fmaf(v[j + 2], aw.z, alpha_acc); alpha_acc = fmaf(v[j + 3], aw.w, alpha_acc);
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) store_split8(sbase + kActHi, sbase + kActLo, row, (c >> 3) + g, v + 8 * g);
                    }
                    if (layer == 5) {
                        // layer 5 consumed the point encoding: reuse its buffer for the direction encoding
                        float dvec[3] = {0.f, 0.f, 0.f};
                        if (valid) {
                            int ray = min(grow / n_samples, n_rays - 1);
                            dvec[0] = viewdirs[3 * (size_t)ray]; dvec[1] = viewdirs[3 * (size_t)ray + 1]; dvec[2] = viewdirs[3 * (size_t)ray + 2];
                        }
                        float v[8];
                        if (part == 0) {
                            enc_group8<0>(dvec, 27, v); store_split8(sbase + kEmbHi, sbase + kEmbLo, row, 0, v);
                            enc_group8<8>(dvec, 27, v); store_split8(sbase + kEmbHi, sbase + kEmbLo, row, 1, v);
                        } else {
                            enc_group8<16>(dvec, 27, v); store_split8(sbase + kEmbHi, sbase + kEmbLo, row, 2, v);
                            enc_group8<24>(dvec, 27, v); store_split8(sbase + kEmbHi, sbase + kEmbLo, row, 3, v);
                        }
                    }
                    if (layer == 7) {
                        if (part == 1) s_alpha[row] = alpha_acc;
                        named_bar_sync(1, kEpiThreads);
                        if (part == 0) alpha_acc += s_alpha[row] + __ldg(misc + kMiscAlphaB);
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(bar_a);
                } else {
                    // views layer (N = 128): ReLU, then rgb_linear as an fp32 dot product
                    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
                    const uint32_t col0 = (uint32_t)part * 64;
#pragma unroll 1
                    for (uint32_t ch = 0; ch < 2; ++ch) {
                        float v[32];
                        tmem_ld32(t_lane + col0 + ch * 32, v);
                        tmem_ld_wait();
                        const uint32_t c = col0 + ch * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float h = fmaxf(v[j] + __ldg(bias + c + j), 0.f);
                            r0 = fmaf(h, __ldg(misc + kMiscRgbW + c + j), r0);
                            r1 = fmaf(h, __ldg(misc + kMiscRgbW + 128 + c + j), r1);
                            r2 = fmaf(h, __ldg(misc + kMiscRgbW + 256 + c + j), r2);
                            if (kSave) v[j] = fminf(h, 65504.f);
                        }
                        if (kSave) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) store_split8(sbase + kActHi, sbase + kActLo, row, (c >> 3) + g, v + 8 * g);
                        }
                    }
                    if (part == 1) { s_rgb[row] = r0; s_rgb[128 + row] = r1; s_rgb[256 + row] = r2; }
                    named_bar_sync(1, kEpiThreads);
                    if (part == 0 && valid) {
                        float4 o;
                        o.x = r0 + s_rgb[row] + __ldg(misc + kMiscRgbB);
                        o.y = r1 + s_rgb[128 + row] + __ldg(misc + kMiscRgbB + 1);
                        o.z = r2 + s_rgb[256 + row] + __ldg(misc + kMiscRgbB + 2);
                        o.w = alpha_acc;
                        *reinterpret_cast<float4*>(raw + 4 * (size_t)grow) = o;
                    }
                    // TMEM reads are ordered before the next tile's prologue arrive (tc_fence_before there)
                    if (kSave) { fence_proxy_async(); tc_fence_before(); mbar_arrive(bar_a); }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem, kTmemCols);
}


// ------------------------------------------------------------------------------------
// unit self-test of descriptors / layout / TMEM mapping:  d[128,n] = a[128,k] b[n,k]^T
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int k, float* __restrict__ d) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int kgs = k / 8;
    const uint32_t a_hi = sbase, a_lo = a_hi + kgs * 2048;
    const uint32_t lbo_b = (uint32_t)n * 16;
    const uint32_t b_hi = a_lo + kgs * 2048, b_lo = b_hi + kgs * lbo_b;
    const uint32_t bar = b_lo + kgs * lbo_b;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + (bar - sbase) + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(bar + 8, 256);
    // operands: thread = row for A; rows of B strided over threads
    for (int kg = 0; kg < kgs; ++kg) {
        float v[8];
        for (int e = 0; e < 8; ++e) v[e] = a[(size_t)threadIdx.x * k + kg * 8 + e];
        store_split8(a_hi, a_lo, threadIdx.x, kg, v);
    }
    for (int r = threadIdx.x; r < n; r += 128)
        for (int kg = 0; kg < kgs; ++kg) {
            float v[8];
            for (int e = 0; e < 8; ++e) v[e] = b[(size_t)r * k + kg * 8 + e];
            uint32_t h[4], l[4];
            for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
            uint32_t off = kg * lbo_b + r * 16;
            st_shared_v4(b_hi + off, h[0], h[1], h[2], h[3]);
            st_shared_v4(b_lo + off, l[0], l[1], l[2], l[3]);
        }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        uint32_t idesc = instr_desc(128, (uint32_t)n);
        auto desc_b = [&](uint32_t addr) {
            return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo_b >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
        };
        for (int ks = 0; ks < k / 16; ++ks) {
            uint64_t ah = smem_desc(a_hi + ks * 2 * kLBO), al = smem_desc(a_lo + ks * 2 * kLBO);
            uint64_t bh = desc_b(b_hi + ks * 2 * lbo_b), bl = desc_b(b_lo + ks * 2 * lbo_b);
            umma_f16(tmem, ah, bh, idesc, ks == 0 ? 0u : 1u);
            umma_f16(tmem, ah, bl, idesc, 1u);
            umma_f16(tmem, al, bh, idesc, 1u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < n; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32 && c + j < n; ++j) d[(size_t)threadIdx.x * n + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// Same product with the A operand in tensor memory (tcgen05.st by the row-owning threads, TS-mode MMA): the
// building block of the kernels that keep activations in TMEM between layers.
__global__ void __launch_bounds__(128, 1)
umma_selftest_ts_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int k, float* __restrict__ d) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int kgs = k / 8;
    const uint32_t lbo_b = (uint32_t)n * 16;
    const uint32_t b_hi = sbase, b_lo = b_hi + kgs * lbo_b;
    const uint32_t bar = b_lo + kgs * lbo_b;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + (bar - sbase) + 8);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(bar + 8, 512);
    for (int r = threadIdx.x; r < n; r += 128)
        for (int kg = 0; kg < kgs; ++kg) {
            float v[8];
            for (int e = 0; e < 8; ++e) v[e] = b[(size_t)r * k + kg * 8 + e];
            uint32_t h[4], l[4];
            for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
            uint32_t off = kg * lbo_b + r * 16;
            st_shared_v4(b_hi + off, h[0], h[1], h[2], h[3]);
            st_shared_v4(b_lo + off, l[0], l[1], l[2], l[3]);
        }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    // A: row = lane, 8 fp16 of K per 4 columns; hi at columns 256.., lo at columns 384..
    for (int kg = 0; kg < kgs; kg += 2) {
        uint32_t h[8], l[8];
        for (int i = 0; i < 8; ++i) {
            float x0 = a[(size_t)threadIdx.x * k + kg * 8 + 2 * i], x1 = a[(size_t)threadIdx.x * k + kg * 8 + 2 * i + 1];
            split_pack2(x0, x1, h[i], l[i]);
        }
        tmem_st8(t_lane + 256 + kg * 4, h);
        tmem_st8(t_lane + 384 + kg * 4, l);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        uint32_t idesc = instr_desc(128, (uint32_t)n);
        auto desc_b = [&](uint32_t addr) {
            return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo_b >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
        };
        for (int ks = 0; ks < k / 16; ++ks) {
            uint64_t bh = desc_b(b_hi + ks * 2 * lbo_b), bl = desc_b(b_lo + ks * 2 * lbo_b);
            umma_f16_ts(tmem, tmem + 256 + ks * 8, bh, idesc, ks == 0 ? 0u : 1u);
            umma_f16_ts(tmem, tmem + 256 + ks * 8, bl, idesc, 1u);
            umma_f16_ts(tmem, tmem + 384 + ks * 8, bh, idesc, 1u);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < n; c += 32) {
        float v[32];
        tmem_ld32(t_lane + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32 && c + j < n; ++j) d[(size_t)threadIdx.x * n + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// Issue-rate microbenchmark: `iters` back-to-back MMAs (M=128, N=n, K=16) on one CTA per SM.
//   mode 0: SS (A and B in shared memory, K-major no swizzle)   mode 1: TS (A in tensor memory)
//   alt != 0: alternate between two accumulators (columns 0 and 256) instead of chaining on one
__global__ void __launch_bounds__(128, 1)
umma_bench_kernel(int mode, int n, int iters, int alt, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_s = sbase, b_s = sbase + 8192, bar = sbase + 8192 + 16384 * 2;
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 8192 + 32768 + 8);   // bar, slot, bar2 (+16), bar3 (+24)
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (8192 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(bar + 8, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *slot;
    if (warp == 0 || (warp == 1 && (alt & 128))) {
        // alt & 128: two issuing warps, each with its own accumulator columns (and its own half of a 2n-row B tile)
        const uint32_t tmem = tmem0 + (uint32_t)warp * 256;
        const uint32_t idesc = instr_desc(128, (uint32_t)n);
        const uint32_t lbo_b = (uint32_t)n * 16 * ((alt & 128) ? 2 : 1);
        const uint32_t bar = sbase + 8192 + 16384 * 2 + 32 * (uint32_t)warp;
        uint64_t bd[4], ad[4];
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
            bd[ks] = (uint64_t)(((b_s + ((alt & 128) ? ks * 2048 + (uint32_t)warp * n * 16 : ks * 2 * lbo_b)) >> 4) & 0x3FFF) | ((uint64_t)(lbo_b >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
            ad[ks] = smem_desc(a_s + ks * 2 * kLBO);
        }
        const uint32_t d1 = (alt && n <= 128) ? tmem + 128 : tmem;
        // loop-structure variants (bits of `alt` above 1): 2 = mbarrier try_wait on a completed barrier per group of 4,
        // 4 = tcgen05.fence::after_thread_sync per group, 8 = tcgen05.commit per group, 16 = second commit per group
        const uint32_t bar2 = bar + 16, bar3 = bar + 24;
        if ((threadIdx.x & 31) == 0) { if (warp == 1) mbar_init(bar, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); fence_barrier_init(); mbar_arrive(bar2); }
        __syncwarp();
        const bool ts = mode != 0;
        long long t0 = clock64();
        uint32_t chain = threadIdx.x;
        for (int i = 0; i < iters; i += 4) {
            if (alt & 2) mbar_wait(bar2, 0);
            if (alt & 4) tc_fence_after();
            if (alt & 32) {                    // ~60 dependent integer multiply-adds (a few hundred cycles of pure ALU latency)
#pragma unroll
                for (int r = 0; r < 60; ++r) chain = chain * 1664525u + 1013904223u + (uint32_t)i;
            }
            if (alt & 64) {                    // 8 loads of a shared-memory word that is not an mbarrier
#pragma unroll
                for (int r = 0; r < 8; ++r) chain += *reinterpret_cast<volatile uint32_t*>(smem + 8192 + 32768 + 40 + 4 * (chain & 1));
            }
            if (elect_one()) {
                if (!ts) {
                    umma_f16(tmem, ad[0], bd[0], idesc, 1u); umma_f16(d1, ad[1], bd[1], idesc, 1u);
                    umma_f16(tmem, ad[2], bd[2], idesc, 1u); umma_f16(d1, ad[3], bd[3], idesc, 1u);
                } else {
                    umma_f16_ts(tmem, tmem + 384, bd[0], idesc, 1u); umma_f16_ts(d1, tmem + 392, bd[1], idesc, 1u);
                    umma_f16_ts(tmem, tmem + 400, bd[2], idesc, 1u); umma_f16_ts(d1, tmem + 408, bd[3], idesc, 1u);
                }
                if (alt & 8) umma_commit(bar3);
                if (alt & 16) umma_commit(bar3);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 0);
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && warp == 0) out[blockIdx.x] = (float)(t1 - t0) / (float)iters + (chain == 0x12345u ? 1.f : 0.f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, 512);
}

}  // namespace cnerf

using namespace cnerf;

namespace cnerf {
int upload_bwd_program_once(int* nblocks);                                   // mlp_bwd_tc.cu
int pack_bwd_stream(const RawParams& p, uint8_t* stream_bwd, int nblocks, cudaStream_t st);
int bwd_stream3_blocks();
int pack_bwd_stream3(const RawParams& p, uint8_t* stream, cudaStream_t st);
int pack_stream3(const RawParams& p, uint8_t* stream3, cudaStream_t st);                                // mlp_fwd3.cu
int stream3_blocks();
int pack_stream4(const RawParams& p, uint8_t* stream4, cudaStream_t st);                                // mlp_fwd4.cu
size_t stream4_bytes();
int launch_fused4(const uint8_t* stream4, const float* misc, const float* pts, const float* viewdirs, int n_points,
                  int n_samples, int n_rays, float* raw, cudaStream_t st);
int launch_fused3(const uint8_t* stream3, const float* misc, const float* pts, const float* viewdirs, int n_points,
                  int n_samples, int n_rays, float* raw, uint8_t* acts, cudaStream_t st);
}

// Kernel generation selected once per process.  Forward (CNERF_MLP_IMPL): default 3 = single-CTA N=256 kernel (mlp_fwd3.cu);
// kept for A/B runs: 4 = CTA-pair ping-pong kernel (mlp_fwd4.cu, correct but slower), 1 = first-generation serial
// kernel.  Backward data chain (CNERF_BWD_IMPL): default 3, 1 = first generation.  Only the streams in use are packed.
static int fwd_impl() {
    static int impl = 0;
    if (!impl) { const char* ev = getenv("CNERF_MLP_IMPL"); impl = (ev && (ev[0] == '1' || ev[0] == '4')) ? ev[0] - '0' : 3; }
    return impl;
}
static int bwd_impl() {
    static int impl = 0;
    if (!impl) { const char* ev = getenv("CNERF_BWD_IMPL"); impl = ((ev && ev[0] == '1') || fwd_impl() == 1) ? 1 : 3; }
    return impl;
}

static int upload_program(int* nblocks) {
    std::vector<BlkInfo> prog = build_program();
    if ((int)prog.size() > kMaxBlocks) return set_error(CNERF_EINVAL, "program too long");
    int n = (int)prog.size();
    cudaError_t e = cudaMemcpyToSymbol(c_blocks, prog.data(), prog.size() * sizeof(BlkInfo));
    if (e != cudaSuccess) return check_cuda(e, "cudaMemcpyToSymbol(c_blocks)");
    e = cudaMemcpyToSymbol(c_num_blocks, &n, sizeof(int));
    if (e != cudaSuccess) return check_cuda(e, "cudaMemcpyToSymbol(c_num_blocks)");
    *nblocks = n;
    return CNERF_OK;
}

extern "C" int cnerf_weights_create(cnerf_weights** out) {
    CNERF_REQUIRE(out, "cnerf_weights_create: null out");
    cnerf_weights* w = new cnerf_weights();
    int rc = upload_program(&w->num_blocks);
    if (rc == CNERF_OK) rc = upload_bwd_program_once(&w->num_blocks_bwd);
    if (rc != CNERF_OK) { delete w; return rc; }
    cudaGetDevice(&w->device);
    cudaError_t e = cudaMalloc(&w->stream, (size_t)w->num_blocks * kBlockBytes);
    if (e == cudaSuccess) e = cudaMalloc(&w->stream_bwd, (size_t)w->num_blocks_bwd * kBlockBytes);
    if (e == cudaSuccess) e = cudaMalloc(&w->stream3, (size_t)stream3_blocks() * kBlockBytes);
    if (e == cudaSuccess) e = cudaMalloc(&w->stream_bwd3, (size_t)bwd_stream3_blocks() * kBlockBytes);
    if (e == cudaSuccess) e = cudaMalloc(&w->stream4, stream4_bytes());
    if (e == cudaSuccess) e = cudaMalloc(&w->misc, kMiscFloats * sizeof(float));
    if (e != cudaSuccess) { cudaFree(w->stream); cudaFree(w->stream_bwd); cudaFree(w->stream3); cudaFree(w->stream_bwd3); cudaFree(w->stream4); delete w; return check_cuda(e, "cudaMalloc(weights)"); }
    *out = w;
    return CNERF_OK;
}

extern "C" void cnerf_weights_destroy(cnerf_weights* w) {
    if (!w) return;
    cudaFree(w->stream);
    cudaFree(w->stream_bwd);
    cudaFree(w->stream3);
    cudaFree(w->stream_bwd3);
    cudaFree(w->stream4);
    cudaFree(w->misc);
    delete w;
}

extern "C" int cnerf_weights_refresh(cnerf_weights* w, const float* const* pts_w, const float* const* pts_b,
                                     const float* feature_w, const float* feature_b, const float* alpha_w,
                                     const float* alpha_b, const float* views_w, const float* views_b,
                                     const float* rgb_w, const float* rgb_b, void* stream) {
    CNERF_REQUIRE(w && pts_w && pts_b && feature_w && feature_b && alpha_w && alpha_b && views_w && views_b && rgb_w && rgb_b,
                  "cnerf_weights_refresh: null pointer");
    RawParams p;
    for (int i = 0; i < 8; ++i) {
        CNERF_REQUIRE(pts_w[i] && pts_b[i], "cnerf_weights_refresh: null pts_linears.%d", i);
        p.w[i] = pts_w[i]; p.b[i] = pts_b[i];
    }
    p.w[8] = feature_w; p.b[8] = feature_b; p.w[9] = views_w; p.b[9] = views_b;
    p.alpha_w = alpha_w; p.alpha_b = alpha_b; p.rgb_w = rgb_w; p.rgb_b = rgb_b;
    for (int i = 0; i < kNumLayers; ++i) p.ld[i] = kLayerLd[i];
    if (fwd_impl() == 1) {
        pack_weights_kernel<<<w->num_blocks, 256, 0, as_stream(stream)>>>(p, w->stream);
        CNERF_LAUNCH_CHECK("pack_weights_kernel");
    }
    pack_misc_kernel<<<ceil_div(kMiscFloats, 256), 256, 0, as_stream(stream)>>>(p, w->misc);
    CNERF_LAUNCH_CHECK("pack_misc_kernel");
    int rc = CNERF_OK;
    if (fwd_impl() != 1) rc = pack_stream3(p, w->stream3, as_stream(stream));
    if (rc == CNERF_OK && fwd_impl() == 4) rc = pack_stream4(p, w->stream4, as_stream(stream));
    if (rc == CNERF_OK) rc = bwd_impl() == 3 ? pack_bwd_stream3(p, w->stream_bwd3, as_stream(stream))
                                             : pack_bwd_stream(p, w->stream_bwd, w->num_blocks_bwd, as_stream(stream));
    if (rc != CNERF_OK) return rc;
    w->packed = true;
    return CNERF_OK;
}

static int launch_mlp(const cnerf_weights* w, const float* pts, const float* viewdirs, int n_rays, int n_samples,
                      float* raw, void* acts, void* stream, const char* who) {
    CNERF_REQUIRE(w && w->packed, "%s: weights handle not packed (call cnerf_weights_refresh)", who);
    CNERF_REQUIRE(pts && viewdirs && raw, "%s: null pointer", who);
    CNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "%s: bad sizes", who);
    int64_t np64 = (int64_t)n_rays * n_samples;
    CNERF_REQUIRE(np64 < (int64_t)1 << 30, "%s: too many points in one call (%lld)", who, (long long)np64);
    if (np64 == 0) return CNERF_OK;
    static bool attr_set = false;
    const int impl = fwd_impl();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);

        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(mlp_fused_kernel)");
        attr_set = true;
    }
    int n_points = (int)np64;
    int tiles = ceil_div(n_points, (int)kRows);
    int grid = tiles < kNumSMs ? tiles : kNumSMs;
    if (impl == 4 && !acts)      // the CTA-pair experiment is inference only; training runs the single-CTA kernel
        return launch_fused4(w->stream4, w->misc, pts, viewdirs, n_points, n_samples, n_rays, raw, as_stream(stream));
    if (impl != 1)
        return launch_fused3(w->stream3, w->misc, pts, viewdirs, n_points, n_samples, n_rays, raw, (uint8_t*)acts, as_stream(stream));
    if (acts)
        mlp_fused_kernel<true><<<grid, kThreads, kSmemTotal, as_stream(stream)>>>(w->stream, w->misc, pts, viewdirs, n_points,
                                                                                n_samples, n_rays, raw, (uint8_t*)acts);
    else
        mlp_fused_kernel<false><<<grid, kThreads, kSmemTotal, as_stream(stream)>>>(w->stream, w->misc, pts, viewdirs, n_points,
                                                                                 n_samples, n_rays, raw, nullptr);
    CNERF_LAUNCH_CHECK("mlp_fused_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_mlp_fwd(const cnerf_weights* w, const float* pts, const float* viewdirs, int n_rays, int n_samples,
                             float* raw, void* stream) {
    return launch_mlp(w, pts, viewdirs, n_rays, n_samples, raw, nullptr, stream, "cnerf_mlp_fwd");
}

extern "C" int64_t cnerf_mlp_acts_bytes(int64_t n_points) {
    return ceil_div64(n_points, kRows) * (int64_t)kTileBytes;
}

extern "C" int cnerf_mlp_fwd_train(const cnerf_weights* w, const float* pts, const float* viewdirs, int n_rays,
                                   int n_samples, float* raw, void* acts, void* stream) {
    CNERF_REQUIRE(acts, "cnerf_mlp_fwd_train: null activation record buffer");
    return launch_mlp(w, pts, viewdirs, n_rays, n_samples, raw, acts, stream, "cnerf_mlp_fwd_train");
}

extern "C" int cnerf_umma_selftest(const float* a, const float* b, int n, int k, float* d, void* stream) {
    CNERF_REQUIRE(a && b && d, "cnerf_umma_selftest: null pointer");
    CNERF_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k >= 16 && k <= 128 && k % 16 == 0, "cnerf_umma_selftest: bad n/k");
    size_t smem = (size_t)(k / 8) * 2048 * 2 + (size_t)(k / 8) * n * 16 * 2 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(umma_selftest_kernel)");
    umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(a, b, n, k, d);
    CNERF_LAUNCH_CHECK("umma_selftest_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_umma_selftest_ts(const float* a, const float* b, int n, int k, float* d, void* stream) {
    CNERF_REQUIRE(a && b && d, "cnerf_umma_selftest_ts: null pointer");
    CNERF_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && k >= 16 && k <= 256 && k % 16 == 0, "cnerf_umma_selftest_ts: bad n/k");
    size_t smem = (size_t)(k / 8) * n * 16 * 2 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(umma_selftest_ts_kernel)");
    umma_selftest_ts_kernel<<<1, 128, smem, as_stream(stream)>>>(a, b, n, k, d);
    CNERF_LAUNCH_CHECK("umma_selftest_ts_kernel");
    return CNERF_OK;
}

// Debug: cycles per tcgen05.mma (M=128, N=n, K=16) for `iters` back-to-back instructions on every SM; out[148] device floats.
extern "C" int cnerf_debug_umma_rate(int mode, int n, int iters, int alt, float* out, void* stream) {
    CNERF_REQUIRE(out && n >= 16 && n <= 256 && n % 16 == 0 && iters > 0, "cnerf_debug_umma_rate: bad arguments");
    size_t smem = 8192 + 32768 + 128;
    cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(umma_bench_kernel)");
    umma_bench_kernel<<<kNumSMs, 128, smem, as_stream(stream)>>>(mode, n, iters, alt, out);
    CNERF_LAUNCH_CHECK("umma_bench_kernel");
    return CNERF_OK;
}
