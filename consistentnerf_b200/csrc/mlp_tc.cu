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
#include "common.cuh"
#include <vector>

namespace cnerf {

// ------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("cnerf: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i <- lane (base+i), v[j] <- column (col+j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------
constexpr uint32_t kRows = 128;                 // tile rows (points) == N rows of one weight block
constexpr uint32_t kLBO = kRows * 16;           // 2048 B between k-groups (8 fp16 of K)
constexpr uint32_t kSBO = 128;                  // 8 rows x 16 B
// K-major, SWIZZLE_NONE, version 1 (sm_100): cute::UMMA::SmemDescriptor
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: c=F32 (bit4), a=b=F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc(uint32_t M, uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }

__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 8 fp32 -> one 16-byte hi store + one 16-byte lo store at (row, kgroup)
__device__ __forceinline__ void store_split8(uint32_t hi_base, uint32_t lo_base, uint32_t row, uint32_t kg, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    uint32_t off = kg * kLBO + row * 16;
    st_shared_v4(hi_base + off, h[0], h[1], h[2], h[3]);
    st_shared_v4(lo_base + off, l[0], l[1], l[2], l[3]);
}

// ------------------------------------------------------------------------------------
// The weight-block program: the order in which 16 KB blocks (128 out-rows x 32 k, hi+lo) are
// streamed and multiplied.  Built once on the host, mirrored in constant memory.
// ------------------------------------------------------------------------------------
constexpr int kNumLayers = 10;                  // 0-7 pts_linears, 8 feature_linear, 9 views_linears.0
constexpr int kMaxBlocks = 160;
constexpr uint32_t kBlockBytes = 16384;
constexpr uint32_t kBlockHalfBytes = 8192;

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

// misc fp32 parameter block
constexpr int kMiscBias = 0;                    // [10][256]
constexpr int kMiscAlphaW = 2560;               // [256]
constexpr int kMiscAlphaB = 2816;               // [1]
constexpr int kMiscRgbW = 2820;                 // [3][128]
constexpr int kMiscRgbB = 3204;                 // [3]
constexpr int kMiscFloats = 3208;

struct RawParams {
    const float* w[kNumLayers];
    const float* b[kNumLayers];
    const float* alpha_w; const float* alpha_b; const float* rgb_w; const float* rgb_b;
    int ld[kNumLayers];
};

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

// ------------------------------------------------------------------------------------
// shared-memory map of the fused kernel
// ------------------------------------------------------------------------------------
constexpr uint32_t kActHi = 0;                          // 32 k-groups x 2048 B  (K = 256)
constexpr uint32_t kActLo = 65536;
constexpr uint32_t kEmbHi = 131072;                     // 8 k-groups (K = 64): point encoding, later dir encoding
constexpr uint32_t kEmbLo = 147456;
constexpr uint32_t kRing = 163840;                      // 4 stages x 16 KB
constexpr int kStages = 4;
constexpr uint32_t kBars = kRing + kStages * kBlockBytes;      // 229376
constexpr uint32_t kTmemSlot = kBars + 96;
constexpr uint32_t kSAlpha = kBars + 128;               // float[128]
constexpr uint32_t kSRgb = kSAlpha + 512;               // float[3][128]
constexpr uint32_t kSmemTotal = kSRgb + 1536;           // 231552 <= 232448

constexpr int kEpiThreads = 256;
constexpr int kThreads = kEpiThreads + 64;
constexpr uint32_t kTmemCols = 512;

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

__global__ void __launch_bounds__(kThreads, 1)
mlp_fused_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc, const float* __restrict__ pts,
                 const float* __restrict__ viewdirs, int n_points, int n_samples, int n_rays, float* __restrict__ raw) {
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
                    if (bi.wait_a) { mbar_wait(bar_a, a_cnt & 1); ++a_cnt; tc_fence_after(); }
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
                    if (bi.last_of_layer) umma_commit(bar_d);  // accumulator of the layer is complete
                }
            }
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

}  // namespace cnerf

using namespace cnerf;

struct cnerf_weights {
    uint8_t* stream = nullptr;     // packed fp16 hi/lo blocks in program order
    float* misc = nullptr;         // biases + alpha/rgb heads (fp32)
    int num_blocks = 0;
    int device = -1;
    bool packed = false;
};

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
    if (rc != CNERF_OK) { delete w; return rc; }
    cudaGetDevice(&w->device);
    cudaError_t e = cudaMalloc(&w->stream, (size_t)w->num_blocks * kBlockBytes);
    if (e == cudaSuccess) e = cudaMalloc(&w->misc, kMiscFloats * sizeof(float));
    if (e != cudaSuccess) { cudaFree(w->stream); delete w; return check_cuda(e, "cudaMalloc(weights)"); }
    *out = w;
    return CNERF_OK;
}

extern "C" void cnerf_weights_destroy(cnerf_weights* w) {
    if (!w) return;
    cudaFree(w->stream);
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
    pack_weights_kernel<<<w->num_blocks, 256, 0, as_stream(stream)>>>(p, w->stream);
    CNERF_LAUNCH_CHECK("pack_weights_kernel");
    pack_misc_kernel<<<ceil_div(kMiscFloats, 256), 256, 0, as_stream(stream)>>>(p, w->misc);
    CNERF_LAUNCH_CHECK("pack_misc_kernel");
    w->packed = true;
    return CNERF_OK;
}

extern "C" int cnerf_mlp_fwd(const cnerf_weights* w, const float* pts, const float* viewdirs, int n_rays, int n_samples,
                             float* raw, void* stream) {
    CNERF_REQUIRE(w && w->packed, "cnerf_mlp_fwd: weights handle not packed (call cnerf_weights_refresh)");
    CNERF_REQUIRE(pts && viewdirs && raw, "cnerf_mlp_fwd: null pointer");
    CNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "cnerf_mlp_fwd: bad sizes");
    int64_t np64 = (int64_t)n_rays * n_samples;
    CNERF_REQUIRE(np64 < (int64_t)1 << 30, "cnerf_mlp_fwd: too many points in one call (%lld)", (long long)np64);
    if (np64 == 0) return CNERF_OK;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(mlp_fused_kernel)");
        attr_set = true;
    }
    int n_points = (int)np64;
    int tiles = ceil_div(n_points, (int)kRows);
    int grid = tiles < kNumSMs ? tiles : kNumSMs;
    mlp_fused_kernel<<<grid, kThreads, kSmemTotal, as_stream(stream)>>>(w->stream, w->misc, pts, viewdirs, n_points,
                                                                      n_samples, n_rays, raw);
    CNERF_LAUNCH_CHECK("mlp_fused_kernel");
    return CNERF_OK;
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
