// K2+K3 fused forward, fourth generation: CTA pairs (tcgen05 cta_group::2) with two 128-point tiles in ping-pong.
//
// What the third generation (mlp_fwd3.cu) left on the table (profiles/r1c_phase_profile_fwd3.txt): tcgen05.mma issue is
// nearly synchronous with execution, so every cycle the issuing warp waits -- for the epilogue of the previous layer of the
// SAME tile (22 k cycles per tile) or for a weight block of a 4-deep ring (21 k) -- idles the tensor pipe.  A second,
// independent tile fills those gaps, but two 128-row operand tiles (2 x 128 KB) do not fit one SM.  An M=128 pair MMA
// (64 rows per CTA) runs at the full per-row rate (64 cycles for N=256, K=16; scripts/umma_rate.py), so:
//
//   cluster   2 CTAs on the two SMs of a TPC; a super-tile is 256 consecutive points = tile X (128) + tile Y (128);
//             CTA r holds rows [64 r, 64 r + 64) of both tiles and HALF of every weight block (N/2 output rows)
//   MMA       issued by the leader CTA only: for layer L: X(L), Y(L) -- while the tensor cores of both SMs work on Y(L),
//             the 16 epilogue warps of each CTA turn D_X into the operand of X(L+1), and vice versa
//   TMEM      M=128 cta_group::2 accumulator: lane l, column j of CTA c  <->  row 64 c + l % 64, column (l / 64) N/2 + j
//             (scripts/pair_layout.py), i.e. 128 columns per tile, all four lane quarters busy in the epilogue
//   SMEM      per tile: operand [64 rows x 256 k] fp16 hi | lo (64 KB) + encoding [64 x 64] (16 KB); 8-stage ring of 8 KB
//             half-blocks (twice the depth of fwd3 in time); every layer's blocks are streamed once per tile
// Inference only (opt-in with CNERF_MLP_IMPL=4; training always runs mlp_fwd3.cu).  Measured slower than the single-CTA kernel
// (DESIGN.md section 3): with synchronous tcgen05.mma issue the per-block overhead of the issuing warps sits behind instructions
// half as long, and 8 KB half-blocks make the weight ring no deeper in time.  Kept as the parity-tested record of that experiment.
#include "mlp_blocks.cuh"

namespace cnerf {

constexpr int k4Threads = 672;                     // 16 epilogue warps; 16/19 loaders, 17/20 MMA issuers (peer CTA: forwarders) of tile X / Y; warp 18 idle
constexpr uint32_t k4LBO = 1024;                   // 64 rows x 16 B between k-groups of an operand tile
constexpr uint32_t k4ActBytes = 65536, k4ActLo = 32768;       // per tile: hi 32 KB | lo 32 KB (32 k-groups each)
constexpr uint32_t k4Emb = 131072, k4EmbBytes = 16384, k4EmbLo = 8192;   // per tile: hi 8 KB | lo 8 KB (8 k-groups each)
constexpr uint32_t k4Ring = 163840;
constexpr int k4Stages = 8;
constexpr uint32_t k4StageBytes = 8192;            // this CTA's half of a weight block: hi 4 KB | lo 4 KB
constexpr uint32_t k4Bars = k4Ring + k4Stages * k4StageBytes;     // 229376
constexpr uint32_t k4BarFull = k4Bars, k4BarEmpty = k4Bars + 64, k4BarPFull = k4Bars + 128, k4BarDFull = k4Bars + 192,
                   k4BarAReady = k4Bars + 208, k4TmemSlot = k4Bars + 256;
constexpr uint32_t k4Smem = k4Bars + 320;
constexpr uint32_t k4TmemCols = 256;               // two tiles x 128 columns

__device__ __forceinline__ int layer_first_block(int L) {
    return L == 0 ? 0 : L <= 4 ? 4 + 17 * (L - 1) : L == 5 ? 72 : L <= 8 ? 92 + 17 * (L - 6) : 143;
}
__device__ __forceinline__ int layer_num_blocks(int L) { return L == 0 ? 4 : L == 5 ? 20 : L == 9 ? 9 : 17; }

// stream4: block b of the fwd3 program, split by output rows: CTA r's half at (2 b + r) * 8 KB = [hi 4 KB | lo 4 KB]
//   layers 0-8: local rows nl = n - 128 r, element (nl, k) at (k/8)*2048 + nl*16 + (k%8)*2
//   layer 9   : local rows nl = n - 64 r,  element (nl, k) at (k/8)*1024 + nl*16 + (k%8)*2
__global__ void __launch_bounds__(256)
pack_weights4_kernel(RawParams p, uint8_t* __restrict__ stream) {
    const int b = blockIdx.x >> 1, r = blockIdx.x & 1;
    const Blk3 bi = block3_info(b);
    const float* W = p.w[bi.layer];
    const int ld = p.ld[bi.layer];
    uint8_t* dst = stream + (size_t)blockIdx.x * k4StageBytes;
    const int rows = bi.layer == 9 ? 64 : 128, kgs = bi.layer == 9 ? 4 : 2;
    for (int u = threadIdx.x; u < rows * kgs; u += 256) {
        const int nl = u % rows, kg = u / rows, n = r * rows + nl;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int k = kg * 8 + e;
            v[e] = (k < bi.kvalid) ? W[(size_t)n * ld + bi.src_k0 + k] : (k == bi.bias_k ? p.b[bi.layer][n] : 0.f);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pack2(v[2 * e], v[2 * e + 1], h[e], l[e]);
        size_t off = (size_t)kg * rows * 16 + (size_t)nl * 16;
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + 4096 + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// 8 fp32 -> hi / lo 16-byte words of one k-group of a 64-row operand tile
__device__ __forceinline__ void emit4(uint32_t hi_base, uint32_t lo_off, uint32_t row, uint32_t kg, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    const uint32_t a = hi_base + kg * k4LBO + row * 16;
    st_shared_v4(a, h[0], h[1], h[2], h[3]);
    st_shared_v4(a + lo_off, l[0], l[1], l[2], l[3]);
}
__device__ unsigned long long g_prof4[16];
__device__ int g_prof4_on;
__device__ int g_dbg4;          // timing experiments only (results become wrong): 1 = no waiting on weight blocks, 2 = commit only once per layer
#define PROF4_T0() long long pt0__ = g_prof4_on ? clock64() : 0
#define PROF4_ADD(var) do { if (g_prof4_on) { long long t__ = clock64(); var += t__ - pt0__; } } while (0)

__global__ void __launch_bounds__(k4Threads, 1)
mlp_fused4_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc,
                  const float* __restrict__ pts, const float* __restrict__ viewdirs, int n_points, int n_samples, int n_rays,
                  float* __restrict__ raw) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int num_super = (n_points + 255) / 256;
    const uint32_t bar_full = sbase + k4BarFull, bar_empty = sbase + k4BarEmpty, bar_pfull = sbase + k4BarPFull;
    const uint32_t bar_dfull = sbase + k4BarDFull, bar_aready = sbase + k4BarAReady;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + k4TmemSlot);

    if (threadIdx.x == 0) {
        for (int s = 0; s < k4Stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_pfull + 8 * s, 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_dfull + 8 * t, 1);
            mbar_init(bar_aready + 8 * t, 32);       // 16 epilogue warps of each CTA (used in the leader only)
        }
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc2(sbase + k4TmemSlot, k4TmemCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // tile t has its own weight ring (stages 4 t .. 4 t + 3), loader warp, MMA-issuing warp (leader CTA) and forwarder (peer CTA):
    // tcgen05.mma issue is nearly synchronous with execution (scripts/umma_rate.py), so one issuing thread cannot hide its own
    // per-block overhead and waits behind 64-cycle instructions -- two independent issuers cover each other's gaps.
    const bool is_loader = warp == 16 || warp == 19, is_mma = warp == 17 || warp == 20;
    const int wt = (warp == 19 || warp == 20) ? 1 : 0;                 // tile served by a loader / MMA / forwarder warp
    constexpr int kRingStages = k4Stages / 2;
    if (is_loader) {
        // ===== weight loader of tile wt: this CTA's half of every block of every layer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int S = pair; S < num_super; S += npairs)
                for (int L = 0; L < 10; ++L) {
                    const int b0 = layer_first_block(L), nb = layer_num_blocks(L);
                    for (int j = 0; j < nb; ++j, ++it) {
                        const uint32_t s = (uint32_t)wt * kRingStages + it % kRingStages, ph = (it / kRingStages) & 1;
                        if (g_dbg4 & 3) continue;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1);
                        mbar_arrive_expect_tx(bar_full + 8 * s, k4StageBytes);
                        bulk_g2s(sbase + k4Ring + s * k4StageBytes, wstream + ((size_t)(b0 + j) * 2 + rank) * k4StageBytes,
                                 k4StageBytes, bar_full + 8 * s);
                    }
                }
        }
        __syncwarp();
    } else if (is_mma && rank != 0) {
        // ===== peer: tell the leader when this CTA's half of a block of tile wt has landed =====
        if (lane == 0) {
            const uint32_t pf0 = mapa_u32(bar_pfull, 0);
            uint32_t it = 0;
            for (int S = pair; S < num_super; S += npairs)
                for (int j = 0; j < k3NumBlocks; ++j, ++it) {
                    const uint32_t s = (uint32_t)wt * kRingStages + it % kRingStages, ph = (it / kRingStages) & 1;
                    if (g_dbg4 & 3) continue;
                    mbar_wait(bar_full + 8 * s, ph);
                    mbar_arrive_cluster(pf0 + 8 * s);
                }
        }
        __syncwarp();
    } else if (is_mma) {
        // ===== leader: MMA issuer of tile wt (warp-uniform walk, one elected lane issues) =====
        constexpr uint32_t idesc256 = instr_desc(128, 256), idesc128 = instr_desc(128, 128);
        constexpr uint32_t kStep = 2 * (k4LBO >> 4);                                // one K=16 step of an operand tile
        const int t = wt;
        const uint32_t d = tmem + (uint32_t)t * 128;
        const uint64_t a_hi = smem_desc_any(sbase + (uint32_t)t * k4ActBytes, k4LBO, kSBO), a_lo = a_hi + (k4ActLo >> 4);
        const uint64_t e_hi = smem_desc_any(sbase + k4Emb + (uint32_t)t * k4EmbBytes, k4LBO, kSBO), e_lo = e_hi + (k4EmbLo >> 4);
        uint32_t it = 0;
        int sl = 0;
        const int dbg = g_dbg4;
        long long pw_a = 0, pw_full = 0, pw_issue = 0, p_start = g_prof4_on ? clock64() : 0;
        for (int S = pair; S < num_super; S += npairs, ++sl) {
#pragma unroll 1
            for (int L = 0; L < 10; ++L) {
                const int n_emb = (L == 0 || L == 5) ? 4 : 0, n_act = L == 0 ? 0 : (L == 9 ? 8 : 16);
                const int nb = layer_num_blocks(L);
                { PROF4_T0(); mbar_wait_cluster(bar_aready + 8 * t, (uint32_t)(sl * 10 + L) & 1); PROF4_ADD(pw_a); }
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < nb; ++j, ++it) {
                    const uint32_t s = (uint32_t)t * kRingStages + it % kRingStages, ph = (it / kRingStages) & 1;
                    if (!(dbg & 1)) { PROF4_T0(); mbar_wait(bar_full + 8 * s, ph); mbar_wait_cluster(bar_pfull + 8 * s, ph); PROF4_ADD(pw_full); }
                    tc_fence_after();
                    PROF4_T0();
                    if (elect_one()) {
                        const uint32_t ring = sbase + k4Ring + s * k4StageBytes;
                        if (L < 9) {
                            const uint64_t bh = smem_desc(ring), bl = bh + (4096 >> 4);
                            if (j < n_emb + n_act) {
                                const bool is_act = j >= n_emb;
                                const uint64_t ah = is_act ? a_hi + (uint64_t)((j - n_emb) * kStep) : e_hi + (uint64_t)(j * kStep);
                                const uint64_t al = is_act ? a_lo + (uint64_t)((j - n_emb) * kStep) : e_lo + (uint64_t)(j * kStep);
                                umma2_f16(d, ah, bh, idesc256, j == 0 ? 0u : 1u);
                                umma2_f16(d, ah, bl, idesc256, 1u);
                                umma2_f16(d, al, bh, idesc256, 1u);
                            } else {                      // bias block: encoding columns 48-63 (column 63 == 1.0) x [0 .. 0, bias]
                                umma2_f16(d, e_hi + 3 * kStep, bh, idesc256, 1u);
                                umma2_f16(d, e_hi + 3 * kStep, bl, idesc256, 1u);
                            }
                        } else {                          // views layer: N = 128, [64 local rows x 32 k] half-blocks
                            const uint64_t bh = smem_desc_any(ring, 1024, kSBO), bl = bh + (4096 >> 4);
                            constexpr uint32_t bStep = 2048 >> 4;
                            const uint64_t ah = j < 8 ? a_hi + (uint64_t)(j * 2 * kStep) : e_hi;
                            const uint64_t al = j < 8 ? a_lo + (uint64_t)(j * 2 * kStep) : e_lo;
                            umma2_f16(d, ah, bh, idesc128, j == 0 ? 0u : 1u);
                            umma2_f16(d, ah, bl, idesc128, 1u);
                            umma2_f16(d, al, bh, idesc128, 1u);
                            umma2_f16(d, ah + kStep, bh + bStep, idesc128, 1u);
                            umma2_f16(d, ah + kStep, bl + bStep, idesc128, 1u);
                            umma2_f16(d, al + kStep, bh + bStep, idesc128, 1u);
                        }
                        if (!(dbg & 2)) umma2_commit(bar_empty + 8 * s, 3);
                        if (j + 1 == nb) umma2_commit(bar_dfull + 8 * t, 3);
                    }
                    __syncwarp();
                    PROF4_ADD(pw_issue);
                }
            }
        }
        if (g_prof4_on && lane == 0 && t == 0) {
            atomicAdd(&g_prof4[0], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_prof4[1], (unsigned long long)pw_a); atomicAdd(&g_prof4[3], (unsigned long long)pw_full);
            atomicAdd(&g_prof4[4], (unsigned long long)pw_issue);
        }
    } else if (warp < 16) {
        // ===== prologue + epilogue warps =====
        // epilogue mapping: TMEM lane quarter q -> local row 32 (q & 1) + lane, column half h = q >> 1; p -> 32 of its 128 columns
        const int q = warp & 3, p = warp >> 2, h = q >> 1;
        const uint32_t irow = (uint32_t)((q & 1) * 32 + lane);
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        // prologue mapping: thread -> (row, k-group) of the 64 x 64 encoding tile
        const int tid = threadIdx.x;
        const uint32_t prow = (uint32_t)(tid & 63), pkg = (uint32_t)(tid >> 6);
        const uint32_t ar0 = mapa_u32(bar_aready, 0);
        long long pw_d = 0, p_start = g_prof4_on ? clock64() : 0;

        auto encode = [&](int S, int t, float* e8) {
            const int gp = S * 256 + t * 128 + (int)rank * 64 + (int)prow;
            float x[3] = {0.f, 0.f, 0.f};
            if (gp < n_points) { x[0] = pts[3 * (size_t)gp]; x[1] = pts[3 * (size_t)gp + 1]; x[2] = pts[3 * (size_t)gp + 2]; }
            switch (pkg) {
                case 0: enc8<0>(x, 63, e8); break;
                case 1: enc8<8>(x, 63, e8); break;
                case 2: enc8<16>(x, 63, e8); break;
                case 3: enc8<24>(x, 63, e8); break;
                case 4: enc8<32>(x, 63, e8); break;
                case 5: enc8<40>(x, 63, e8); break;
                case 6: enc8<48>(x, 63, e8); break;
                default: enc8<56>(x, 63, e8); e8[7] = 1.f; break;      // column 63: the constant 1 that carries the biases
            }
        };
        auto publish = [&](int t) {                           // operands of the next layer of tile t are in shared memory
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(ar0 + 8 * t);
            }
        };

        float enc[2][8];
        float alpha_part[2] = {0.f, 0.f};
        if (pair < num_super) { encode(pair, 0, enc[0]); encode(pair, 1, enc[1]); }
        int sl = 0;
        for (int S = pair; S < num_super; S += npairs, ++sl) {
            // ---- prologue: publish the (pre-computed) point encodings of both tiles
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                emit4(sbase + k4Emb + (uint32_t)t * k4EmbBytes, k4EmbLo, prow, pkg, enc[t]);
                publish(t);
            }
#pragma unroll 1
            for (int L = 0; L < 9; ++L) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    { PROF4_T0(); mbar_wait(bar_dfull + 8 * t, (uint32_t)(sl * 10 + L) & 1); PROF4_ADD(pw_d); }
                    tc_fence_after();
                    if (L == 5 && tid < 256) {
                        // every MMA of layer 5 of this tile (the last reader of point-encoding columns 0-31) is done: k-groups 0-3
                        // take the direction encoding (column 31 = 1.0 for the bias), published by this layer's arrivals
                        const int gp = S * 256 + t * 128 + (int)rank * 64 + (int)prow;
                        float dvec[3] = {0.f, 0.f, 0.f};
                        if (gp < n_points) {
                            int ray = min(gp / n_samples, n_rays - 1);
                            dvec[0] = viewdirs[3 * (size_t)ray]; dvec[1] = viewdirs[3 * (size_t)ray + 1]; dvec[2] = viewdirs[3 * (size_t)ray + 2];
                        }
                        float v[8];
                        if (pkg == 0)      enc8<0>(dvec, 27, v);
                        else if (pkg == 1) enc8<8>(dvec, 27, v);
                        else if (pkg == 2) enc8<16>(dvec, 27, v);
                        else             { enc8<24>(dvec, 27, v); v[7] = 1.f; }
                        emit4(sbase + k4Emb + (uint32_t)t * k4EmbBytes, k4EmbLo, prow, pkg, v);
                    }
                    float v[32];
                    if (!(g_dbg4 & 4)) {
                        tmem_ld32(t_lane + (uint32_t)t * 128 + (uint32_t)p * 32, v);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.01f * (float)(j + lane);
                    }
                    const bool relu = L != 8;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = relu ? fmaxf(v[j], 0.f) : fmaxf(v[j], -65504.f);
                        v[j] = fminf(x, 65504.f);
                    }
                    const uint32_t c0 = (uint32_t)h * 128 + (uint32_t)p * 32;       // first of this thread's 32 output columns
                    if (L == 7) {
                        float acc = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c0 + j));
                            acc = fmaf(v[j], a.x, acc); acc = fmaf(v[j + 1], a.y, acc); acc = fmaf(v[j + 2], a.z, acc); acc = fmaf(v[j + 3], a.w, acc);
                        }
                        alpha_part[t] = acc;
                    }
                    const uint32_t abase = sbase + (uint32_t)t * k4ActBytes;
#pragma unroll
                    for (uint32_t k4 = 0; k4 < 4; ++k4) if (!(g_dbg4 & 8)) emit4(abase, k4ActLo, irow, (c0 >> 3) + k4, v + 8 * k4);
                    publish(t);
                }
            }
            // the next super-tile's encodings are computed while the tensor cores work on the views layers
            if (S + npairs < num_super) { encode(S + npairs, 0, enc[0]); encode(S + npairs, 1, enc[1]); }
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                // views layer: ReLU (bias already in the accumulator), then rgb_linear as an fp32 dot product; 16 of 128 columns per thread
                { PROF4_T0(); mbar_wait(bar_dfull + 8 * t, (uint32_t)(sl * 10 + 9) & 1); PROF4_ADD(pw_d); }
                tc_fence_after();
                const uint32_t c0 = (uint32_t)h * 64 + (uint32_t)p * 16;
                float v[16];
                tmem_ld16(t_lane + (uint32_t)t * 128 + (uint32_t)p * 16, v);
                tmem_ld_wait();
                float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float hv = fmaxf(v[j], 0.f);
                    r0 = fmaf(hv, __ldg(misc + kMiscRgbW + c0 + j), r0);
                    r1 = fmaf(hv, __ldg(misc + kMiscRgbW + 128 + c0 + j), r1);
                    r2 = fmaf(hv, __ldg(misc + kMiscRgbW + 256 + c0 + j), r2);
                    v[j] = fminf(hv, 65504.f);
                }
                // partial sums of the two narrow heads -> scratch in this tile's encoding buffer (free: its last readers were the
                // views-layer MMAs and, in training, record stores that completed before the one waited for above)
                float* part = reinterpret_cast<float*>(smem + k4Emb + (uint32_t)t * k4EmbBytes);
                const int slot = h * 4 + p;
                part[(0 * 8 + slot) * 64 + irow] = r0;
                part[(1 * 8 + slot) * 64 + irow] = r1;
                part[(2 * 8 + slot) * 64 + irow] = r2;
                part[(3 * 8 + slot) * 64 + irow] = alpha_part[t];
                named_bar_sync(1, 512);
                if (tid < 64) {
                    const int gp = S * 256 + t * 128 + (int)rank * 64 + tid;
                    float o[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float a = 0.f;
#pragma unroll
                        for (int sl8 = 0; sl8 < 8; ++sl8) a += part[(c * 8 + sl8) * 64 + tid];
                        o[c] = a;
                    }
                    if (gp < n_points)
                        *reinterpret_cast<float4*>(raw + 4 * (size_t)gp) =
                            make_float4(o[0] + __ldg(misc + kMiscRgbB), o[1] + __ldg(misc + kMiscRgbB + 1), o[2] + __ldg(misc + kMiscRgbB + 2),
                                        o[3] + __ldg(misc + kMiscAlphaB));
                }
                tc_fence_before();
                named_bar_sync(1, 512);      // the scratch aliases the encoding tile the next prologue rewrites
            }
        }
        if (g_prof4_on && lane == 0 && warp == 0) {
            atomicAdd(&g_prof4[8], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_prof4[9], (unsigned long long)pw_d);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 17) tmem_dealloc2(tmem, k4TmemCols);
}

int pack_stream4(const RawParams& p, uint8_t* stream4, cudaStream_t st) {
    pack_weights4_kernel<<<2 * k3NumBlocks, 256, 0, st>>>(p, stream4);
    CNERF_LAUNCH_CHECK("pack_weights4_kernel");
    return CNERF_OK;
}
size_t stream4_bytes() { return (size_t)2 * k3NumBlocks * k4StageBytes; }

int launch_fused4(const uint8_t* stream4, const float* misc, const float* pts, const float* viewdirs, int n_points,
                  int n_samples, int n_rays, float* raw, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4Smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(mlp_fused4_kernel)");
        attr_set = true;
    }
    const int num_super = ceil_div(n_points, 256);
    const int pairs = num_super < kNumSMs / 2 ? num_super : kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(k4Threads); cfg.dynamicSmemBytes = k4Smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_fused4_kernel, stream4, misc, pts, viewdirs, n_points, n_samples, n_rays, raw);
    if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(mlp_fused4_kernel)");
    return CNERF_OK;
}

}  // namespace cnerf

// Debug: in-kernel phase profile of mlp_fused4_kernel (cycles summed over the 74 leader MMA warps / epilogue warp 0 of all CTAs):
//  [0] MMA warp total  [1] wait operand tiles  [3] wait weights  [4] MMA issue + commit   [8] epilogue total  [9] wait D
extern "C" int cnerf_debug_profile4(int enable, unsigned long long* out16) {
    using namespace cnerf;
    unsigned long long zero[16] = {0};
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out16) e = cudaMemcpyFromSymbol(out16, g_prof4, sizeof(zero));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_prof4, zero, sizeof(zero));
    const int on = enable & 1, dbg = enable >> 1;
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_prof4_on, &on, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_dbg4, &dbg, sizeof(int));
    if (e != cudaSuccess) return check_cuda(e, "cnerf_debug_profile4");
    return CNERF_OK;
}
