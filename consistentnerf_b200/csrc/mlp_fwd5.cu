// K2+K3 fused forward, fp16-operand generation: ONE tcgen05.mma per MAC (fp16 operands, fp32 accumulation in TMEM) and TWO
// 128-point tiles in flight per SM.
//
// Why: the three-term kernel (mlp_fwd3.cu) needs the hi AND lo halves of the A operand in shared memory (128 KB per tile), so
// only one tile fits per SM and every layer boundary stalls the tensor pipe on that tile's epilogue (phase profile: 35 % of a
// tile is operand / weight waits).  Measured on workload A (scripts/mma_terms.py, profiles/r2_mma_terms.txt): with plain fp16
// operands the rendered maps stay within 2e-5 of the fp64 oracle (bar: 1e-4).  With hi halves only the A operand is 64 KB per
// tile, two tiles fit, and the MMA warp alternates between them layer by layer: while the epilogue warps turn tile X's
// accumulator into its next A operand, the tensor pipe runs tile Y's layer.
//
//   TMEM   one 256-column fp32 accumulator per tile slot (2 x 256 = all 512 columns)
//   SMEM   per slot: A operand (K = 256, fp16, 64 KB) + encoding tile (K = 64, 16 KB);  8-stage ring of 8 KB weight units
//          ([256 out-rows x 16 k] fp16: the hi halves of the three-term kernel's blocks, same stream, same order)
//   sync   per slot: a_ready (16 warp arrivals: the slot's A operand / encoding is complete AND its accumulator has been read)
//          and d_full (tcgen05.commit: the layer's MMAs are complete).  Layer granularity -- no k-block pipelining is needed
//          because the other slot's layer fills the pipe.
//
// The training variant streams every A operand (hi halves) and the ReLU sign bits to the activation record, in the layout the
// backward kernels read (mlp_layout.cuh); it serves dw_terms == 1 only (the lo halves are never produced here).
#include "mlp_blocks.cuh"

namespace cnerf {

constexpr int k5Threads = 576;                            // 16 epilogue warps + loader warp + MMA warp
constexpr uint32_t k5Act = 0;                             // + slot * 65536: 32 k-groups x 2048 B
constexpr uint32_t k5Emb = 131072;                        // + slot * 16384: 8 k-groups
constexpr uint32_t k5Ring = 163840;
constexpr int k5Stages = 8;                             // power of two: the ring cursor advances with an AND, not a division
constexpr uint32_t k5Unit = kBlockHalfBytes;              // 8 KB: the hi half of a weight block
constexpr uint32_t k5Bars = k5Ring + k5Stages * k5Unit;   // 229376
constexpr uint32_t k5TmemSlot = k5Bars + 192;
constexpr uint32_t k5Smem = k5Bars + 256;
// first unit of layer l in the block stream (mlp_blocks.cuh): 0, 4, 21, 38, 55, 72, 92, 109, 126, 143, 152
__device__ __forceinline__ int k5_layer_first(int l) { return l == 0 ? 0 : l <= 5 ? 4 + 17 * (l - 1) : l <= 9 ? 92 + 17 * (l - 6) : 152; }

__device__ __forceinline__ void emit_hi(uint32_t base, uint32_t row, uint32_t kg, const float* v) {
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __half2 t = __floats2half2_rn(v[2 * i], v[2 * i + 1]); h[i] = *reinterpret_cast<uint32_t*>(&t); }
    st_shared_v4(base + kg * kLBO + row * 16, h[0], h[1], h[2], h[3]);
}

__device__ unsigned long long g_prof5[16];
#define PROF5_T0() long long pt0__ = kProf ? clock64() : 0
#define PROF5_ADD(var) do { if (kProf) { long long t__ = clock64(); var += t__ - pt0__; } } while (0)

// kProf: instrumented instantiation for the phase profile (cnerf_debug_profile5); the production kernels carry none of it
template <int kSave, bool kProf>
__global__ void __launch_bounds__(k5Threads, 1)
mlp_fused5_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc, const float* __restrict__ pts,
                  const float* __restrict__ viewdirs, int n_points, int n_samples, int n_rays, float* __restrict__ raw,
                  uint8_t* __restrict__ acts) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + k5Bars, bar_empty = bar_full + 8 * k5Stages;
    const uint32_t bar_dfull = bar_empty + 8 * k5Stages;        // [2]
    const uint32_t bar_aready = bar_dfull + 16;                  // [2]  16 warp arrivals
    const uint32_t bar_hv = bar_aready + 16;                     // [2]  training: views-layer output staged in the slot's A tile
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + k5TmemSlot);
    const int num_tiles = (n_points + (int)kRows - 1) / (int)kRows;
    // tile n of this CTA = blockIdx.x + n * gridDim.x; slot s works on n = 2 it + s
    const int my_tiles = ((int)blockIdx.x < num_tiles) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_iter = (my_tiles + 1) / 2;
    wstream += (size_t)(blockIdx.x % kWeightReplicas) * k3NumBlocks * kBlockBytes;      // this CTA's copy of the stream (mlp_layout.cuh)

    if (threadIdx.x == 0) {
        for (int s = 0; s < k5Stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_aready + 8 * s, 16); mbar_init(bar_hv + 8 * s, 16); }
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc(sbase + k5TmemSlot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        // ===== weight loader: units in exactly the order the MMA warp consumes them =====
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            uint32_t st = 0, ph = 0;
            for (int it = 0; it < n_iter; ++it)
                for (int layer = 0; layer < 10; ++layer)
                    for (int s = 0; s < 2; ++s) {
                        if (2 * it + s >= my_tiles) continue;
                        const uint8_t* src = wstream + (size_t)k5_layer_first(layer) * kBlockBytes;
                        for (int nb = k5_layer_first(layer + 1) - k5_layer_first(layer); nb > 0; --nb, src += kBlockBytes) {
                            mbar_wait(bar_empty + 8 * st, ph ^ 1);
                            mbar_arrive_expect_tx(bar_full + 8 * st, k5Unit);
                            bulk_g2s_hint(sbase + k5Ring + st * k5Unit, src, k5Unit, bar_full + 8 * st, keep);
                            st = (st + 1) & (k5Stages - 1);
                            ph ^= (st == 0);
                        }
                    }
        }
    } else if (warp == 17) {
        // ===== MMA issuer: alternates between the two slots layer by layer (warp-uniform walk, one elected lane issues) =====
        constexpr uint32_t idesc256 = instr_desc(128, 256), idesc128 = instr_desc(128, 128);
        constexpr uint32_t kStep = 2 * (kLBO >> 4);                               // two k-groups = one K=16 step of an A tile
        const uint64_t b256 = smem_desc_any(sbase + k5Ring, 4096, 128), b128 = smem_desc(sbase + k5Ring);
        const uint64_t stream_pol = l2_policy_evict_first();
        uint32_t st = 0, ph = 0;                                                   // ring cursor: stage, phase parity
        long long pw_a = 0, pw_full = 0, pw_issue = 0, pw_rec = 0, p_start = kProf ? clock64() : 0;
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll 1
            for (int layer = 0; layer < 10; ++layer) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const int n = 2 * it + s;
                    if (n >= my_tiles) continue;
                    const int tile = (int)blockIdx.x + n * (int)gridDim.x;
                    uint8_t* rec = kSave ? acts + (size_t)tile * kTileBytes : nullptr;
                    const uint32_t act = sbase + k5Act + (uint32_t)s * 65536, emb = sbase + k5Emb + (uint32_t)s * 16384;
                    const uint64_t act_d = smem_desc(act), emb_d = smem_desc(emb);
                    const uint32_t d = tmem + (uint32_t)s * 256;
                    { PROF5_T0(); mbar_wait(bar_aready + 8 * s, (uint32_t)(it * 10 + layer) & 1); PROF5_ADD(pw_a); }
                    tc_fence_after();
                    if (kSave && elect_one()) {      // the operand this layer reads is final: stream it to the record (hi halves)
                        if (layer == 0) bulk_s2g_hint(rec + kSlotE, emb, 16384, stream_pol);
                        else if (layer <= 8) bulk_s2g_hint(rec + kSlotH0 + (size_t)(layer - 1) * 131072, act, 65536, stream_pol);
                        else bulk_s2g_hint(rec + kSlotF, act, 65536, stream_pol);
                        if (layer == 6) bulk_s2g_hint(rec + kSlotV, emb, 16384, stream_pol);       // direction encoding (k-groups 0-3)
                        bulk_commit();
                    }
                    __syncwarp();
                    // The issue loop carries as little as possible: tcgen05.mma issue is synchronous with execution, so every instruction
                    // between two MMAs is added to the time per MMA (scripts/umma_rate.py).  One running A descriptor per segment, ring
                    // cursor advanced with an AND.
                    uint32_t acc = 0u;
                    auto segment = [&](uint64_t a, int count, uint32_t a_step) {
#pragma unroll 1
                        for (int j = 0; j < count; ++j, a += a_step) {
                            { PROF5_T0(); mbar_wait(bar_full + 8 * st, ph); PROF5_ADD(pw_full); }
                            tc_fence_after();
                            PROF5_T0();
                            if (elect_one()) {
                                if (layer < 9) {
                                    umma_f16(d, a, b256 + (uint64_t)(st * (k5Unit >> 4)), idesc256, acc);
                                } else {                                              // views layer: [128 x 32] units, N = 128, two K = 16 steps
                                    const uint64_t b = b128 + (uint64_t)(st * (k5Unit >> 4));
                                    umma_f16(d, a, b, idesc128, acc);
                                    umma_f16(d, a + kStep, b + kStep, idesc128, 1u);
                                }
                                umma_commit(bar_empty + 8 * st);
                            }
                            __syncwarp();
                            acc = 1u;
                            st = (st + 1) & (k5Stages - 1);
                            ph ^= (st == 0);
                            PROF5_ADD(pw_issue);
                        }
                    };
                    if (layer == 0) segment(emb_d, 4, kStep);
                    else if (layer == 5) { segment(emb_d, 4, kStep); segment(act_d, 16, kStep); }
                    else if (layer < 9) { segment(act_d, 16, kStep); segment(emb_d + 3 * kStep, 1, 0); }      // bias unit: encoding columns 48-63 (column 63 == 1.0)
                    else { segment(act_d, 8, 2 * kStep); segment(emb_d, 1, 0); }                              // feature columns, then the direction encoding
                    // training: the previous tile's views output (staged in this slot's A tile by its last epilogue) goes to the record
                    // now, behind layer 0's MMAs (which read the encoding tile only); it must have left before epilogue 0 rewrites A
                    if (kSave && layer == 0 && it > 0) {
                        mbar_wait(bar_hv + 8 * s, (uint32_t)(it - 1) & 1);
                        if (elect_one()) { bulk_s2g_hint(rec - (size_t)2 * gridDim.x * kTileBytes + kSlotHV, act, 32768, stream_pol); bulk_commit(); }
                        __syncwarp();
                    }
                    {
                        PROF5_T0();
                        if (elect_one()) {
                            if (kSave) bulk_wait_read0();       // the epilogue overwrites the operand tile once it sees this layer done
                            umma_commit(bar_dfull + 8 * s);
                        }
                        __syncwarp();
                        PROF5_ADD(pw_rec);
                    }
                }
            }
        }
        if (kProf && lane == 0) {
            atomicAdd(&g_prof5[0], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_prof5[1], (unsigned long long)pw_a); atomicAdd(&g_prof5[3], (unsigned long long)pw_full);
            atomicAdd(&g_prof5[4], (unsigned long long)pw_issue); atomicAdd(&g_prof5[5], (unsigned long long)pw_rec);
        }
        if (kSave) {      // views outputs of the last tile of each slot
            for (int s = 0; s < 2; ++s) {
                const int last_it = (my_tiles - 1 - s) / 2;      // last iteration in which slot s had a tile
                if (my_tiles <= s) continue;
                const int tile = (int)blockIdx.x + (2 * last_it + s) * (int)gridDim.x;
                mbar_wait(bar_hv + 8 * s, (uint32_t)last_it & 1);
                if (elect_one()) {
                    bulk_s2g_hint(acts + (size_t)tile * kTileBytes + kSlotHV, sbase + k5Act + (uint32_t)s * 65536, 32768, stream_pol);
                    bulk_commit();
                }
                __syncwarp();
            }
            if (elect_one()) bulk_wait0();
            __syncwarp();
        }
    } else {
        // ===== prologue + epilogue warps: thread = (row, p); per 32-column k-block it owns columns 8p..8p+7; both slots in turn =====
        const int q = warp & 3, p = warp >> 2;
        const uint32_t row = (uint32_t)(q * 32 + lane);
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        // point encoding of this thread's 16 columns (k-groups 2p, 2p+1) -> the slot's encoding tile; column 63 carries the biases
        auto publish_encoding = [&](int tile, int s) {
            const int gr = tile * (int)kRows + (int)row;
            float x[3] = {0.f, 0.f, 0.f}, e16[16];
            if (gr < n_points) { x[0] = pts[3 * (size_t)gr]; x[1] = pts[3 * (size_t)gr + 1]; x[2] = pts[3 * (size_t)gr + 2]; }
            if (p == 0)      { enc8<0>(x, 63, e16);  enc8<8>(x, 63, e16 + 8); }
            else if (p == 1) { enc8<16>(x, 63, e16); enc8<24>(x, 63, e16 + 8); }
            else if (p == 2) { enc8<32>(x, 63, e16); enc8<40>(x, 63, e16 + 8); }
            else             { enc8<48>(x, 63, e16); enc8<56>(x, 63, e16 + 8); e16[15] = 1.f; }
            const uint32_t eb = sbase + k5Emb + (uint32_t)s * 16384;
            emit_hi(eb, row, 2 * (uint32_t)p, e16);
            emit_hi(eb, row, 2 * (uint32_t)p + 1, e16 + 8);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready + 8 * s);
        };
        for (int s = 0; s < 2; ++s)
            if (s < my_tiles) publish_encoding((int)blockIdx.x + s * (int)gridDim.x, s);
        long long pw_d = 0, pw_ld = 0, pw_body = 0, p_start = kProf ? clock64() : 0;
        float alpha0 = 0.f, alpha1 = 0.f;                 // alpha_linear partial dot products of the two slots (scalars: no dynamic indexing)
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll 1
            for (int layer = 0; layer < 10; ++layer) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const int n = 2 * it + s;
                    if (n >= my_tiles) continue;
                    const int tile = (int)blockIdx.x + n * (int)gridDim.x;
                    const int grow = tile * (int)kRows + (int)row;
                    const bool valid = grow < n_points;
                    const uint32_t ab = sbase + k5Act + (uint32_t)s * 65536, eb = sbase + k5Emb + (uint32_t)s * 16384;
                    { PROF5_T0(); mbar_wait(bar_dfull + 8 * s, (uint32_t)(it * 10 + layer) & 1); PROF5_ADD(pw_d); }
                    tc_fence_after();
                    PROF5_T0();
                    if (layer < 9) {
                        if (layer == 0) { if (s) alpha1 = 0.f; else alpha0 = 0.f; }
                        if (layer == 5) {
                            // every MMA of layer 5 (the last reader of encoding columns 0-31) is done: k-groups 0-3 take the direction
                            // encoding (column 31 = 1.0 for the views bias)
                            float dvec[3] = {0.f, 0.f, 0.f}, v[8];
                            if (valid) {
                                const int ray = min(grow / n_samples, n_rays - 1);
                                dvec[0] = viewdirs[3 * (size_t)ray]; dvec[1] = viewdirs[3 * (size_t)ray + 1]; dvec[2] = viewdirs[3 * (size_t)ray + 2];
                            }
                            if (p == 0)      enc8<0>(dvec, 27, v);
                            else if (p == 1) enc8<8>(dvec, 27, v);
                            else if (p == 2) enc8<16>(dvec, 27, v);
                            else             { enc8<24>(dvec, 27, v); v[7] = 1.f; }
                            emit_hi(eb, row, (uint32_t)p, v);
                        }
                        const bool relu = layer != 8;
                        uint32_t mbits_lo = 0, mbits_hi = 0;
                        const uint32_t dcol = t_lane + (uint32_t)s * 256 + (uint32_t)p * 8;
#pragma unroll
                        for (uint32_t half = 0; half < 2; ++half) {
                            float v[32];
                            long long tl0 = kProf ? clock64() : 0;
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4) tmem_ld8(dcol + (half * 4 + k4) * 32, v + 8 * k4);
                            tmem_ld_wait();
                            if (kProf) pw_ld += clock64() - tl0;
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4) {
                                const uint32_t kb = half * 4 + k4;
                                float* w = v + 8 * k4;
                                if (kSave) {
                                    const uint32_t bits = sign_clear_bits8(w);
                                    if (half == 0) mbits_lo |= bits << (8 * k4); else mbits_hi |= bits << (8 * k4);
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float t = relu ? fmaxf(w[j], 0.f) : fmaxf(w[j], -65504.f);
                                    w[j] = fminf(t, 65504.f);
                                }
                                if (layer == 7) {
                                    const uint32_t c = kb * 32 + (uint32_t)p * 8;
                                    const float4 a0 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c)), a1 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c + 4));
                                    float acc = s ? alpha1 : alpha0;
                                    acc = fmaf(w[0], a0.x, acc); acc = fmaf(w[1], a0.y, acc); acc = fmaf(w[2], a0.z, acc); acc = fmaf(w[3], a0.w, acc);
                                    acc = fmaf(w[4], a1.x, acc); acc = fmaf(w[5], a1.y, acc); acc = fmaf(w[6], a1.z, acc); acc = fmaf(w[7], a1.w, acc);
                                    if (s) alpha1 = acc; else alpha0 = acc;
                                }
                                emit_hi(ab, row, kb * 4 + (uint32_t)p, w);
                            }
                        }
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_aready + 8 * s);
                        if (kSave && relu)
                            *reinterpret_cast<uint2*>(acts + (size_t)tile * kTileBytes + kSlotM + (size_t)layer * 4096 + (size_t)p * 1024 + row * 8) =
                                make_uint2(mbits_lo, mbits_hi);
                    } else {
                        // views layer: ReLU (bias already in the accumulator), rgb_linear as an fp32 dot product; 32 of 128 columns per thread
                        const uint32_t c = (uint32_t)p * 32;
                        float v[32];
                        tmem_ld32(t_lane + (uint32_t)s * 256 + c, v);
                        tmem_ld_wait();
                        tc_fence_before();
                        // the accumulator is in registers and every reader of the slot's encoding tile is done: hand the slot's NEXT
                        // tile to the MMA warp first, so that its layer 0 runs while these warps finish this tile
                        if (n + 2 < my_tiles) publish_encoding(tile + 2 * (int)gridDim.x, s);
                        if (kSave) {
                            uint32_t word = 0;
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) word |= sign_clear_bits8(v + 8 * k4) << (8 * k4);
                            *reinterpret_cast<uint32_t*>(acts + (size_t)tile * kTileBytes + kSlotM + 32768 + row * 16 + (c >> 3)) = word;
                        }
                        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float hv = fmaxf(v[j], 0.f);
                            r0 = fmaf(hv, __ldg(misc + kMiscRgbW + c + j), r0);
                            r1 = fmaf(hv, __ldg(misc + kMiscRgbW + 128 + c + j), r1);
                            r2 = fmaf(hv, __ldg(misc + kMiscRgbW + 256 + c + j), r2);
                            v[j] = fminf(hv, 65504.f);
                        }
                        // scratch for the head reductions: the upper half of the slot's A tile (its last readers, this layer's MMAs, are
                        // done; training stages hv in k-groups 0-15; the next writer is the slot's next epilogue 0, after the barrier below)
                        float* s_rgb = reinterpret_cast<float*>(smem + k5Act + (size_t)s * 65536 + 32768);
                        float* s_alpha = s_rgb + 1152;
                        if (kSave) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) emit_hi(ab, row, (c >> 3) + k4, v + 8 * k4);
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_hv + 8 * s);
                        }
                        if (p > 0) { float* o = s_rgb + (p - 1) * 384; o[row] = r0; o[128 + row] = r1; o[256 + row] = r2; }
                        s_alpha[p * 128 + row] = s ? alpha1 : alpha0;
                        named_bar_sync(1, 512);
                        if (p == 0 && valid) {
                            float4 o;
                            o.x = r0 + s_rgb[row] + s_rgb[384 + row] + s_rgb[768 + row] + __ldg(misc + kMiscRgbB);
                            o.y = r1 + s_rgb[128 + row] + s_rgb[512 + row] + s_rgb[896 + row] + __ldg(misc + kMiscRgbB + 1);
                            o.z = r2 + s_rgb[256 + row] + s_rgb[640 + row] + s_rgb[1024 + row] + __ldg(misc + kMiscRgbB + 2);
                            o.w = s_alpha[row] + s_alpha[128 + row] + s_alpha[256 + row] + s_alpha[384 + row] + __ldg(misc + kMiscAlphaB);
                            *reinterpret_cast<float4*>(raw + 4 * (size_t)grow) = o;
                        }
                        named_bar_sync(1, 512);      // the scratch lives in the A tile the slot's next epilogue rewrites
                    }
                    PROF5_ADD(pw_body);
                }
            }
        }
        if (kProf && lane == 0 && warp == 0) {
            atomicAdd(&g_prof5[8], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_prof5[9], (unsigned long long)pw_d); atomicAdd(&g_prof5[10], (unsigned long long)pw_ld);
            atomicAdd(&g_prof5[11], (unsigned long long)pw_body);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem, 512);
}

static int g_prof5_host = 0;

int launch_fused5(const uint8_t* stream3, const float* misc, const float* pts, const float* viewdirs, int n_points,
                  int n_samples, int n_rays, float* raw, uint8_t* acts, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused5_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fused5_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fused5_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5Smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fused5_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5Smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(mlp_fused5_kernel)");
        attr_set = true;
    }
    const int tiles = ceil_div(n_points, (int)kRows);
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
#define CNERF_F5(S, P) mlp_fused5_kernel<S, P><<<grid, k5Threads, k5Smem, st>>>(stream3, misc, pts, viewdirs, n_points, n_samples, n_rays, raw, acts)
    if (g_prof5_host) { if (acts) CNERF_F5(1, true); else CNERF_F5(0, true); }
    else { if (acts) CNERF_F5(1, false); else CNERF_F5(0, false); }
#undef CNERF_F5
    CNERF_LAUNCH_CHECK("mlp_fused5_kernel");
    return CNERF_OK;
}

}  // namespace cnerf

// Debug: in-kernel phase profile of mlp_fused5_kernel (cycles summed over CTAs):
//  [0] MMA warp total  [1] wait A operand  [3] wait weights  [4] MMA issue + commit  [5] record read-wait + layer commit
//  [8] epilogue warp 0 total  [9] wait D  [10] tcgen05.ld + wait  [11] epilogue body (incl. [10])
extern "C" int cnerf_debug_profile5(int enable, unsigned long long* out16) {
    using namespace cnerf;
    unsigned long long zero[16] = {0};
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out16) e = cudaMemcpyFromSymbol(out16, g_prof5, sizeof(zero));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_prof5, zero, sizeof(zero));
    if (e != cudaSuccess) return check_cuda(e, "cnerf_debug_profile5");
    g_prof5_host = enable & 1;
    return CNERF_OK;
}
