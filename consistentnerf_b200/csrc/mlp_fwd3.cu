// K2+K3 fused forward, third generation: N=256 MMAs, double-buffered accumulators, k-block pipelining.
//
// Measured on B200 (scripts/umma_rate.py, profiles/r1e_umma_issue_rate.txt): one tcgen05.mma M=128 x N=256 x K=16 runs at
// 128 cycles (the pipe rate) back to back, N=128 costs the same 128 cycles (operand-A fetch bound), and issue is SYNCHRONOUS
// with execution: every instruction the issuing warp executes between MMAs is added to the time per MMA.  Consequences:
// every 256-wide layer is issued as N=256 instructions, and the issue loop carries nothing it does not need (the phase
// profile is a separate template instantiation, kProf).
//
//   TMEM   two 256-column fp32 accumulators; layer L accumulates into D[L & 1]
//   SMEM   A operand: activations of the current layer (K=256, fp16 hi/lo, 128 KB) + encoding tile (32 KB),
//          4-stage ring of 16 KB weight blocks ([256 out-rows x 16 k], hi 8 KB + lo 8 KB), fetched with the L2 evict_last policy
//
// The epilogue of layer L (16 warps) walks the accumulator k-block by k-block (32 columns of D = one 32-wide k-block
// of the next layer's A operand) and signals each finished k-block through its own mbarrier; the MMA warp starts
// layer L+1 on k-block 0 while the epilogue is still converting k-blocks 1..7 -- it writes the OTHER accumulator, so
// the only serialisation left is the first k-block.  The next tile's encoding is published ahead of the last (views-layer)
// epilogue, so its layer 0 overlaps that epilogue.  The training variant (kSave) streams every finished A k-block to the
// activation record with bulk S2G copies issued by the MMA warp (L2 evict_first); measured alternatives -- a dedicated
// streaming warp, direct st.global from the epilogue -- were slower.
#include "mlp_blocks.cuh"

namespace cnerf {

__device__ unsigned long long g_prof3[16];
__device__ int g_prof3_on;
__device__ int g_dbg3;          // timing experiments only (results become wrong): 1 skip tcgen05.ld, 2 skip operand stores, 4 skip proxy fences
// the phase profile is a separate template instantiation (kProf): every instruction in the MMA-issuing warp delays the tensor
// pipe (tcgen05.mma issue is synchronous with execution, scripts/umma_rate.py), so the production kernels carry none of it
#define PROF_T0() long long pt0__ = kProf ? clock64() : 0
#define PROF_ADD(var) do { if (kProf) { long long t__ = clock64(); var += t__ - pt0__; } } while (0)

constexpr int k3Threads = 576;                           // 16 epilogue warps + loader warp + MMA warp
constexpr uint32_t k3ActHi = 0, k3ActLo = 65536;         // 32 k-groups x 2048 B each
constexpr uint32_t k3EmbHi = 131072, k3EmbLo = 147456;   // 8 k-groups each
constexpr uint32_t k3Ring = 163840;
constexpr int k3Stages = 4;
constexpr uint32_t k3Bars = k3Ring + k3Stages * kBlockBytes;      // 229376
constexpr uint32_t k3TmemSlot = k3Bars + 192;
constexpr uint32_t k3Smem = k3Bars + 256;
__global__ void __launch_bounds__(256)
pack_weights3_kernel(RawParams p, uint8_t* __restrict__ stream) {
    const Blk3 bi = block3_info(blockIdx.x);
    const float* W = p.w[bi.layer];
    const int ld = p.ld[bi.layer];
    uint8_t* dst = stream + ((size_t)blockIdx.y * k3NumBlocks + blockIdx.x) * kBlockBytes;      // blockIdx.y: replica
    const int rows = bi.layer == 9 ? 128 : 256, kgs = bi.layer == 9 ? 4 : 2;
    for (int u = threadIdx.x; u < rows * kgs; u += 256) {
        const int n = u % rows, kg = u / rows;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int k = kg * 8 + e;
            v[e] = (k < bi.kvalid) ? W[(size_t)n * ld + bi.src_k0 + k] : (k == bi.bias_k ? p.b[bi.layer][n] : 0.f);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pack2(v[2 * e], v[2 * e + 1], h[e], l[e]);
        size_t off = (size_t)kg * rows * 16 + (size_t)n * 16;
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + kBlockHalfBytes + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
// (__maxnreg__(104/112) instead of the launch bound would avoid the few spills of the training variant, but 576 threads then
// exceed the register file's allocation granularity: "too many resources requested for launch")
// kSave: 0 inference; 1 training, the record carries the fp16 hi halves only (weight gradients in one fp16 MMA per MAC);
//        2 training, hi + lo halves (weight gradients with the full three-term split).
// kVar:  measurement variants only (cnerf_debug_mlp_fwd_terms): `terms` selects which of the three partial products are issued
//        (bit 0 a_hi*w_hi, bit 1 a_hi*w_lo, bit 2 a_lo*w_hi); the production instantiations issue all three unconditionally.
template <int kSave, bool kProf, bool kVar>
__global__ void __launch_bounds__(k3Threads, 1)
mlp_fused3_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc, const float* __restrict__ pts,
                  const float* __restrict__ viewdirs, int n_points, int n_samples, int n_rays, float* __restrict__ raw,
                  uint8_t* __restrict__ acts, int terms) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + k3Bars, bar_empty = bar_full + 8 * k3Stages;
    const uint32_t bar_dfull = bar_empty + 8 * k3Stages;        // [2]  accumulator D[i] complete
    const uint32_t bar_aready = bar_dfull + 16;                  // [8]  A k-block written (16 warp arrivals)
    const uint32_t bar_eready = bar_aready + 64;                 //      encoding tile written (16 warp arrivals)
    const uint32_t bar_hv = bar_eready + 8;                      //      training: views-layer output staged in SMEM
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + k3TmemSlot);
    const int num_tiles = (n_points + (int)kRows - 1) / (int)kRows;
    wstream += (size_t)(blockIdx.x % kWeightReplicas) * k3NumBlocks * kBlockBytes;

    if (threadIdx.x == 0) {
        for (int s = 0; s < k3Stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_dfull, 1); mbar_init(bar_dfull + 8, 1);
        for (int k = 0; k < 8; ++k) mbar_init(bar_aready + 8 * k, 16);      // one arrival per epilogue warp
        mbar_init(bar_eready, 16);
        mbar_init(bar_hv, 16);
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc(sbase + k3TmemSlot, 512);
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
                for (int b = 0; b < k3NumBlocks; ++b, ++it) {
                    const uint32_t s = it % k3Stages, ph = (it / k3Stages) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(bar_full + 8 * s, kBlockBytes);
                    bulk_g2s_hint(sbase + k3Ring + s * kBlockBytes, wstream + (size_t)b * kBlockBytes, kBlockBytes, bar_full + 8 * s, keep);
                }
        }
    } else if (warp == 17) {
        // ===== MMA issuer: the whole warp walks the static program (warp-uniform control flow), one elected lane issues =====
        constexpr uint32_t idesc256 = instr_desc(128, 256), idesc128 = instr_desc(128, 128);
        const uint64_t b256 = smem_desc_any(sbase + k3Ring, 4096, 128);          // + stage*1024 ; lo: +512
        const uint64_t b128 = smem_desc(sbase + k3Ring);
        const uint64_t act_hi = smem_desc(sbase + k3ActHi), act_lo = smem_desc(sbase + k3ActLo);
        const uint64_t emb_hi = smem_desc(sbase + k3EmbHi), emb_lo = smem_desc(sbase + k3EmbLo);
        constexpr uint32_t kStep = 2 * (kLBO >> 4);                               // two k-groups = one K=16 step of the A tile
        const uint64_t stream_pol = l2_policy_evict_first();
        uint32_t it = 0;
        long long pw_a = 0, pw_full = 0, pw_e = 0, pw_issue = 0, p_start = kProf ? clock64() : 0;
        int tl = 0;
        auto store_hv = [&](uint8_t* rec_of_tile, uint32_t tl_of_tile) {      // whole warp; record slot HV of a finished tile
            mbar_wait(bar_hv, tl_of_tile & 1);
            if (elect_one()) {
                bulk_s2g_hint(rec_of_tile + kSlotHV, sbase + k3ActHi, 32768, stream_pol);
                if (kSave == 2) bulk_s2g_hint(rec_of_tile + kSlotHV + 32768, sbase + k3ActLo, 32768, stream_pol);
                bulk_commit();
            }
            __syncwarp();
        };
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            uint8_t* rec = kSave ? acts + (size_t)tile * kTileBytes : nullptr;
            { PROF_T0(); mbar_wait(bar_eready, (uint32_t)tl & 1); PROF_ADD(pw_e); }
            if (kSave && elect_one()) { bulk_s2g_hint(rec + kSlotE, sbase + k3EmbHi, kSave == 2 ? 32768 : 16384, stream_pol); bulk_commit(); }
            __syncwarp();
#pragma unroll 1
            for (int layer = 0; layer < 9; ++layer) {
                const uint32_t d = tmem + (uint32_t)(layer & 1) * 256;
                const int n_emb = (layer == 0 || layer == 5) ? 4 : 0, n_act = layer == 0 ? 0 : 16, n_bias = n_emb ? 0 : 1;
                const int nb = n_emb + n_act + n_bias;
                const uint32_t aph = (uint32_t)(tl * 9 + layer - 1) & 1;
                uint8_t* slot = rec + kSlotH0 + (size_t)(layer - 1) * 131072;             // A of this layer = output of layer - 1
#pragma unroll 1
                for (int j = 0; j < nb; ++j, ++it) {
                    const bool is_act = j >= n_emb && j < n_emb + n_act;
                    const int ja = j - n_emb;
                    if (is_act && !(ja & 1)) {
                        const int kb = ja >> 1;
                        { PROF_T0(); mbar_wait(bar_aready + 8 * kb, aph); PROF_ADD(pw_a); }
                        if (kSave && elect_one()) {           // this k-block of the A operand is final: stream it to the record
                            bulk_s2g_hint(slot + (size_t)kb * 8192, sbase + k3ActHi + kb * 8192, 8192, stream_pol);
                            if (kSave == 2) bulk_s2g_hint(slot + 65536 + (size_t)kb * 8192, sbase + k3ActLo + kb * 8192, 8192, stream_pol);
                            if (layer == 6 && kb == 0) bulk_s2g_hint(rec + kSlotV, sbase + k3EmbHi, kSave == 2 ? 32768 : 16384, stream_pol);
                            bulk_commit();
                        }
                        __syncwarp();
                    }
                    // the previous tile's views output must leave the activation tile before this tile's first epilogue rewrites it
                    // (bulk_wait_read0 ahead of the commit below): stored here, after layer 0's MMAs were issued
                    if (kSave && layer == 0 && tl > 0 && j + 1 == nb) store_hv(rec - (size_t)gridDim.x * kTileBytes, (uint32_t)(tl - 1));
                    const uint32_t s = it % k3Stages, ph = (it / k3Stages) & 1;
                    { PROF_T0(); mbar_wait(bar_full + 8 * s, ph); PROF_ADD(pw_full); }
                    tc_fence_after();
                    PROF_T0();
                    if (elect_one()) {
                        const uint64_t bh = b256 + (uint64_t)(s * (kBlockBytes >> 4)), bl = bh + (kBlockHalfBytes >> 4);
                        if (is_act || j < n_emb) {
                            const uint64_t ah = (is_act ? act_hi + (uint64_t)(ja * kStep) : emb_hi + (uint64_t)(j * kStep));
                            const uint64_t al = (is_act ? act_lo + (uint64_t)(ja * kStep) : emb_lo + (uint64_t)(j * kStep));
                            umma_f16(d, ah, bh, idesc256, j == 0 ? 0u : 1u);
                            if (!kVar || (terms & 2)) umma_f16(d, ah, bl, idesc256, 1u);
                            if (!kVar || (terms & 4)) umma_f16(d, al, bh, idesc256, 1u);
                        } else {                              // bias block: encoding columns 48-63 (column 63 == 1.0) x [0 .. 0, bias]
                            umma_f16(d, emb_hi + 3 * kStep, bh, idesc256, 1u);
                            if (!kVar || (terms & 2)) umma_f16(d, emb_hi + 3 * kStep, bl, idesc256, 1u);
                        }
                        umma_commit(bar_empty + 8 * s);
                        if (j + 1 == nb) {
                            if (kSave) bulk_wait_read0();     // the epilogue may overwrite the operand tiles once it sees this layer done
                            umma_commit(bar_dfull + 8 * (layer & 1));
                        }
                    }
                    __syncwarp();
                    PROF_ADD(pw_issue);
                }
            }
            {   // views layer: N = 128, [128 x 32] blocks, accumulator D[1] columns 256..383
                const uint32_t d = tmem + 256;
                const uint32_t aph = (uint32_t)(tl * 9 + 8) & 1;
#pragma unroll 1
                for (int j = 0; j < 9; ++j, ++it) {
                    if (j < 8) {
                        { PROF_T0(); mbar_wait(bar_aready + 8 * j, aph); PROF_ADD(pw_a); }
                        if (kSave && elect_one()) {
                            bulk_s2g_hint(rec + kSlotF + (size_t)j * 8192, sbase + k3ActHi + j * 8192, 8192, stream_pol);
                            if (kSave == 2) bulk_s2g_hint(rec + kSlotF + 65536 + (size_t)j * 8192, sbase + k3ActLo + j * 8192, 8192, stream_pol);
                            bulk_commit();
                        }
                        __syncwarp();
                    }
                    const uint32_t s = it % k3Stages, ph = (it / k3Stages) & 1;
                    { PROF_T0(); mbar_wait(bar_full + 8 * s, ph); PROF_ADD(pw_full); }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bh = b128 + (uint64_t)(s * (kBlockBytes >> 4)), bl = bh + (kBlockHalfBytes >> 4);
                        const uint64_t ah = j < 8 ? act_hi + (uint64_t)(j * 2 * kStep) : emb_hi;
                        const uint64_t al = j < 8 ? act_lo + (uint64_t)(j * 2 * kStep) : emb_lo;
                        umma_f16(d, ah, bh, idesc128, j == 0 ? 0u : 1u);
                        if (!kVar || (terms & 2)) umma_f16(d, ah, bl, idesc128, 1u);
                        if (!kVar || (terms & 4)) umma_f16(d, al, bh, idesc128, 1u);
                        umma_f16(d, ah + kStep, bh + kStep, idesc128, 1u);
                        if (!kVar || (terms & 2)) umma_f16(d, ah + kStep, bl + kStep, idesc128, 1u);
                        if (!kVar || (terms & 4)) umma_f16(d, al + kStep, bh + kStep, idesc128, 1u);
                        umma_commit(bar_empty + 8 * s);
                        if (j == 8) {
                            if (kSave) bulk_wait_read0();
                            umma_commit(bar_dfull + 8);
                        }
                    }
                    __syncwarp();
                }
            }
            // training: this tile's views-layer output (staged in the activation tile by the last epilogue) is stored after the NEXT
            // tile's layer-0 MMAs have been issued (store_hv below), so that layer 0 overlaps the last epilogue; the final tile here
            if (kSave && tile + (int)gridDim.x >= num_tiles) store_hv(rec, (uint32_t)tl);
        }
        if (kSave) { if (elect_one()) bulk_wait0(); __syncwarp(); }
        if (kProf && lane == 0) {
            atomicAdd(&g_prof3[0], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_prof3[1], (unsigned long long)pw_a); atomicAdd(&g_prof3[2], (unsigned long long)pw_e);
            atomicAdd(&g_prof3[3], (unsigned long long)pw_full);
            atomicAdd(&g_prof3[4], (unsigned long long)pw_issue);
        }
    } else {
        // ===== prologue + epilogue warps: thread = (row, p); per k-block of 32 columns it owns columns 8p..8p+7 =====
        const int q = warp & 3, p = warp >> 2;
        const uint32_t row = (uint32_t)(q * 32 + lane);
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t ah = sbase + k3ActHi, al = sbase + k3ActLo, eh = sbase + k3EmbHi, el = sbase + k3EmbLo;
        long long pw_d = 0, p_start = kProf ? clock64() : 0;
        const int dbg = kProf ? g_dbg3 : 0;
        int tl = 0;
        // point encoding of this thread's 16 columns (k-groups 2p, 2p+1); column 63 is the constant 1 that carries the biases
        auto encode = [&](int tile, float* e16) {
            const int gr = tile * (int)kRows + (int)row;
            float x[3] = {0.f, 0.f, 0.f};
            if (gr < n_points) { x[0] = pts[3 * (size_t)gr]; x[1] = pts[3 * (size_t)gr + 1]; x[2] = pts[3 * (size_t)gr + 2]; }
            if (p == 0)      { enc8<0>(x, 63, e16);  enc8<8>(x, 63, e16 + 8); }
            else if (p == 1) { enc8<16>(x, 63, e16); enc8<24>(x, 63, e16 + 8); }
            else if (p == 2) { enc8<32>(x, 63, e16); enc8<40>(x, 63, e16 + 8); }
            else             { enc8<48>(x, 63, e16); enc8<56>(x, 63, e16 + 8); e16[15] = 1.f; }
        };
        float encv[16];
        if ((int)blockIdx.x < num_tiles) encode(blockIdx.x, encv);
        auto publish_encoding = [&]() {       // the (pre-computed) point encoding of the next tile to run -> encoding tile
            emit_kgroup<false>(eh, el, nullptr, 0, row, 2 * (uint32_t)p, encv);
            emit_kgroup<false>(eh, el, nullptr, 0, row, 2 * (uint32_t)p + 1, encv + 8);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_eready);
        };
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            const int grow = tile * (int)kRows + (int)row;
            const bool valid = grow < n_points;
            if (tl == 0) publish_encoding();      // later tiles: published at the end of the previous tile, ahead of its last epilogue
            float alpha_acc = 0.f;
#pragma unroll 1
            for (int layer = 0; layer < 9; ++layer) {
                { PROF_T0(); mbar_wait(bar_dfull + 8 * (layer & 1), (uint32_t)(tl * 5 + (layer >> 1)) & 1); PROF_ADD(pw_d); }
                tc_fence_after();
                if (layer == 5) {
                    // every MMA of layer 5 (the last reader of the point encoding columns 0-31) is done: k-groups 0-3 now take
                    // the direction encoding (column 31 = 1.0 for the bias); published by this layer's a_ready arrivals
                    float dvec[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        int ray = min(grow / n_samples, n_rays - 1);
                        dvec[0] = viewdirs[3 * (size_t)ray]; dvec[1] = viewdirs[3 * (size_t)ray + 1]; dvec[2] = viewdirs[3 * (size_t)ray + 2];
                    }
                    float v[8];
                    if (p == 0)      enc8<0>(dvec, 27, v);
                    else if (p == 1) enc8<8>(dvec, 27, v);
                    else if (p == 2) enc8<16>(dvec, 27, v);
                    else             { enc8<24>(dvec, 27, v); v[7] = 1.f; }
                    emit_kgroup<false>(eh, el, nullptr, 0, row, (uint32_t)p, v);
                }
                const bool relu = layer != 8;
                uint32_t mbits_lo = 0, mbits_hi = 0;      // training: this thread's 8 sign-bit bytes of the layer (k-blocks 0-3 | 4-7)
                const uint32_t dcol = t_lane + (uint32_t)(layer & 1) * 256 + (uint32_t)p * 8;
#pragma unroll 1
                for (uint32_t kb = 0; kb < 8; kb += 2) {
                    float v[16];
                    if (!(dbg & 1)) {
                        tmem_ld8(dcol + kb * 32, v);
                        tmem_ld8(dcol + kb * 32 + 32, v + 8);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = (float)(j + lane) * 0.01f;
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const uint32_t c = (kb + u) * 32 + (uint32_t)p * 8;
                        float* w = v + 8 * u;
                        if (kSave) {              // ReLU sign bits of this k-group (byte kb + u of this thread's eight), from the pre-activations
                            const uint32_t bits = sign_clear_bits8(w);
                            if (kb + u < 4) mbits_lo |= bits << (8 * (kb + u)); else mbits_hi |= bits << (8 * (kb + u - 4));
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float t = relu ? fmaxf(w[j], 0.f) : fmaxf(w[j], -65504.f);
                            w[j] = fminf(t, 65504.f);
                        }
                        if (layer == 7) {
                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c)), a1 = __ldg(reinterpret_cast<const float4*>(misc + kMiscAlphaW + c + 4));
                            alpha_acc = fmaf(w[0], a0.x, alpha_acc); alpha_acc = fmaf(w[1], a0.y, alpha_acc);
                            alpha_acc = fmaf(w[2], a0.z, alpha_acc); alpha_acc = fmaf(w[3], a0.w, alpha_acc);
                            alpha_acc = fmaf(w[4], a1.x, alpha_acc); alpha_acc = fmaf(w[5], a1.y, alpha_acc);
                            alpha_acc = fmaf(w[6], a1.z, alpha_acc); alpha_acc = fmaf(w[7], a1.w, alpha_acc);
                        }
                        if (!(dbg & 2)) emit_kgroup<false>(ah, al, nullptr, 0, row, (kb + u) * 4 + (uint32_t)p, w);
                        if (!(dbg & 4)) fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_aready + 8 * (kb + u));
                    }
                }
                if (kSave && relu)        // record slot M: one 8-byte store per thread and layer (256 contiguous bytes per warp)
                    *reinterpret_cast<uint2*>(acts + (size_t)tile * kTileBytes + kSlotM + (size_t)layer * 4096 + (size_t)p * 1024 + row * 8) =
                        make_uint2(mbits_lo, mbits_hi);
            }
            // the next tile's encoding is computed while the tensor core works on the views layer
            const bool has_next = tile + (int)gridDim.x < num_tiles;
            if (has_next) encode(tile + (int)gridDim.x, encv);
            {   // views layer: ReLU (bias already in the accumulator), then rgb_linear as an fp32 dot product; 32 of 128 columns per thread
                { PROF_T0(); mbar_wait(bar_dfull + 8, (uint32_t)(tl * 5 + 4) & 1); PROF_ADD(pw_d); }
                tc_fence_after();
                // Every reader of the encoding tile (the views-layer MMAs were the last) is done: hand the NEXT tile's encoding to the
                // MMA warp first, so that its layer 0 runs while these warps finish this tile (the views accumulator D[1] is not
                // written again before layer 1, which waits for these same warps).
                if (has_next) publish_encoding();
                // scratch for the head reductions: the upper half of the activation tile's hi part (free since the views-layer MMAs;
                // training stages hv in k-groups 0-15 only; the next writer is the next tile's first epilogue, after the barrier below)
                float* s_rgb = reinterpret_cast<float*>(smem + k3ActHi + 32768);
                float* s_alpha = s_rgb + 1152;
                const uint32_t c = (uint32_t)p * 32;
                float v[32];
                tmem_ld32(t_lane + 256 + c, v);
                tmem_ld_wait();
                if (kSave) {      // ReLU sign bits of the views-layer pre-activations: k-groups 4p..4p+3 of this row, one 4-byte store
                    uint32_t word = 0;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) word |= sign_clear_bits8(v + 8 * k4) << (8 * k4);
                    *reinterpret_cast<uint32_t*>(acts + (size_t)tile * kTileBytes + kSlotM + 32768 + row * 16 + (c >> 3)) = word;
                }
                float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float hv = fmaxf(v[j], 0.f);
                    r0 = fmaf(hv, __ldg(misc + kMiscRgbW + c + j), r0);
                    r1 = fmaf(hv, __ldg(misc + kMiscRgbW + 128 + c + j), r1);
                    r2 = fmaf(hv, __ldg(misc + kMiscRgbW + 256 + c + j), r2);
                    v[j] = fminf(hv, 65504.f);
                }

                if (kSave) {      // stage hv in the activation tile (its last reader, this layer's MMAs, is done)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) emit_kgroup<false>(ah, al, nullptr, 0, row, (c >> 3) + k4, v + 8 * k4);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_hv);
                }
                if (p > 0) { float* o = s_rgb + (p - 1) * 384; o[row] = r0; o[128 + row] = r1; o[256 + row] = r2; }
                s_alpha[p * 128 + row] = alpha_acc;
                named_bar_sync(1, 512);
                if (p == 0 && valid) {
                    float4 o;
                    o.x = r0 + s_rgb[row] + s_rgb[384 + row] + s_rgb[768 + row] + __ldg(misc + kMiscRgbB);
                    o.y = r1 + s_rgb[128 + row] + s_rgb[512 + row] + s_rgb[896 + row] + __ldg(misc + kMiscRgbB + 1);
                    o.z = r2 + s_rgb[256 + row] + s_rgb[640 + row] + s_rgb[1024 + row] + __ldg(misc + kMiscRgbB + 2);
                    o.w = s_alpha[row] + s_alpha[128 + row] + s_alpha[256 + row] + s_alpha[384 + row] + __ldg(misc + kMiscAlphaB);
                    *reinterpret_cast<float4*>(raw + 4 * (size_t)grow) = o;
                }
                tc_fence_before();
                named_bar_sync(1, 512);      // the scratch lives in the activation tile the next tile's first epilogue rewrites
            }
        }
        if (kProf && lane == 0 && warp == 0) {
            atomicAdd(&g_prof3[8], (unsigned long long)(clock64() - p_start));
            atomicAdd(&g_prof3[9], (unsigned long long)pw_d);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem, 512);
}

static int g_prof3_host = 0;        // cnerf_debug_profile3: launch the instrumented instantiation

int pack_stream3(const RawParams& p, uint8_t* stream3, cudaStream_t st) {
    pack_weights3_kernel<<<dim3(k3NumBlocks, kWeightReplicas), 256, 0, st>>>(p, stream3);
    CNERF_LAUNCH_CHECK("pack_weights3_kernel");
    return CNERF_OK;
}
int stream3_blocks() { return k3NumBlocks; }

template <int kSave, bool kProf, bool kVar>
static int launch_fused3_inst(int grid, const uint8_t* stream3, const float* misc, const float* pts, const float* viewdirs,
                              int n_points, int n_samples, int n_rays, float* raw, uint8_t* acts, int terms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused3_kernel<kSave, kProf, kVar>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3Smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(mlp_fused3_kernel)");
        attr_set = true;
    }
    mlp_fused3_kernel<kSave, kProf, kVar><<<grid, k3Threads, k3Smem, st>>>(stream3, misc, pts, viewdirs, n_points, n_samples, n_rays, raw, acts, terms);
    CNERF_LAUNCH_CHECK("mlp_fused3_kernel");
    return CNERF_OK;
}

// record_lo: the training record also carries the lo halves (three-term weight gradients); terms: 7 = production, anything else
// selects the measurement instantiation (inference only)
int launch_fused3(const uint8_t* stream3, const float* misc, const float* pts, const float* viewdirs, int n_points,
                  int n_samples, int n_rays, float* raw, uint8_t* acts, int record_lo, int terms, cudaStream_t st) {
    const int tiles = ceil_div(n_points, (int)kRows);
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
#define CNERF_F3(S, P, V) launch_fused3_inst<S, P, V>(grid, stream3, misc, pts, viewdirs, n_points, n_samples, n_rays, raw, acts, terms, st)
    if (terms != 7) {
        CNERF_REQUIRE(!acts, "mlp_fused3: the partial-product variants are inference only");
        return CNERF_F3(0, false, true);
    }
    if (g_prof3_host) return acts ? (record_lo ? CNERF_F3(2, true, false) : CNERF_F3(1, true, false)) : CNERF_F3(0, true, false);
    return acts ? (record_lo ? CNERF_F3(2, false, false) : CNERF_F3(1, false, false)) : CNERF_F3(0, false, false);
#undef CNERF_F3
}

}  // namespace cnerf

// Debug: in-kernel phase profile of mlp_fused3_kernel (cycles summed over CTAs):
//  [0] MMA warp total  [1] wait A k-blocks  [2] wait encoding  [3] wait weights  [4] MMA issue + commit  [8] epilogue total  [9] wait D
// enable: bit 0 = profile on, bits 1.. = g_dbg3 timing experiments
extern "C" int cnerf_debug_profile3(int enable, unsigned long long* out16) {
    using namespace cnerf;
    unsigned long long zero[16] = {0};
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out16) e = cudaMemcpyFromSymbol(out16, g_prof3, sizeof(zero));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_prof3, zero, sizeof(zero));
    const int on = enable & 1, dbg = enable >> 1;
    g_prof3_host = on;
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_prof3_on, &on, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_dbg3, &dbg, sizeof(int));
    if (e != cudaSuccess) return check_cuda(e, "cnerf_debug_profile3");
    return CNERF_OK;
}
