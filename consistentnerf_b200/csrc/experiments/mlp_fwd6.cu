// K2+K3 fused forward, fp16 operands, CTA PAIRS (tcgen05 cta_group::2, M = 256) with two super-tiles in flight per pair.
//
// Why: the single-CTA two-tile kernel (mlp_fwd5.cu) uses every weight unit for exactly one M = 128 MMA, so per
// 128-cycle MMA an SM moves 8 KB of bulk-copy writes + 8 KB of B-operand reads + 4 KB of A reads + 4 KB of epilogue stores = 192 B/clk
// through 128 B/clk of shared-memory bandwidth: two thirds of the MMA rate is its ceiling, and the bulk copies starve.  In a CTA pair one
// M = 256 instruction drives the tensor cores of both SMs of a TPC; each CTA holds its own 128 rows of A and of D and only HALF of B (N/2
// weight rows).  Per SM and MMA that is 4 KB written + 4 KB of B read + 4 KB of A + 4 KB of epilogue stores = 128 B/clk -- feasible at the
// full rate -- and the L2 traffic of the weight stream halves as well.
//
//   cluster   2 CTAs; a super-tile is 256 consecutive points: CTA r works on tile 2 S + r (128 points)
//   slots     two super-tiles in flight (ping-pong, as in mlp_fwd5.cu): while the epilogue warps of BOTH CTAs turn slot X's accumulators
//             into the next A operands, the pair's tensor cores run slot Y's layer
//   MMA       issued by the leader CTA (rank 0) only, M = 256, N = 256 (views layer: 128), K = 16; tcgen05.commit multicasts to both CTAs
//   SMEM      per CTA and slot: A operand (64 KB) + encoding tile (16 KB); 16-stage ring of 4 KB weight half-units (this CTA's N/2 rows)
//   TMEM      per CTA: one 256-column accumulator per slot (lane = local row)
//   sync      full[st] (local bulk copy landed), pfull[st] (leader: the peer's half landed; forwarded by the peer's warp 17),
//             empty[st] / d_full[slot] (multicast commits), a_ready[slot] (leader: 32 warp arrivals, 16 from each CTA)
// Inference only; training runs mlp_fwd5.cu.
//
// STATUS: written at the end of round 2, compiles for sm_100a (SASS: UTCHMMA.2CTA), but NOT RUN ON A GPU -- the round's GPU budget was
// spent before it could be tested (scripts/test_pair.py is the parity + timing check to run first).  It is therefore built only with
// `python -m consistentnerf_b200.build --experiments` and selected only with CNERF_FWD_PAIR=1; nothing in the product uses it.
#include "mlp_blocks.cuh"

namespace cnerf {

constexpr int k6Threads = 576;                            // 16 epilogue warps + loader warp + MMA (leader) / forwarder (peer) warp
constexpr uint32_t k6Act = 0;                             // + slot * 65536
constexpr uint32_t k6Emb = 131072;                        // + slot * 16384
constexpr uint32_t k6Ring = 163840;
constexpr int k6Stages = 16;
constexpr uint32_t k6Half = 4096;                         // this CTA's half of a weight unit
constexpr uint32_t k6Bars = k6Ring + k6Stages * k6Half;   // 229376
constexpr uint32_t k6BarFull = k6Bars, k6BarEmpty = k6Bars + 128, k6BarPFull = k6Bars + 256, k6BarDFull = k6Bars + 384,
                   k6BarAReady = k6Bars + 400, k6TmemSlot = k6Bars + 416;
constexpr uint32_t k6Smem = k6Bars + 448;

__device__ __forceinline__ int k6_layer_first(int l) { return l == 0 ? 0 : l <= 5 ? 4 + 17 * (l - 1) : l <= 9 ? 92 + 17 * (l - 6) : 152; }

// stream6: unit b of the forward program (mlp_blocks.cuh), hi halves only, split by output rows: CTA r's half at (2 b + r) * 4 KB
//   layers 0-8: local rows nl = n - 128 r, element (nl, k) at (k/8)*2048 + nl*16 + (k%8)*2
//   layer 9   : local rows nl = n - 64 r,  element (nl, k) at (k/8)*1024 + nl*16 + (k%8)*2
__global__ void __launch_bounds__(256)
pack_weights6_kernel(RawParams p, uint8_t* __restrict__ stream) {
    const int b = blockIdx.x >> 1, r = blockIdx.x & 1;
    const Blk3 bi = block3_info(b);
    const float* W = p.w[bi.layer];
    const int ld = p.ld[bi.layer];
    uint8_t* dst = stream + (size_t)blockIdx.x * k6Half;
    const int rows = bi.layer == 9 ? 64 : 128, kgs = bi.layer == 9 ? 4 : 2;
    for (int u = threadIdx.x; u < rows * kgs; u += 256) {
        const int nl = u % rows, kg = u / rows, n = r * rows + nl;
        uint32_t h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k = kg * 8 + 2 * e + q;
                v[q] = (k < bi.kvalid) ? W[(size_t)n * ld + bi.src_k0 + k] : (k == bi.bias_k ? p.b[bi.layer][n] : 0.f);
            }
            __half2 t = __floats2half2_rn(v[0], v[1]);
            h[e] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(dst + (size_t)kg * rows * 16 + (size_t)nl * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

__device__ __forceinline__ void emit_hi6(uint32_t base, uint32_t row, uint32_t kg, const float* v) {
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __half2 t = __floats2half2_rn(v[2 * i], v[2 * i + 1]); h[i] = *reinterpret_cast<uint32_t*>(&t); }
    st_shared_v4(base + kg * kLBO + row * 16, h[0], h[1], h[2], h[3]);
}

__global__ void __launch_bounds__(k6Threads, 1)
mlp_fused6_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ misc, const float* __restrict__ pts,
                  const float* __restrict__ viewdirs, int n_points, int n_samples, int n_rays, float* __restrict__ raw) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t bar_full = sbase + k6BarFull, bar_empty = sbase + k6BarEmpty, bar_pfull = sbase + k6BarPFull;
    const uint32_t bar_dfull = sbase + k6BarDFull, bar_aready = sbase + k6BarAReady;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + k6TmemSlot);
    const int num_tiles = (n_points + (int)kRows - 1) / (int)kRows;
    const int num_super = (num_tiles + 1) / 2;
    // super-tile n of this pair = pair + n * npairs; slot s works on n = 2 it + s
    const int my_super = (pair < num_super) ? (num_super - 1 - pair) / npairs + 1 : 0;
    const int n_iter = (my_super + 1) / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < k6Stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_pfull + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_aready + 8 * s, 32); }
        fence_barrier_init();
    }
    if (warp == 17) tmem_alloc2(sbase + k6TmemSlot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 16) {
        // ===== weight loader: this CTA's half of every unit, in the order the leader consumes them =====
        if (lane == 0) {
            const uint64_t keep = l2_policy_evict_last();
            uint32_t st = 0, ph = 0;
            for (int it = 0; it < n_iter; ++it)
                for (int layer = 0; layer < 10; ++layer)
                    for (int s = 0; s < 2; ++s) {
                        if (2 * it + s >= my_super) continue;
                        const uint8_t* src = wstream + ((size_t)k6_layer_first(layer) * 2 + rank) * k6Half;
                        for (int nb = k6_layer_first(layer + 1) - k6_layer_first(layer); nb > 0; --nb, src += 2 * k6Half) {
                            mbar_wait(bar_empty + 8 * st, ph ^ 1);
                            mbar_arrive_expect_tx(bar_full + 8 * st, k6Half);
                            bulk_g2s_hint(sbase + k6Ring + st * k6Half, src, k6Half, bar_full + 8 * st, keep);
                            st = (st + 1) & (k6Stages - 1);
                            ph ^= (st == 0);
                        }
                    }
        }
    } else if (warp == 17 && rank != 0) {
        // ===== peer: tell the leader when this CTA's half of a unit has landed =====
        if (lane == 0) {
            const uint32_t pf0 = mapa_u32(bar_pfull, 0);
            uint32_t st = 0, ph = 0;
            for (int it = 0; it < n_iter; ++it)
                for (int layer = 0; layer < 10; ++layer)
                    for (int s = 0; s < 2; ++s) {
                        if (2 * it + s >= my_super) continue;
                        for (int nb = k6_layer_first(layer + 1) - k6_layer_first(layer); nb > 0; --nb) {
                            mbar_wait(bar_full + 8 * st, ph);
                            mbar_arrive_cluster(pf0 + 8 * st);
                            st = (st + 1) & (k6Stages - 1);
                            ph ^= (st == 0);
                        }
                    }
        }
    } else if (warp == 17) {
        // ===== leader: MMA issuer for the pair, alternating between the two slots layer by layer =====
        constexpr uint32_t idesc256 = instr_desc(256, 256), idesc128 = instr_desc(256, 128);
        constexpr uint32_t kStep = 2 * (kLBO >> 4);                               // two k-groups = one K=16 step of an A tile
        constexpr uint32_t bStep9 = 2048 >> 4;                                    // views layer: two k-groups of a 64-row half-unit
        uint32_t st = 0, ph = 0;
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll 1
            for (int layer = 0; layer < 10; ++layer) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    if (2 * it + s >= my_super) continue;
                    const uint64_t act_d = smem_desc(sbase + k6Act + (uint32_t)s * 65536), emb_d = smem_desc(sbase + k6Emb + (uint32_t)s * 16384);
                    const uint32_t d = tmem + (uint32_t)s * 256;
                    mbar_wait_cluster(bar_aready + 8 * s, (uint32_t)(it * 10 + layer) & 1);
                    tc_fence_after();
                    uint32_t acc = 0u;
                    auto segment = [&](uint64_t a, int count, uint32_t a_step) {
#pragma unroll 1
                        for (int j = 0; j < count; ++j, a += a_step) {
                            mbar_wait(bar_full + 8 * st, ph);
                            mbar_wait_cluster(bar_pfull + 8 * st, ph);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t ring = sbase + k6Ring + st * k6Half;
                                if (layer < 9) {
                                    umma2_f16(d, a, smem_desc(ring), idesc256, acc);
                                } else {                                              // views layer: [64 local rows x 32 k] half-units, N = 128
                                    const uint64_t b = smem_desc_any(ring, 1024, kSBO);
                                    umma2_f16(d, a, b, idesc128, acc);
                                    umma2_f16(d, a + kStep, b + bStep9, idesc128, 1u);
                                }
                                umma2_commit(bar_empty + 8 * st, 3);
                            }
                            __syncwarp();
                            acc = 1u;
                            st = (st + 1) & (k6Stages - 1);
                            ph ^= (st == 0);
                        }
                    };
                    if (layer == 0) segment(emb_d, 4, kStep);
                    else if (layer == 5) { segment(emb_d, 4, kStep); segment(act_d, 16, kStep); }
                    else if (layer < 9) { segment(act_d, 16, kStep); segment(emb_d + 3 * kStep, 1, 0); }      // bias unit: encoding columns 48-63
                    else { segment(act_d, 8, 2 * kStep); segment(emb_d, 1, 0); }                              // feature columns, then the direction encoding
                    if (elect_one()) umma2_commit(bar_dfull + 8 * s, 3);
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== prologue + epilogue warps of THIS CTA's tiles: thread = (row, p); both slots in turn (as mlp_fwd5.cu) =====
        const int q = warp & 3, p = warp >> 2;
        const uint32_t row = (uint32_t)(q * 32 + lane);
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t ar0 = mapa_u32(bar_aready, 0);                       // the leader's a_ready barriers
        auto arrive_ready = [&](int s) {
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(ar0 + 8 * s);
        };
        auto tile_of = [&](int n) { return 2 * (pair + n * npairs) + (int)rank; };
        auto publish_encoding = [&](int tile, int s) {
            const int gr = tile * (int)kRows + (int)row;
            float x[3] = {0.f, 0.f, 0.f}, e16[16];
            if (gr < n_points) { x[0] = pts[3 * (size_t)gr]; x[1] = pts[3 * (size_t)gr + 1]; x[2] = pts[3 * (size_t)gr + 2]; }
            if (p == 0)      { enc8<0>(x, 63, e16);  enc8<8>(x, 63, e16 + 8); }
            else if (p == 1) { enc8<16>(x, 63, e16); enc8<24>(x, 63, e16 + 8); }
            else if (p == 2) { enc8<32>(x, 63, e16); enc8<40>(x, 63, e16 + 8); }
            else             { enc8<48>(x, 63, e16); enc8<56>(x, 63, e16 + 8); e16[15] = 1.f; }
            const uint32_t eb = sbase + k6Emb + (uint32_t)s * 16384;
            emit_hi6(eb, row, 2 * (uint32_t)p, e16);
            emit_hi6(eb, row, 2 * (uint32_t)p + 1, e16 + 8);
            arrive_ready(s);
        };
        for (int s = 0; s < 2; ++s)
            if (s < my_super) publish_encoding(tile_of(s), s);
        float alpha0 = 0.f, alpha1 = 0.f;
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll 1
            for (int layer = 0; layer < 10; ++layer) {
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const int n = 2 * it + s;
                    if (n >= my_super) continue;
                    const int tile = tile_of(n);
                    const int grow = tile * (int)kRows + (int)row;
                    const bool valid = grow < n_points;
                    const uint32_t ab = sbase + k6Act + (uint32_t)s * 65536, eb = sbase + k6Emb + (uint32_t)s * 16384;
                    mbar_wait(bar_dfull + 8 * s, (uint32_t)(it * 10 + layer) & 1);
                    tc_fence_after();
                    if (layer < 9) {
                        if (layer == 0) { if (s) alpha1 = 0.f; else alpha0 = 0.f; }
                        if (layer == 5) {
                            float dvec[3] = {0.f, 0.f, 0.f}, v[8];
                            if (valid) {
                                const int ray = min(grow / n_samples, n_rays - 1);
                                dvec[0] = viewdirs[3 * (size_t)ray]; dvec[1] = viewdirs[3 * (size_t)ray + 1]; dvec[2] = viewdirs[3 * (size_t)ray + 2];
                            }
                            if (p == 0)      enc8<0>(dvec, 27, v);
                            else if (p == 1) enc8<8>(dvec, 27, v);
                            else if (p == 2) enc8<16>(dvec, 27, v);
                            else             { enc8<24>(dvec, 27, v); v[7] = 1.f; }
                            emit_hi6(eb, row, (uint32_t)p, v);
                        }
                        const bool relu = layer != 8;
                        const uint32_t dcol = t_lane + (uint32_t)s * 256 + (uint32_t)p * 8;
#pragma unroll
                        for (uint32_t half = 0; half < 2; ++half) {
                            float v[32];
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4) tmem_ld8(dcol + (half * 4 + k4) * 32, v + 8 * k4);
                            tmem_ld_wait();
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4) {
                                const uint32_t kb = half * 4 + k4;
                                float* w = v + 8 * k4;
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
                                emit_hi6(ab, row, kb * 4 + (uint32_t)p, w);
                            }
                        }
                        arrive_ready(s);
                    } else {
                        // views layer: ReLU, rgb_linear as an fp32 dot product; 32 of 128 columns per thread
                        const uint32_t c = (uint32_t)p * 32;
                        float v[32];
                        tmem_ld32(t_lane + (uint32_t)s * 256 + c, v);
                        tmem_ld_wait();
                        tc_fence_before();
                        // accumulator in registers, every reader of the slot's encoding tile done: hand the slot's NEXT super-tile to the leader
                        if (n + 2 < my_super) publish_encoding(tile_of(n + 2), s);
                        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float hv = fmaxf(v[j], 0.f);
                            r0 = fmaf(hv, __ldg(misc + kMiscRgbW + c + j), r0);
                            r1 = fmaf(hv, __ldg(misc + kMiscRgbW + 128 + c + j), r1);
                            r2 = fmaf(hv, __ldg(misc + kMiscRgbW + 256 + c + j), r2);
                        }
                        // scratch: the upper half of the slot's A tile (its last readers, this layer's MMAs, are done; the next writer is the
                        // slot's next epilogue 0, after the barrier below)
                        float* s_rgb = reinterpret_cast<float*>(smem + k6Act + (size_t)s * 65536 + 32768);
                        float* s_alpha = s_rgb + 1152;
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
                        named_bar_sync(1, 512);
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 17) tmem_dealloc2(tmem, 512);
}

int pack_stream6(const RawParams& p, uint8_t* stream6, cudaStream_t st) {
    pack_weights6_kernel<<<2 * k3NumBlocks, 256, 0, st>>>(p, stream6);
    CNERF_LAUNCH_CHECK("pack_weights6_kernel");
    return CNERF_OK;
}
size_t stream6_bytes() { return (size_t)2 * k3NumBlocks * k6Half; }

int launch_fused6(const uint8_t* stream6, const float* misc, const float* pts, const float* viewdirs, int n_points,
                  int n_samples, int n_rays, float* raw, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused6_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k6Smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(mlp_fused6_kernel)");
        attr_set = true;
    }
    const int num_super = ceil_div(ceil_div(n_points, (int)kRows), 2);
    const int pairs = num_super < kNumSMs / 2 ? num_super : kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(k6Threads); cfg.dynamicSmemBytes = k6Smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_fused6_kernel, stream6, misc, pts, viewdirs, n_points, n_samples, n_rays, raw);
    if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(mlp_fused6_kernel)");
    return CNERF_OK;
}

}  // namespace cnerf
