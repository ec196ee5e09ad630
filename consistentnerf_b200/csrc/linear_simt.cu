// K2 (stand-alone positional encoding) and the generic fp32 layer kernels (CUDA cores).
//
// This is the any-architecture path behind the NeRF nn.Module surface (arbitrary D / W / skips /
// use_viewdirs, NP/run_nerf_helpers.py:67-130) and the fp32 backward of the MLP.  It is a plain
// 128x128x16 register-tiled SGEMM with fused bias / ReLU / ReLU-mask prologue; the canonical
// 8x256 network takes the tcgen05 path (mlp_fwd3.cu / mlp_bwd_tc.cu) instead.
#include "common.cuh"

namespace cnerf {

// ------------------------------------------------------------------------------------
// positional encoding: one thread per OUTPUT element so stores are coalesced.
// column j of the encoding of a C-vector: j < C -> x[j]; else block b=(j-C)/C: octave b/2,
// sin for even b, cos for odd b (NP/run_nerf_helpers.py:24-46).
// ------------------------------------------------------------------------------------
__global__ void posenc_kernel(const float* __restrict__ x, int ldx, int n, int C, int L, int repeat,
                              float* __restrict__ out, int ldo, int col0) {
    int E = C * (1 + 2 * L);
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * E) return;
    int row = (int)(idx / E), j = (int)(idx - (int64_t)row * E);
    int src = row / repeat;
    float v;
    if (j < C) {
        v = x[(size_t)src * ldx + j];
    } else {
        int b = (j - C) / C, c = (j - C) - b * C;
        float arg = x[(size_t)src * ldx + c] * exp2f((float)(b >> 1));   // exact power-of-two scaling
        v = (b & 1) ? cosf(arg) : sinf(arg);
    }
    out[(size_t)row * ldo + col0 + j] = v;
}

// ------------------------------------------------------------------------------------
// SGEMM  C[M,N] = sum_k A(m,k) * B(n,k)
// ------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct GemmArgs {
    const float* A; int lda;
    const float* Amask; int ldam;      // optional: A(m,k) is zeroed where Amask(m,k) <= 0 (ReLU backward)
    const float* B; int ldb;
    float* C; int ldc;
    int M, N, K;
    const float* bias; int relu; int accumulate;
    int k_chunk;                       // split-K: reduction range per blockIdx.z (0 = whole K)
    float* partial;                    // split-K partial tiles [splits][M][N]
};

template <bool KCONTIG>
__device__ __forceinline__ void load_tile(const float* __restrict__ base, int ld, const float* __restrict__ msk,
                                          int ldm, int row0, int nrows, int k0, int kend, float (*dst)[BM + PAD], int t) {
    if (KCONTIG) {
        int kk = t & 15, r0 = t >> 4;
        int k = k0 + kk;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = r0 + 16 * i, gr = row0 + r;
            float v = 0.f;
            if (gr < nrows && k < kend) {
                v = base[(size_t)gr * ld + k];
                if (msk && !(msk[(size_t)gr * ldm + k] > 0.f)) v = 0.f;
            }
            dst[kk][r] = v;
        }
    } else {
        int r = t & 127, kk0 = t >> 7;
        int gr = row0 + r;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int kk = kk0 + 2 * i, k = k0 + kk;
            float v = 0.f;
            if (gr < nrows && k < kend) {
                v = base[(size_t)k * ld + gr];
                if (msk && !(msk[(size_t)k * ldm + gr] > 0.f)) v = 0.f;
            }
            dst[kk][r] = v;
        }
    }
}

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(256)
sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int kbeg = 0, kend = g.K;
    if (g.k_chunk > 0) { kbeg = blockIdx.z * g.k_chunk; kend = min(g.K, kbeg + g.k_chunk); }
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        load_tile<A_KCONTIG>(g.A, g.lda, g.Amask, g.ldam, m0, g.M, k0, kend, As, t);
        load_tile<B_KCONTIG>(g.B, g.ldb, nullptr, 0, n0, g.N, k0, kend, Bs, t);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
            *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    if (g.k_chunk > 0) {
        float* P = g.partial + (size_t)blockIdx.z * g.M * g.N;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int m = m0 + ty * 8 + i;
            if (m >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int n = n0 + tx * 8 + j;
                if (n < g.N) P[(size_t)m * g.N + n] = acc[i][j];
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + tx * 8 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[n];
            if (g.relu) v = fmaxf(v, 0.f);
            float* c = g.C + (size_t)m * g.ldc + n;
            *c = g.accumulate ? *c + v : v;
        }
    }
}

// fold split-K partials in split order (deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t mn, float* __restrict__ out,
                                     int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mn) return;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += partial[(size_t)s * mn + i];
    out[i] = accumulate ? out[i] + v : v;
}

// db partials: block (x: 32 columns, y: row split); thread (tx column, ty row lane)
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy, int m, int n,
                      int rows_per_split, float* __restrict__ partial) {
    __shared__ float sh[8][33];
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int col = blockIdx.x * 32 + tx;
    int r0 = blockIdx.y * rows_per_split, r1 = min(m, r0 + rows_per_split);
    float s = 0.f;
    if (col < n) {
        for (int r = r0 + ty; r < r1; r += 8) {
            float v = dy[(size_t)r * lddy + col];
            if (y && !(y[(size_t)r * ldy + col] > 0.f)) v = 0.f;
            s += v;
        }
    }
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && col < n) {
        float tot = 0.f;
        for (int k = 0; k < 8; ++k) tot += sh[k][tx];
        partial[(size_t)blockIdx.y * n + col] = tot;
    }
}

static int bwd_weight_splits(int m) {
    int s = ceil_div(m, 2048);
    if (s < 1) s = 1;
    if (s > 296) s = 296;
    return s;
}

}  // namespace cnerf

using namespace cnerf;

extern "C" int cnerf_posenc(const float* x, int ldx, int n, int C, int n_freqs, int repeat, float* out, int ldo,
                            int col0, void* stream) {
    CNERF_REQUIRE(x && out, "cnerf_posenc: null pointer");
    CNERF_REQUIRE(n >= 0 && C >= 1 && n_freqs >= 0 && n_freqs <= 24 && repeat >= 1 && ldx >= C, "cnerf_posenc: bad sizes");
    CNERF_REQUIRE(ldo >= col0 + C * (1 + 2 * n_freqs), "cnerf_posenc: output row too short");
    if (n == 0) return CNERF_OK;
    int64_t total = (int64_t)n * C * (1 + 2 * n_freqs);
    posenc_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(x, ldx, n, C, n_freqs, repeat, out, ldo, col0);
    CNERF_LAUNCH_CHECK("posenc_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_linear_fwd(const float* x, int ldx, const float* w, const float* b, int m, int n, int k, int relu,
                                float* y, int ldy, void* stream) {
    CNERF_REQUIRE(x && w && y, "cnerf_linear_fwd: null pointer");
    CNERF_REQUIRE(m >= 0 && n >= 1 && k >= 1 && ldx >= k && ldy >= n, "cnerf_linear_fwd: bad sizes m=%d n=%d k=%d", m, n, k);
    if (m == 0) return CNERF_OK;
    GemmArgs g = {x, ldx, nullptr, 0, w, k, y, ldy, m, n, k, b, relu, 0, 0, nullptr};
    dim3 grid(ceil_div(n, BN), ceil_div(m, BM), 1);
    sgemm_kernel<true, true><<<grid, 256, 0, as_stream(stream)>>>(g);
    CNERF_LAUNCH_CHECK("sgemm_kernel<fwd>");
    return CNERF_OK;
}

extern "C" int cnerf_linear_bwd_data(const float* dy, int lddy, const float* y, int ldy, const float* w, int m, int n,
                                     int k, float* dx, int lddx, int accumulate, void* stream) {
    CNERF_REQUIRE(dy && w && dx, "cnerf_linear_bwd_data: null pointer");
    CNERF_REQUIRE(m >= 0 && n >= 1 && k >= 1 && lddy >= n && lddx >= k, "cnerf_linear_bwd_data: bad sizes");
    if (m == 0) return CNERF_OK;
    // dx[m,k] = sum_n dy[m,n] w[n,k]:  M=m, N=k, K=n;  B(k', n) = w[n*k + k'] is N-contiguous
    GemmArgs g = {dy, lddy, y, ldy, w, k, dx, lddx, m, k, n, nullptr, 0, accumulate, 0, nullptr};
    dim3 grid(ceil_div(k, BN), ceil_div(m, BM), 1);
    sgemm_kernel<true, false><<<grid, 256, 0, as_stream(stream)>>>(g);
    CNERF_LAUNCH_CHECK("sgemm_kernel<bwd_data>");
    return CNERF_OK;
}

extern "C" int64_t cnerf_linear_bwd_weight_workspace(int m, int n, int k) {
    int s = bwd_weight_splits(m);
    return (int64_t)s * ((int64_t)n * k + n) * (int64_t)sizeof(float);
}

extern "C" int cnerf_linear_bwd_weight(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx,
                                       int m, int n, int k, float* dw, float* db, int accumulate, void* workspace,
                                       void* stream) {
    CNERF_REQUIRE(dy && x && dw && workspace, "cnerf_linear_bwd_weight: null pointer");
    CNERF_REQUIRE(m >= 1 && n >= 1 && k >= 1 && lddy >= n && ldx >= k, "cnerf_linear_bwd_weight: bad sizes");
    int splits = bwd_weight_splits(m);
    int chunk = ceil_div(ceil_div(m, splits), BK) * BK;
    splits = ceil_div(m, chunk);
    float* part = reinterpret_cast<float*>(workspace);
    // dw[n,k] = sum_m dy[m,n] x[m,k]:  M=n, N=k, K=m; both operands are output-contiguous
    GemmArgs g = {dy, lddy, y, ldy, x, ldx, nullptr, 0, n, k, m, nullptr, 0, 0, chunk, part};
    dim3 grid(ceil_div(k, BN), ceil_div(n, BM), splits);
    sgemm_kernel<false, false><<<grid, 256, 0, as_stream(stream)>>>(g);
    CNERF_LAUNCH_CHECK("sgemm_kernel<bwd_weight>");
    int64_t mn = (int64_t)n * k;
    splitk_reduce_kernel<<<(unsigned)ceil_div64(mn, 256), 256, 0, as_stream(stream)>>>(part, splits, mn, dw, accumulate);
    CNERF_LAUNCH_CHECK("splitk_reduce_kernel");
    if (db) {
        float* bpart = part + (size_t)splits * mn;
        dim3 g2(ceil_div(n, 32), splits);
        colsum_partial_kernel<<<g2, 256, 0, as_stream(stream)>>>(dy, lddy, y, ldy, m, n, chunk, bpart);
        CNERF_LAUNCH_CHECK("colsum_partial_kernel");
        splitk_reduce_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(bpart, splits, n, db, accumulate);
        CNERF_LAUNCH_CHECK("splitk_reduce_kernel(db)");
    }
    return CNERF_OK;
}
