// Ray preparation, stratified sampling (K1) and hierarchical sampling (K5).
//
// These are HBM-bound element/row kernels.  Where the reference's result feeds an
// integer decision (the inverse-CDF bin search) the arithmetic is written with explicit
// round-to-nearest intrinsics in the reference's operation order so the compiler cannot
// contract it into FMAs: identical (cdf, u) give bit-identical indices and samples.
#include "common.cuh"

namespace cnerf {

// ------------------------------------------------------------------------------------
// ray packing  (render(): NP/run_nerf.py:101-126; ndc_rays: NP/run_nerf_helpers.py:186-202)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void ndc_transform(float H, float W, float focal, float near_, float* o, float* d) {
    float t = -(near_ + o[2]) / d[2];
    float ox = o[0] + t * d[0], oy = o[1] + t * d[1], oz = o[2] + t * d[2];
    float sx = -1.f / (W / (2.f * focal)), sy = -1.f / (H / (2.f * focal));
    float o0 = sx * ox / oz, o1 = sy * oy / oz, o2 = 1.f + 2.f * near_ / oz;
    float d0 = sx * (d[0] / d[2] - ox / oz), d1 = sy * (d[1] / d[2] - oy / oz), d2 = -2.f * near_ / oz;
    o[0] = o0; o[1] = o1; o[2] = o2;
    d[0] = d0; d[1] = d1; d[2] = d2;
}

__device__ __forceinline__ void write_ray(float* row, const float* o, const float* d, float near_, float far_,
                                          bool viewdirs, const float* vd) {
    row[0] = o[0]; row[1] = o[1]; row[2] = o[2];
    row[3] = d[0]; row[4] = d[1]; row[5] = d[2];
    row[6] = near_; row[7] = far_;
    if (viewdirs) { row[8] = vd[0]; row[9] = vd[1]; row[10] = vd[2]; }
}

__device__ __forceinline__ void unit_dir(const float* d, float* vd) {
    float n2 = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
    float n = __fsqrt_rn(n2);
    vd[0] = __fdiv_rn(d[0], n); vd[1] = __fdiv_rn(d[1], n); vd[2] = __fdiv_rn(d[2], n);
}

__global__ void pack_rays_kernel(const float* __restrict__ ro, const float* __restrict__ rd, int n, float near_,
                                 float far_, int use_viewdirs, int ndc, float H, float W, float focal,
                                 float* __restrict__ rays) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o[3] = {ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]};
    float d[3] = {rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]};
    float vd[3] = {0.f, 0.f, 0.f};
    if (use_viewdirs) unit_dir(d, vd);
    if (ndc) ndc_transform(H, W, focal, 1.f, o, d);
    write_ray(rays + (size_t)i * (use_viewdirs ? 11 : 8), o, d, near_, far_, use_viewdirs, vd);
}

__global__ void image_rays_kernel(int H, int W, Mat3 K, Mat34 c2w, float near_, float far_, int use_viewdirs,
                                  int ndc, float* __restrict__ rays) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    int py = i / W, px = i - py * W;
    // get_rays: dirs = ((i-cx)/fx, -(j-cy)/fy, -1), rays_d = sum(dirs * c2w[:3,:3], -1)
    float dx = __fdiv_rn(__fsub_rn((float)px, K.m[2]), K.m[0]);
    float dy = -__fdiv_rn(__fsub_rn((float)py, K.m[5]), K.m[4]);
    float dz = -1.f;
    float d[3], o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w.m[4 * r]), __fmul_rn(dy, c2w.m[4 * r + 1])),
                         __fmul_rn(dz, c2w.m[4 * r + 2]));
        o[r] = c2w.m[4 * r + 3];
    }
    float vd[3] = {0.f, 0.f, 0.f};
    if (use_viewdirs) unit_dir(d, vd);
    if (ndc) ndc_transform((float)H, (float)W, K.m[0], 1.f, o, d);
    write_ray(rays + (size_t)i * (use_viewdirs ? 11 : 8), o, d, near_, far_, use_viewdirs, vd);
}

// Batch sampler back end (train(): NP/run_nerf_view.py:1443-1517, NP/run_nerf.py:718-760): rays of SELECTED pixels of one
// view generated from the pose, plus the gathers of the target colour / prior depth / consistency mask at the same
// pixels -- one launch instead of a full-image get_rays, three host->device image copies and five fancy-index gathers.
__global__ void gather_rays_kernel(int H, int W, Mat3 K, Mat34 c2w, const int32_t* __restrict__ pix, int n, float near_,
                                   float far_, int use_viewdirs, int ndc, const float* __restrict__ img,
                                   const float* __restrict__ depth, const float* __restrict__ mask, float* __restrict__ rays,
                                   float* __restrict__ target, float* __restrict__ depth_out, float* __restrict__ mask_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pid = pix[i];
    const int py = pid / W, px = pid - py * W;
    float dx = __fdiv_rn(__fsub_rn((float)px, K.m[2]), K.m[0]);
    float dy = -__fdiv_rn(__fsub_rn((float)py, K.m[5]), K.m[4]);
    float dz = -1.f;
    float d[3], o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w.m[4 * r]), __fmul_rn(dy, c2w.m[4 * r + 1])), __fmul_rn(dz, c2w.m[4 * r + 2]));
        o[r] = c2w.m[4 * r + 3];
    }
    float vd[3] = {0.f, 0.f, 0.f};
    if (use_viewdirs) unit_dir(d, vd);
    if (ndc) ndc_transform((float)H, (float)W, K.m[0], 1.f, o, d);
    write_ray(rays + (size_t)i * (use_viewdirs ? 11 : 8), o, d, near_, far_, use_viewdirs, vd);
    if (target && img) {
        target[3 * (size_t)i] = __ldg(img + 3 * (size_t)pid); target[3 * (size_t)i + 1] = __ldg(img + 3 * (size_t)pid + 1);
        target[3 * (size_t)i + 2] = __ldg(img + 3 * (size_t)pid + 2);
    }
    if (depth_out && depth) depth_out[i] = __ldg(depth + pid);
    if (mask_out && mask) mask_out[i] = __ldg(mask + pid);
}

// ------------------------------------------------------------------------------------
// K1 stratified z  (NP/run_nerf.py:360-384)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float z_edge(float near_, float far_, float t, int lindisp) {
    float omt = __fsub_rn(1.f, t);
    if (!lindisp) return __fadd_rn(__fmul_rn(near_, omt), __fmul_rn(far_, t));
    float a = __fmul_rn(__fdiv_rn(1.f, near_), omt);
    float b = __fmul_rn(__fdiv_rn(1.f, far_), t);
    return __fdiv_rn(1.f, __fadd_rn(a, b));
}

__global__ void stratified_z_kernel(const float* __restrict__ rays, int stride, const float* __restrict__ t_vals,
                                    const float* __restrict__ t_rand, int n, int S, int lindisp,
                                    float* __restrict__ z, float* __restrict__ pts) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * S) return;
    int r = (int)(idx / S), s = (int)(idx - (int64_t)r * S);
    const float* ray = rays + (size_t)r * stride;
    float near_ = ray[6], far_ = ray[7];
    float zc = z_edge(near_, far_, t_vals[s], lindisp);
    if (t_rand != nullptr) {
        float lower = zc, upper = zc;
        if (s > 0) lower = __fmul_rn(0.5f, __fadd_rn(zc, z_edge(near_, far_, t_vals[s - 1], lindisp)));
        if (s < S - 1) upper = __fmul_rn(0.5f, __fadd_rn(z_edge(near_, far_, t_vals[s + 1], lindisp), zc));
        zc = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[idx]));
    }
    z[idx] = zc;
    if (pts != nullptr) {
        pts[3 * idx + 0] = __fadd_rn(ray[0], __fmul_rn(ray[3], zc));
        pts[3 * idx + 1] = __fadd_rn(ray[1], __fmul_rn(ray[4], zc));
        pts[3 * idx + 2] = __fadd_rn(ray[2], __fmul_rn(ray[5], zc));
    }
}

__global__ void ray_points_kernel(const float* __restrict__ rays, int stride, const float* __restrict__ z, int n,
                                  int S, float* __restrict__ pts) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * S) return;
    int r = (int)(idx / S);
    const float* ray = rays + (size_t)r * stride;
    float zc = z[idx];
    pts[3 * idx + 0] = __fadd_rn(ray[0], __fmul_rn(ray[3], zc));
    pts[3 * idx + 1] = __fadd_rn(ray[1], __fmul_rn(ray[4], zc));
    pts[3 * idx + 2] = __fadd_rn(ray[2], __fmul_rn(ray[5], zc));
}

// ------------------------------------------------------------------------------------
// K5 inverse-CDF sampling: one warp per ray, cdf and bins staged in shared memory.
// ------------------------------------------------------------------------------------
constexpr int kPdfWarps = 4;

// ---- sum(weights + 1e-5) in the summation order of ATen's CPU sum kernel ---------------------------
// The reference normalises the pdf with torch.sum(weights, -1) (NP/run_nerf_helpers.py:209).  Float
// addition is not associative and the rounded total decides the low bits of every cdf entry, hence the
// searchsorted indices at near-ties, so "bit-exact bin indices" needs a named summation order.  The
// oracle device is the CPU; its kernel (aten/src/ATen/native/cpu/SumKernel.cpp: vectorized_inner_sum ->
// row_sum -> multi_row_sum) adds 8-float vectors lane-wise with 4 interleaved accumulators and a
// 16-step cascade, then the scalar tail, then the 8 lane partials in order.  This restates that order
// (checked against torch.sum for every length 1..4095 in oracle/make_golden.py's container).
constexpr int kAtenVec = 8;

// row_sum over `size` elements x(i) = w[off + i*stride] + 1e-5, sequential on one thread
__device__ float aten_row_sum(const float* __restrict__ w, int off, int stride, int size) {
    auto X = [&](int i) { return __fadd_rn(w[off + i * stride], 1e-5f); };
    const int size_ilp = size / 4;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
    int i = 0;
    while (i + 16 <= size_ilp) {                      // level_step = 1 << max(4, CeilLog2(size)/4) = 16 for size < 2^20
        for (int j = 0; j < 16; ++j, ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[0][k] = __fadd_rn(acc[0][k], X(4 * i + k));
#pragma unroll
        for (int j = 1; j < 4; ++j) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { acc[j][k] = __fadd_rn(acc[j][k], acc[j - 1][k]); acc[j - 1][k] = 0.f; }
            if ((i & (15 << (4 * j))) != 0) break;
        }
    }
    for (; i < size_ilp; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[0][k] = __fadd_rn(acc[0][k], X(4 * i + k));
#pragma unroll
    for (int j = 1; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[0][k] = __fadd_rn(acc[0][k], acc[j][k]);
    float p0 = acc[0][0];
    for (int r = size_ilp * 4; r < size; ++r) p0 = __fadd_rn(p0, X(r));
    p0 = __fadd_rn(p0, acc[0][1]); p0 = __fadd_rn(p0, acc[0][2]); p0 = __fadd_rn(p0, acc[0][3]);
    return p0;
}

// whole-warp call; returns sum_j (w[j] + 1e-5) on every lane
__device__ float aten_cpu_sum(const float* __restrict__ w, int n, int lane) {
    float fin = 0.f;
    if (n >= kAtenVec) {
        const int vs = n / kAtenVec;
        float part = lane < kAtenVec ? aten_row_sum(w, lane, kAtenVec, vs) : 0.f;
        if (lane == 0) for (int k = vs * kAtenVec; k < n; ++k) fin = __fadd_rn(fin, __fadd_rn(w[k], 1e-5f));
#pragma unroll
        for (int j = 0; j < kAtenVec; ++j) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, part, j));
    } else if (lane == 0) {
        fin = aten_row_sum(w, 0, 1, n);
    }
    return __shfl_sync(0xffffffffu, fin, 0);
}

// Builds cdf[0..B-1] in shared memory from B-1 weights.  Follows sample_pdf
// (NP/run_nerf_helpers.py:208-211): w += 1e-5; pdf = w / sum(w); cdf = [0, cumsum(pdf)].
// The running cumsum is accumulated in double and rounded to fp32 per element, which is what torch's
// CPU cumsum produces (hard part 3); the total follows aten_cpu_sum above.
__device__ void warp_build_cdf(const float* __restrict__ w, int nw, float* cdf_s, int lane) {
    float total = aten_cpu_sum(w, nw, lane);
    // contiguous chunk per lane -> sequential inside, exclusive scan across lanes
    int per = (nw + 31) / 32;
    int j0 = lane * per, j1 = min(j0 + per, nw);
    double local = 0.0;
    for (int j = j0; j < j1; ++j) local += (double)__fdiv_rn(__fadd_rn(w[j], 1e-5f), total);
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    double run = incl - local;
    if (lane == 0) cdf_s[0] = 0.f;
    for (int j = j0; j < j1; ++j) {
        run += (double)__fdiv_rn(__fadd_rn(w[j], 1e-5f), total);
        cdf_s[j + 1] = (float)run;
    }
    __syncwarp();
}

// searchsorted(right=True) + the lerp of :233-248.
__device__ __forceinline__ float invert_cdf(const float* cdf_s, const float* bins_s, int B, float u, int* below_o,
                                            int* above_o) {
    int lo = 0, hi = B;                 // first index with cdf[idx] > u
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf_s[mid] <= u) lo = mid + 1; else hi = mid;
    }
    int below = max(lo - 1, 0), above = min(lo, B - 1);
    float cb = cdf_s[below], ca = cdf_s[above];
    float denom = __fsub_rn(ca, cb);
    if (denom < 1e-5f) denom = 1.f;
    float t = __fdiv_rn(__fsub_rn(u, cb), denom);
    float bb = bins_s[below], ba = bins_s[above];
    *below_o = below; *above_o = above;
    return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, const float* __restrict__ u,
                  const float* __restrict__ u_det, int n, int B, int M, float* __restrict__ samples,
                  float* __restrict__ cdf_out, int32_t* __restrict__ below_out, int32_t* __restrict__ above_out) {
    extern __shared__ float smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ray = blockIdx.x * kPdfWarps + warp;
    if (ray >= n) return;
    float* cdf_s = smem + (size_t)warp * 2 * B;
    float* bins_s = cdf_s + B;
    for (int j = lane; j < B; j += 32) bins_s[j] = bins[(size_t)ray * B + j];
    warp_build_cdf(weights + (size_t)ray * (B - 1), B - 1, cdf_s, lane);
    if (cdf_out) for (int j = lane; j < B; j += 32) cdf_out[(size_t)ray * B + j] = cdf_s[j];
    for (int k = lane; k < M; k += 32) {
        float uu = u ? u[(size_t)ray * M + k] : u_det[k];
        int below, above;
        float s = invert_cdf(cdf_s, bins_s, B, uu, &below, &above);
        samples[(size_t)ray * M + k] = s;
        if (below_out) below_out[(size_t)ray * M + k] = below;
        if (above_out) above_out[(size_t)ray * M + k] = above;
    }
}

// Fused fine sampling: mid-point bins, inverse CDF, bitonic merge with the coarse z, std.
__global__ void __launch_bounds__(kPdfWarps * 32)
sample_fine_kernel(const float* __restrict__ z, const float* __restrict__ weights, const float* __restrict__ u,
                   const float* __restrict__ u_det, int n, int S, int M, int P2, float* __restrict__ z_samples,
                   float* __restrict__ z_fine, float* __restrict__ z_std) {
    extern __shared__ float smem[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ray = blockIdx.x * kPdfWarps + warp;
    if (ray >= n) return;
    const int B = S - 1;
    float* cdf_s = smem + (size_t)warp * (2 * B + P2);
    float* bins_s = cdf_s + B;
    float* sort_s = bins_s + B;
    const float* zr = z + (size_t)ray * S;
    for (int j = lane; j < S; j += 32) sort_s[j] = zr[j];
    for (int j = lane; j < B; j += 32) bins_s[j] = __fmul_rn(0.5f, __fadd_rn(zr[j + 1], zr[j]));   // :393
    warp_build_cdf(weights + (size_t)ray * S + 1, B - 1, cdf_s, lane);                              // weights[...,1:-1]
    double s1 = 0.0;
    for (int k = lane; k < M; k += 32) {
        float uu = u ? u[(size_t)ray * M + k] : u_det[k];
        int below, above;
        float s = invert_cdf(cdf_s, bins_s, B, uu, &below, &above);
        sort_s[S + k] = s;
        s1 += (double)s;
        if (z_samples) z_samples[(size_t)ray * M + k] = s;
    }
    for (int j = S + M + lane; j < P2; j += 32) sort_s[j] = __int_as_float(0x7f800000);   // +inf padding
    __syncwarp();
    if (z_std) {      // population std (torch.std(unbiased=False), NP/run_nerf.py:415)
        double mean = warp_sum(s1) / (double)M, s2 = 0.0;
        for (int k = lane; k < M; k += 32) { double dlt = (double)sort_s[S + k] - mean; s2 += dlt * dlt; }
        s2 = warp_sum(s2);
        if (lane == 0) z_std[ray] = (float)sqrt(s2 / (double)M);
    }
    // bitonic sort of P2 values (ascending); value-only so the result equals torch.sort exactly
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P2; i += 32) {
                int p = i ^ j;
                if (p > i) {
                    float a = sort_s[i], b = sort_s[p];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { sort_s[i] = b; sort_s[p] = a; }
                }
            }
            __syncwarp();
        }
    }
    for (int j = lane; j < S + M; j += 32) z_fine[(size_t)ray * (S + M) + j] = sort_s[j];
}

}  // namespace cnerf

using namespace cnerf;

extern "C" int cnerf_pack_rays(const float* rays_o, const float* rays_d, int n, float near_, float far_,
                               int use_viewdirs, int ndc, int H, int W, float focal, float* rays, void* stream) {
    CNERF_REQUIRE(n >= 0 && rays_o && rays_d && rays, "cnerf_pack_rays: null pointer or negative n");
    if (n == 0) return CNERF_OK;
    pack_rays_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(rays_o, rays_d, n, near_, far_, use_viewdirs, ndc,
                                                                      (float)H, (float)W, focal, rays);
    CNERF_LAUNCH_CHECK("pack_rays_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_image_rays(int H, int W, const float* K_host, const float* c2w_host, float near_, float far_,
                                int use_viewdirs, int ndc, float* rays, void* stream) {
    CNERF_REQUIRE(H > 0 && W > 0 && K_host && c2w_host && rays, "cnerf_image_rays: bad arguments");
    Mat3 K; Mat34 P;
    for (int i = 0; i < 9; ++i) K.m[i] = K_host[i];
    for (int i = 0; i < 12; ++i) P.m[i] = c2w_host[i];
    image_rays_kernel<<<ceil_div(H * W, 256), 256, 0, as_stream(stream)>>>(H, W, K, P, near_, far_, use_viewdirs, ndc, rays);
    CNERF_LAUNCH_CHECK("image_rays_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_gather_rays(int H, int W, const float* K_host, const float* c2w_host, const int32_t* pix, int n,
                                 float near_, float far_, int use_viewdirs, int ndc, const float* img, const float* depth,
                                 const float* mask, float* rays, float* target, float* depth_out, float* mask_out, void* stream) {
    CNERF_REQUIRE(H > 0 && W > 0 && K_host && c2w_host && pix && rays && n >= 0, "cnerf_gather_rays: bad arguments");
    CNERF_REQUIRE(!(target && !img) && !(depth_out && !depth) && !(mask_out && !mask), "cnerf_gather_rays: gather output without source");
    if (n == 0) return CNERF_OK;
    Mat3 K; Mat34 P;
    for (int i = 0; i < 9; ++i) K.m[i] = K_host[i];
    for (int i = 0; i < 12; ++i) P.m[i] = c2w_host[i];
    gather_rays_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(H, W, K, P, pix, n, near_, far_, use_viewdirs, ndc, img, depth,
                                                                        mask, rays, target, depth_out, mask_out);
    CNERF_LAUNCH_CHECK("gather_rays_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_stratified_z(const float* rays, int ray_stride, const float* t_vals, const float* t_rand,
                                  int n_rays, int n_samples, int lindisp, float* z, float* pts, void* stream) {
    CNERF_REQUIRE(rays && t_vals && z, "cnerf_stratified_z: null pointer");
    CNERF_REQUIRE(ray_stride >= 8 && n_rays >= 0 && n_samples >= 1, "cnerf_stratified_z: bad sizes (stride %d, S %d)",
                  ray_stride, n_samples);
    if (n_rays == 0) return CNERF_OK;
    int64_t total = (int64_t)n_rays * n_samples;
    stratified_z_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(rays, ray_stride, t_vals, t_rand,
                                                                                         n_rays, n_samples, lindisp, z, pts);
    CNERF_LAUNCH_CHECK("stratified_z_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_ray_points(const float* rays, int ray_stride, const float* z, int n_rays, int n_samples,
                                float* pts, void* stream) {
    CNERF_REQUIRE(rays && z && pts && ray_stride >= 6, "cnerf_ray_points: bad arguments");
    if (n_rays == 0 || n_samples == 0) return CNERF_OK;
    int64_t total = (int64_t)n_rays * n_samples;
    ray_points_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(rays, ray_stride, z, n_rays, n_samples, pts);
    CNERF_LAUNCH_CHECK("ray_points_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_sample_pdf(const float* bins, const float* weights, const float* u, const float* u_det,
                                int n_rays, int n_bins, int n_new, float* samples, float* cdf, int32_t* below,
                                int32_t* above, void* stream) {
    CNERF_REQUIRE(bins && weights && samples && (u || u_det), "cnerf_sample_pdf: null pointer");
    CNERF_REQUIRE(n_bins >= 2 && n_bins <= 4096 && n_new >= 1, "cnerf_sample_pdf: n_bins %d must be in [2,4096]", n_bins);
    if (n_rays == 0) return CNERF_OK;
    size_t smem = (size_t)kPdfWarps * 2 * n_bins * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(sample_pdf_kernel)");
    }
    sample_pdf_kernel<<<ceil_div(n_rays, kPdfWarps), kPdfWarps * 32, smem, as_stream(stream)>>>(
        bins, weights, u, u_det, n_rays, n_bins, n_new, samples, cdf, below, above);
    CNERF_LAUNCH_CHECK("sample_pdf_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_sample_fine(const float* z, const float* weights, const float* u, const float* u_det, int n_rays,
                                 int n_samples, int n_new, float* z_samples, float* z_fine, float* z_std, void* stream) {
    CNERF_REQUIRE(z && weights && z_fine && (u || u_det), "cnerf_sample_fine: null pointer");
    CNERF_REQUIRE(n_samples >= 4 && n_new >= 1 && n_samples + n_new <= 8192, "cnerf_sample_fine: bad sizes S=%d M=%d",
                  n_samples, n_new);
    if (n_rays == 0) return CNERF_OK;
    int P2 = 1;
    while (P2 < n_samples + n_new) P2 <<= 1;
    size_t smem = (size_t)kPdfWarps * (2 * (n_samples - 1) + P2) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sample_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(sample_fine_kernel)");
    }
    sample_fine_kernel<<<ceil_div(n_rays, kPdfWarps), kPdfWarps * 32, smem, as_stream(stream)>>>(
        z, weights, u, u_det, n_rays, n_samples, n_new, P2, z_samples, z_fine, z_std);
    CNERF_LAUNCH_CHECK("sample_fine_kernel");
    return CNERF_OK;
}
