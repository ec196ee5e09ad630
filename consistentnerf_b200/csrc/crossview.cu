// K6: cross-view warp / project / gather / occlusion test, and K7: masked consistency losses.
//
// Both are HBM-bound integer/float row kernels (28 B in + a 16 B gather + ~13 B out per pixel pair for
// K6, 36 B per ray for K7).  The projection arithmetic is evaluated left to right with explicit
// round-to-nearest intrinsics (no FMA contraction) so that the rounded pixel coordinates are
// reproducible bit for bit against the oracle (oracle/nerf_oracle.py: project_points).
#include "common.cuh"
#include "soft_weight.h"

namespace cnerf {

struct Cam { float R[9]; float T[3]; float K[9]; };

__device__ __forceinline__ float dot3_rn(float x, float y, float z, float a, float b, float c) {
    float v = __fmul_rn(x, a);
    v = __fadd_rn(v, __fmul_rn(y, b));
    v = __fadd_rn(v, __fmul_rn(z, c));
    return v;
}

struct Projection { float px, py, xc, yc, zc; bool inb; };

// get_ref_rays / get_test_label, NP/run_nerf_view.py:590-613
__device__ __forceinline__ Projection project(const Cam& cam, float x, float y, float z, int H, int W) {
    Projection p;
    float c0 = __fadd_rn(dot3_rn(x, y, z, cam.R[0], cam.R[1], cam.R[2]), cam.T[0]);
    float c1 = __fadd_rn(dot3_rn(x, y, z, cam.R[3], cam.R[4], cam.R[5]), cam.T[1]);
    float c2 = __fadd_rn(dot3_rn(x, y, z, cam.R[6], cam.R[7], cam.R[8]), cam.T[2]);
    p.xc = c0; p.yc = -c1; p.zc = -c2;                         // @ diag(1,-1,-1)
    float u = dot3_rn(p.xc, p.yc, p.zc, cam.K[0], cam.K[1], cam.K[2]);
    float v = dot3_rn(p.xc, p.yc, p.zc, cam.K[3], cam.K[4], cam.K[5]);
    float w = dot3_rn(p.xc, p.yc, p.zc, cam.K[6], cam.K[7], cam.K[8]);
    p.px = rintf(__fadd_rn(__fdiv_rn(u, w), 0.0f));            // torch.round: half to even
    p.py = rintf(__fadd_rn(__fdiv_rn(v, w), 0.0f));
    float nx = __fdiv_rn(p.px, (float)(W - 1)), ny = __fdiv_rn(p.py, (float)(H - 1));
    p.inb = (nx > 0.f) && (nx < 1.f) && (ny > 0.f) && (ny < 1.f);   // strict on both sides
    return p;
}

__global__ void project_gather_kernel(const float* __restrict__ pts, int n, Cam cam, Mat34 c2w, int have_c2w,
                                      const float* __restrict__ img, int C, const float* __restrict__ depth, int H,
                                      int W, float* __restrict__ px, float* __restrict__ py,
                                      uint8_t* __restrict__ mask, float* __restrict__ cam_pts,
                                      float* __restrict__ rgb_ref, float* __restrict__ depth_ref,
                                      float* __restrict__ ray_o, float* __restrict__ ray_d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Projection p = project(cam, pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], H, W);
    if (px) px[i] = p.px;
    if (py) py[i] = p.py;
    if (mask) mask[i] = p.inb ? 1 : 0;
    if (cam_pts) { cam_pts[3 * (size_t)i] = p.xc; cam_pts[3 * (size_t)i + 1] = p.yc; cam_pts[3 * (size_t)i + 2] = p.zc; }
    int xi = p.inb ? (int)p.px : 0, yi = p.inb ? (int)p.py : 0;
    if (rgb_ref && img) {
        for (int c = 0; c < C; ++c)
            rgb_ref[(size_t)i * C + c] = p.inb ? __ldg(img + ((size_t)c * H + yi) * W + xi) : 0.f;
    }
    if (depth_ref && depth) depth_ref[i] = p.inb ? __ldg(depth + (size_t)yi * W + xi) : 0.f;
    if (ray_d && have_c2w) {
        // directions through the rounded pixel, rotated by c2w (:615-620, get_rays_ref :553-574)
        float dx = __fdiv_rn(__fsub_rn(p.px, cam.K[2]), cam.K[0]);
        float dy = __fdiv_rn(__fsub_rn(p.py, cam.K[5]), cam.K[4]);
        for (int r = 0; r < 3; ++r) {
            ray_d[3 * (size_t)i + r] = dx * c2w.m[4 * r] + dy * c2w.m[4 * r + 1] + c2w.m[4 * r + 2];
            if (ray_o) ray_o[3 * (size_t)i + r] = c2w.m[4 * r + 3];
        }
    }
}

// One block per `chunk` pixels (the reference's threshold doubling is defined per chunk,
// NP/run_nerf_view.py:1014,1026-1029, so the chunk is a semantic parameter).
__global__ void __launch_bounds__(256)
hard_mask_kernel(const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ dt, int n,
                 Cam cam, const float* __restrict__ depth_ref, int H, int W, float thr0, int chunk, int accumulate,
                 uint8_t* __restrict__ mask) {
    __shared__ float s_min[8];
    __shared__ float s_thr;
    int c0 = blockIdx.x * chunk, c1 = min(c0 + chunk, n);
    float best = __int_as_float(0x7f800000);
    for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
        float d = dt[i];
        float x = __fadd_rn(ro[3 * (size_t)i], __fmul_rn(d, rd[3 * (size_t)i]));
        float y = __fadd_rn(ro[3 * (size_t)i + 1], __fmul_rn(d, rd[3 * (size_t)i + 1]));
        float z = __fadd_rn(ro[3 * (size_t)i + 2], __fmul_rn(d, rd[3 * (size_t)i + 2]));
        Projection p = project(cam, x, y, z, H, W);
        if (p.inb) {
            float diff = fabsf(__fsub_rn(p.zc, __ldg(depth_ref + (size_t)((int)p.py) * W + (int)p.px)));
            best = fminf(best, diff);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_min[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fminf(m, s_min[w]);
        float thr = -1.f;                                   // no in-bounds pixel: nothing passes
        if (m < __int_as_float(0x7f800000)) {
            thr = thr0;
            for (int it = 0; it < 200 && !(m < thr); ++it) thr = 2.f * thr;
        }
        s_thr = thr;
    }
    __syncthreads();
    float thr = s_thr;
    for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
        float d = dt[i];
        float x = __fadd_rn(ro[3 * (size_t)i], __fmul_rn(d, rd[3 * (size_t)i]));
        float y = __fadd_rn(ro[3 * (size_t)i + 1], __fmul_rn(d, rd[3 * (size_t)i + 1]));
        float z = __fadd_rn(ro[3 * (size_t)i + 2], __fmul_rn(d, rd[3 * (size_t)i + 2]));
        Projection p = project(cam, x, y, z, H, W);
        uint8_t hit = 0;
        if (p.inb) {
            float diff = fabsf(__fsub_rn(p.zc, __ldg(depth_ref + (size_t)((int)p.py) * W + (int)p.px)));
            hit = diff < thr ? 1 : 0;
        }
        mask[i] = accumulate ? (uint8_t)(mask[i] | hit) : hit;
    }
}

// ------------------------------------------------------------------------------------
// K7 masked MSE.  Deterministic: per-block partials in fixed slots, one block folds them in order.
// ------------------------------------------------------------------------------------
constexpr int kLossBlocks = 148;
constexpr int kLossThreads = 256;

struct LossAcc { double s1, s0, n1, n0, sall, msum; };

__global__ void __launch_bounds__(kLossThreads)
masked_mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                          const float* __restrict__ mask, int n, int C, float divisor, double* __restrict__ part) {
    LossAcc a = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float m = mask ? mask[i] : 1.f;
        double e = 0.0;
        for (int c = 0; c < C; ++c) {
            float d = __fsub_rn(__fdiv_rn(pred[(size_t)i * C + c], divisor), __fdiv_rn(target[(size_t)i * C + c], divisor));
            e += (double)(d * d);
        }
        a.sall += e; a.msum += (double)m;
        if (m == 1.f) { a.s1 += e; a.n1 += 1.0; }
        else if (m == 0.f) { a.s0 += e; a.n0 += 1.0; }
    }
    __shared__ double sh[kLossThreads / 32][6];
    double v[6] = {a.s1, a.s0, a.n1, a.n0, a.sall, a.msum};
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 6; ++k) sh[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) t += sh[w][threadIdx.x];
        part[(size_t)blockIdx.x * 6 + threadIdx.x] = t;
    }
}

__global__ void masked_mse_final_kernel(const double* __restrict__ part, int nblocks, int n, int C, float coef,
                                        float n_ref, int use_unmasked, const float* __restrict__ global_counts,
                                        float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nblocks; ++b)
        for (int k = 0; k < 6; ++k) v[k] += part[(size_t)b * 6 + k];
    double nrows = (double)n;
    if (global_counts) {      // this rank's share of a loss whose means run over the GLOBAL batch: denominators from all ranks
        v[2] = (double)global_counts[0]; v[3] = (double)global_counts[1]; v[5] = (double)global_counts[2];
        nrows = (double)global_counts[3];
    }
    double loss = v[0] / (v[2] * C);
    if (use_unmasked && v[5] != (double)n_ref) loss += (double)coef * (v[1] / (v[3] * C));
    out[0] = (float)loss;
    out[1] = (float)v[2];
    out[2] = (float)v[3];
    out[3] = (float)(v[4] / (nrows * C));
    out[4] = (float)v[5];
}

__global__ void masked_mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                      const float* __restrict__ mask, int n, int C, float divisor, float coef,
                                      float n_ref, int use_unmasked, const float* __restrict__ out,
                                      const float* __restrict__ g_loss, float* __restrict__ d_pred) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * C) return;
    int i = (int)(idx / C);
    float m = mask ? mask[i] : 1.f;
    float n1 = out[1], n0 = out[2], msum = out[4];
    float wgt = 0.f;
    if (m == 1.f) wgt = 1.f / (n1 * C);
    else if (m == 0.f && use_unmasked && msum != n_ref) wgt = coef / (n0 * C);
    float d = pred[idx] / divisor - target[idx] / divisor;
    d_pred[idx] = g_loss[0] * wgt * 2.f * d / divisor;
}

// ------------------------------------------------------------------------------------
// K7b soft-weighted MSE (img2mse_softmask / img2mse_depth_softmask / img2mse_softLpmask, NP/run_nerf_view.py:50-58);
// element arithmetic in soft_weight.h.  Same deterministic two-stage reduction as K7.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLossThreads)
soft_mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t total, float divisor,
                        int kind, float param_host, const float* __restrict__ param_dev, double* __restrict__ part) {
    const float param = param_dev ? param_dev[0] : param_host;
    double num = 0.0, den = 0.0, s4 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float d = cnerf_soft_residual(pred[i], target[i], divisor);
        double w = (double)cnerf_soft_weight(d, kind, param), e = (double)(d * d);
        num += w * e; den += w; s4 += w * e * e;
    }
    __shared__ double sh[kLossThreads / 32][3];
    double v[3] = {warp_sum(num), warp_sum(den), warp_sum(s4)};
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; ++k) sh[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) t += sh[w][threadIdx.x];
        part[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

__global__ void soft_mse_final_kernel(const double* __restrict__ part, int nblocks, int kind, float param_host,
                                      const float* __restrict__ param_dev, float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    const float param = param_dev ? param_dev[0] : param_host;
    double v[3] = {0, 0, 0};
    for (int b = 0; b < nblocks; ++b)
        for (int k = 0; k < 3; ++k) v[k] += part[(size_t)b * 3 + k];
    double loss, dparam;
    cnerf_soft_finish(v[0], v[1], v[2], kind, (double)param, &loss, &dparam);
    out[0] = (float)loss;
    out[1] = (float)v[0];
    out[2] = (float)v[1];
    out[3] = (float)dparam;
    out[4] = param;
}

__global__ void soft_mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t total,
                                    float divisor, int kind, const float* __restrict__ out,
                                    const float* __restrict__ g_loss, float* __restrict__ d_pred) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float d = cnerf_soft_residual(pred[i], target[i], divisor);
    d_pred[i] = g_loss[0] * cnerf_soft_dnum(d, kind, out[4]) / out[2] / divisor;
}

static Cam make_cam(const float* w2c, const float* K) {
    Cam c;
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) c.R[3 * r + k] = w2c[4 * r + k];
        c.T[r] = w2c[4 * r + 3];
    }
    for (int i = 0; i < 9; ++i) c.K[i] = K[i];
    return c;
}

}  // namespace cnerf

using namespace cnerf;

extern "C" int cnerf_project_gather(const float* pts_w, int n, const float* w2c_host, const float* K_host,
                                    const float* c2w_host, const float* img, int C, const float* depth, int H, int W,
                                    float* px, float* py, uint8_t* mask, float* cam, float* rgb_ref, float* depth_ref,
                                    float* ref_rays_o, float* ref_rays_d, void* stream) {
    CNERF_REQUIRE(pts_w && w2c_host && K_host, "cnerf_project_gather: null pointer");
    CNERF_REQUIRE(n >= 0 && H > 1 && W > 1 && C >= 0, "cnerf_project_gather: bad sizes");
    CNERF_REQUIRE(!(rgb_ref && !img) && !(depth_ref && !depth), "cnerf_project_gather: gather output without source");
    if (n == 0) return CNERF_OK;
    Mat34 P = {};
    if (c2w_host) for (int i = 0; i < 12; ++i) P.m[i] = c2w_host[i];
    project_gather_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        pts_w, n, make_cam(w2c_host, K_host), P, c2w_host != nullptr, img, C, depth, H, W, px, py, mask, cam, rgb_ref,
        depth_ref, ref_rays_o, ref_rays_d);
    CNERF_LAUNCH_CHECK("project_gather_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_hard_mask_pair(const float* rays_o, const float* rays_d, const float* depth_tgt, int n,
                                    const float* w2c_host, const float* K_host, const float* depth_ref, int H, int W,
                                    float thr0, int chunk, int accumulate, uint8_t* mask, void* stream) {
    CNERF_REQUIRE(rays_o && rays_d && depth_tgt && w2c_host && K_host && depth_ref && mask, "cnerf_hard_mask_pair: null pointer");
    CNERF_REQUIRE(n >= 0 && chunk > 0 && H > 1 && W > 1 && thr0 > 0.f, "cnerf_hard_mask_pair: bad sizes");
    if (n == 0) return CNERF_OK;
    hard_mask_kernel<<<ceil_div(n, chunk), 256, 0, as_stream(stream)>>>(rays_o, rays_d, depth_tgt, n,
                                                                          make_cam(w2c_host, K_host), depth_ref, H, W,
                                                                          thr0, chunk, accumulate, mask);
    CNERF_LAUNCH_CHECK("hard_mask_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_masked_mse_fwd(const float* pred, const float* target, const float* mask, int n, int C,
                                    float divisor, float coef, float n_ref, int use_unmasked, const float* global_counts,
                                    float* out, void* workspace, void* stream) {
    CNERF_REQUIRE(pred && target && out && workspace, "cnerf_masked_mse_fwd: null pointer");
    CNERF_REQUIRE(n >= 0 && C >= 1, "cnerf_masked_mse_fwd: bad sizes");
    int blocks = n == 0 ? 1 : (ceil_div(n, kLossThreads) < kLossBlocks ? ceil_div(n, kLossThreads) : kLossBlocks);
    double* part = reinterpret_cast<double*>(workspace);
    masked_mse_partial_kernel<<<blocks, kLossThreads, 0, as_stream(stream)>>>(pred, target, mask, n, C, divisor, part);
    CNERF_LAUNCH_CHECK("masked_mse_partial_kernel");
    masked_mse_final_kernel<<<1, 32, 0, as_stream(stream)>>>(part, blocks, n, C, coef, n_ref, use_unmasked, global_counts, out);
    CNERF_LAUNCH_CHECK("masked_mse_final_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_masked_mse_bwd(const float* pred, const float* target, const float* mask, int n, int C,
                                    float divisor, float coef, float n_ref, int use_unmasked, const float* out,
                                    const float* g_loss, float* d_pred, void* stream) {
    CNERF_REQUIRE(pred && target && out && g_loss && d_pred, "cnerf_masked_mse_bwd: null pointer");
    if (n == 0) return CNERF_OK;
    int64_t total = (int64_t)n * C;
    masked_mse_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, as_stream(stream)>>>(
        pred, target, mask, n, C, divisor, coef, n_ref, use_unmasked, out, g_loss, d_pred);
    CNERF_LAUNCH_CHECK("masked_mse_bwd_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_soft_mse_fwd(const float* pred, const float* target, int64_t n_elems, float divisor, int kind,
                                  float param, const float* param_dev, float* out, void* workspace, void* stream) {
    CNERF_REQUIRE(((pred && target) || n_elems == 0) && out && workspace, "cnerf_soft_mse_fwd: null pointer");
    CNERF_REQUIRE(n_elems >= 0 && (kind == 0 || kind == 1) && divisor != 0.f, "cnerf_soft_mse_fwd: bad arguments");
    CNERF_REQUIRE(param_dev || (kind == 0 ? param > 0.f : param >= 1.f),
                  "cnerf_soft_mse_fwd: temp must be > 0 (kind 0), the exponent >= 1 (kind 1)");
    int64_t want = ceil_div64(n_elems, kLossThreads);
    int blocks = want < 1 ? 1 : (want < kLossBlocks ? (int)want : kLossBlocks);
    double* part = reinterpret_cast<double*>(workspace);
    soft_mse_partial_kernel<<<blocks, kLossThreads, 0, as_stream(stream)>>>(pred, target, n_elems, divisor, kind, param,
                                                                            param_dev, part);
    CNERF_LAUNCH_CHECK("soft_mse_partial_kernel");
    soft_mse_final_kernel<<<1, 32, 0, as_stream(stream)>>>(part, blocks, kind, param, param_dev, out);
    CNERF_LAUNCH_CHECK("soft_mse_final_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_soft_mse_bwd(const float* pred, const float* target, int64_t n_elems, float divisor, int kind,
                                  const float* out, const float* g_loss, float* d_pred, void* stream) {
    CNERF_REQUIRE(((pred && target && d_pred) || n_elems == 0) && out && g_loss, "cnerf_soft_mse_bwd: null pointer");
    CNERF_REQUIRE(n_elems >= 0 && (kind == 0 || kind == 1) && divisor != 0.f, "cnerf_soft_mse_bwd: bad arguments");
    if (n_elems == 0) return CNERF_OK;
    soft_mse_bwd_kernel<<<(unsigned)ceil_div64(n_elems, 256), 256, 0, as_stream(stream)>>>(pred, target, n_elems, divisor,
                                                                                           kind, out, g_loss, d_pred);
    CNERF_LAUNCH_CHECK("soft_mse_bwd_kernel");
    return CNERF_OK;
}
