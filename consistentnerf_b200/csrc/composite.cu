// K4: alpha compositing along the ray (raw2outputs, NP/run_nerf.py:265-308) forward and backward.
//
// HBM-bound: 16 B of raw + 4 B of z per sample in, 4 B of weight out.  One warp owns one ray;
// lane l handles samples l, l+32, ... so every load is a fully coalesced 128-bit access, and the
// exclusive transmittance product is a 5-step shuffle scan per 32-sample chunk with a scalar carry.
#include "common.cuh"

namespace cnerf {

constexpr int kRaysPerBlock = 8;   // 8 warps

struct SampleTerms {
    float alpha, factor, dist;     // factor = 1 - alpha + 1e-10 (identity 1 for padding lanes)
    float c0, c1, c2;              // sigmoid(raw rgb)
    float zval;
    float sig;                     // raw sigma + noise
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ SampleTerms load_sample(const float* __restrict__ raw, const float* __restrict__ z,
                                                   const float* __restrict__ noise, int S, int s, float dnorm) {
    SampleTerms t;
    if (s < S) {
        float4 r = ldg_f4(raw + 4 * (size_t)s);
        float zc = z[s];
        float d = (s == S - 1) ? 1e10f : (z[s + 1] - zc);       // :280-281
        t.dist = d * dnorm;                                      // :283
        t.sig = noise ? r.w + noise[s] : r.w;
        t.alpha = 1.f - expf(-fmaxf(t.sig, 0.f) * t.dist);       // :278
        t.factor = (1.f - t.alpha) + 1e-10f;                     // :298
        t.c0 = sigmoidf_(r.x); t.c1 = sigmoidf_(r.y); t.c2 = sigmoidf_(r.z);
        t.zval = zc;
    } else {
        t.alpha = 0.f; t.factor = 1.f; t.dist = 0.f; t.c0 = t.c1 = t.c2 = 0.f; t.zval = 0.f; t.sig = 0.f;
    }
    return t;
}

// inclusive product scan over the warp; returns the exclusive value for this lane and the warp total
__device__ __forceinline__ float warp_excl_prod(float v, int lane, float* total) {
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= n;
    }
    *total = __shfl_sync(0xffffffffu, incl, 31);
    float ex = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 1.f : ex;
}

__device__ __forceinline__ float ray_norm(const float* d) {
    return sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

// torch: 1 / max(1e-10, depth/acc), NaN (0/0) propagates (:302)
__device__ __forceinline__ float disparity(float depth, float acc) {
    float q = depth / acc;
    if (q != q) return q;
    return 1.f / fmaxf(1e-10f, q);
}

__global__ void __launch_bounds__(kRaysPerBlock * 32)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     int d_stride, const float* __restrict__ noise, int n, int S, int white,
                     float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc,
                     float* __restrict__ depth, float* __restrict__ weights) {
    int lane = threadIdx.x & 31;
    int ray = blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
    if (ray >= n) return;
    const float* rr = raw + (size_t)ray * S * 4;
    const float* zr = z + (size_t)ray * S;
    const float* nr = noise ? noise + (size_t)ray * S : nullptr;
    float dnorm = ray_norm(rays_d + (size_t)ray * d_stride);
    float carry = 1.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, ad = 0.f, aw = 0.f;
    for (int base = 0; base < S; base += 32) {
        int s = base + lane;
        SampleTerms t = load_sample(rr, zr, nr, S, s, dnorm);
        float tot;
        float T = carry * warp_excl_prod(t.factor, lane, &tot);
        carry *= tot;
        float w = t.alpha * T;
        if (s < S && weights) weights[(size_t)ray * S + s] = w;
        a0 += w * t.c0; a1 += w * t.c1; a2 += w * t.c2; ad += w * t.zval; aw += w;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); ad = warp_sum(ad); aw = warp_sum(aw);
    if (lane == 0) {
        if (white) { float bg = 1.f - aw; a0 += bg; a1 += bg; a2 += bg; }   // :305-306
        rgb[3 * (size_t)ray + 0] = a0; rgb[3 * (size_t)ray + 1] = a1; rgb[3 * (size_t)ray + 2] = a2;
        if (disp) disp[ray] = disparity(ad, aw);
        if (acc) acc[ray] = aw;
        if (depth) depth[ray] = ad;
    }
}

// Backward.  Pass 1 re-runs the forward scan and parks (T, alpha, dist) in the d_raw slot of each
// sample; pass 2 walks the chunks in reverse with a suffix sum of G_j w_j:
//   dL/dalpha_s = G_s T_s - (sum_{j>s} G_j w_j) / (1 - alpha_s + 1e-10)
// Each lane only re-reads what it wrote itself, so no block-level synchronisation is needed.
__global__ void __launch_bounds__(kRaysPerBlock * 32)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     int d_stride, const float* __restrict__ noise, int n, int S, int white,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_disp,
                     const float* __restrict__ g_acc, const float* __restrict__ g_depth,
                     const float* __restrict__ g_weights, float* __restrict__ d_raw) {
    int lane = threadIdx.x & 31;
    int ray = blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
    if (ray >= n) return;
    const float* rr = raw + (size_t)ray * S * 4;
    const float* zr = z + (size_t)ray * S;
    const float* nr = noise ? noise + (size_t)ray * S : nullptr;
    float4* out = reinterpret_cast<float4*>(d_raw + (size_t)ray * S * 4);
    float dnorm = ray_norm(rays_d + (size_t)ray * d_stride);

    float carry = 1.f, ad = 0.f, aw = 0.f;
    for (int base = 0; base < S; base += 32) {
        int s = base + lane;
        SampleTerms t = load_sample(rr, zr, nr, S, s, dnorm);
        float tot;
        float T = carry * warp_excl_prod(t.factor, lane, &tot);
        carry *= tot;
        float w = t.alpha * T;
        ad += w * t.zval; aw += w;
        if (s < S) out[s] = make_float4(T, t.alpha, t.dist, 0.f);
    }
    ad = warp_sum(ad); aw = warp_sum(aw);

    float gr0 = g_rgb[3 * (size_t)ray], gr1 = g_rgb[3 * (size_t)ray + 1], gr2 = g_rgb[3 * (size_t)ray + 2];
    float gA = g_acc ? g_acc[ray] : 0.f;
    float gD = g_depth ? g_depth[ray] : 0.f;
    if (white) gA -= gr0 + gr1 + gr2;
    if (g_disp) {
        float q = ad / aw;
        if (q == q && q > 1e-10f) {               // max() selected depth/acc
            float gq = -g_disp[ray] / (q * q);    // d(1/q)/dq
            gD += gq / aw;
            gA += -gq * q / aw;
        }
    }
    const float* gw = g_weights ? g_weights + (size_t)ray * S : nullptr;

    float suffix = 0.f;                            // sum over later chunks of G_j w_j
    int last_base = ((S - 1) / 32) * 32;
    for (int base = last_base; base >= 0; base -= 32) {
        int s = base + lane;
        float T = 0.f, alpha = 0.f, dist = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, sig = 0.f, zc = 0.f;
        if (s < S) {
            float4 p = out[s];
            T = p.x; alpha = p.y; dist = p.z;
            float4 r = ldg_f4(rr + 4 * (size_t)s);
            c0 = sigmoidf_(r.x); c1 = sigmoidf_(r.y); c2 = sigmoidf_(r.z);
            sig = nr ? r.w + nr[s] : r.w;
            zc = zr[s];
        }
        float w = alpha * T;
        float G = gr0 * c0 + gr1 * c1 + gr2 * c2 + gD * zc + gA + ((gw && s < S) ? gw[s] : 0.f);
        float gwv = (s < S) ? G * w : 0.f;
        // exclusive suffix sum over lanes above this one
        float incl = gwv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float v = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += v;
        }
        float after = (incl - gwv) + suffix;
        suffix += __shfl_sync(0xffffffffu, incl, 0);
        if (s < S) {
            float factor = (1.f - alpha) + 1e-10f;
            float dalpha = G * T - after / factor;
            float dsig = (sig > 0.f) ? dalpha * dist * (1.f - alpha) : 0.f;
            out[s] = make_float4(gr0 * w * c0 * (1.f - c0), gr1 * w * c1 * (1.f - c1), gr2 * w * c2 * (1.f - c2), dsig);
        }
    }
}

}  // namespace cnerf

using namespace cnerf;

extern "C" int cnerf_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                   const float* noise, int n_rays, int n_samples, int white_bkgd, float* rgb,
                                   float* disp, float* acc, float* depth, float* weights, void* stream) {
    CNERF_REQUIRE(raw && z && rays_d && rgb, "cnerf_composite_fwd: null pointer");
    CNERF_REQUIRE(n_rays >= 0 && n_samples >= 1 && d_stride >= 3, "cnerf_composite_fwd: bad sizes");
    if (n_rays == 0) return CNERF_OK;
    composite_fwd_kernel<<<ceil_div(n_rays, kRaysPerBlock), kRaysPerBlock * 32, 0, as_stream(stream)>>>(
        raw, z, rays_d, d_stride, noise, n_rays, n_samples, white_bkgd, rgb, disp, acc, depth, weights);
    CNERF_LAUNCH_CHECK("composite_fwd_kernel");
    return CNERF_OK;
}

extern "C" int cnerf_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                   const float* noise, int n_rays, int n_samples, int white_bkgd, const float* g_rgb,
                                   const float* g_disp, const float* g_acc, const float* g_depth,
                                   const float* g_weights, float* d_raw, void* stream) {
    CNERF_REQUIRE(raw && z && rays_d && g_rgb && d_raw, "cnerf_composite_bwd: null pointer");
    CNERF_REQUIRE(n_rays >= 0 && n_samples >= 1 && d_stride >= 3, "cnerf_composite_bwd: bad sizes");
    if (n_rays == 0) return CNERF_OK;
    composite_bwd_kernel<<<ceil_div(n_rays, kRaysPerBlock), kRaysPerBlock * 32, 0, as_stream(stream)>>>(
        raw, z, rays_d, d_stride, noise, n_rays, n_samples, white_bkgd, g_rgb, g_disp, g_acc, g_depth, g_weights, d_raw);
    CNERF_LAUNCH_CHECK("composite_bwd_kernel");
    return CNERF_OK;
}
