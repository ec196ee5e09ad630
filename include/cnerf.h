/*
 * cnerf.h -- C ABI of the B200-native ConsistentNeRF per-ray hot path.
 *
 * This is the drop-in boundary: a plain `extern "C"` shared library (libcnerf.so) with
 * raw device pointers, sizes and an explicit stream.  No torch types cross it.  The
 * Python host side (consistentnerf_b200/) binds it with ctypes and re-exposes the
 * reference's own function surface (render / batchify_rays / render_rays / raw2outputs /
 * sample_pdf / run_network / NeRF / get_embedder).  `NP/` below abbreviates
 * /root/reference/nerf-pytorch-master/ -- each entry point cites what it replaces.
 *
 * Conventions
 *   - every `const float*` / `float*` is a DEVICE pointer to contiguous fp32 unless the
 *     comment says "host"; small camera matrices are HOST pointers and are passed to the
 *     kernels by value;
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); all work is stream
 *     ordered, nothing synchronises, the library never allocates caller-visible memory
 *     (the only device allocation it owns is inside a cnerf_weights handle);
 *   - return value: 0 on success, a CNERF_E* code otherwise; the message is retrievable
 *     per thread with cnerf_last_error().  Nothing throws across the boundary.
 */
#ifndef CNERF_H_
#define CNERF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNERF_VERSION 200          /* 0.2.0: gradient-precision arguments on the K3b entry points */

#define CNERF_OK        0
#define CNERF_EINVAL    1          /* bad argument (shape, null pointer, unsupported size) */
#define CNERF_ECUDA     2          /* CUDA runtime error; message holds cudaGetErrorString */
#define CNERF_EUNSUP    3          /* configuration not handled by this kernel */

int cnerf_version(void);
/* Copies the calling thread's last error message (NUL terminated) into buf; returns its length. */
int cnerf_last_error(char* buf, int len);
/* Compute capability major*10+minor of the current device, SM count in *sms (may be NULL). */
int cnerf_device_info(int* cc, int* sms);

/* ------------------------------------------------------------------------------------------
 * Ray preparation -- render() NP/run_nerf.py:95-126, get_rays NP/run_nerf_helpers.py:164-173,
 * ndc_rays :186-202.
 * ---------------------------------------------------------------------------------------- */

/* rays_o/rays_d [n,3] -> rays [n, 8 or 11] = [o, d, near, far, (d/|d|)].  ndc != 0 applies
 * ndc_rays(H, W, focal, near=1) to o,d after the view direction was taken. */
int cnerf_pack_rays(const float* rays_o, const float* rays_d, int n, float near_, float far_,
                    int use_viewdirs, int ndc, int H, int W, float focal, float* rays, void* stream);

/* Pinhole rays of a whole H x W image from a camera-to-world pose, packed like above.
 * K_host: 3x3 row-major, c2w_host: 3x4 row-major (both HOST). */
int cnerf_image_rays(int H, int W, const float* K_host, const float* c2w_host, float near_, float far_,
                     int use_viewdirs, int ndc, float* rays, void* stream);

/* Batch-sampler back end (train(): NP/run_nerf_view.py:1443-1517, NP/run_nerf.py:718-760): packed rays of the SELECTED pixels
 * pix[n] (int32 row-major pixel ids, device) of one view, and the gathers of target colour img [H,W,3], prior depth [H,W]
 * and mask [H,W] (device fp32, any may be NULL together with its output) at the same pixels. */
int cnerf_gather_rays(int H, int W, const float* K_host, const float* c2w_host, const int32_t* pix, int n, float near_,
                      float far_, int use_viewdirs, int ndc, const float* img, const float* depth, const float* mask,
                      float* rays, float* target, float* depth_out, float* mask_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1 stratified sampling -- render_rays NP/run_nerf.py:354-384.
 * ---------------------------------------------------------------------------------------- */

/* z[n,S] = near(1-t)+far t   (lindisp: 1/(1/near(1-t)+1/far t)); t_vals [S] is the
 * torch.linspace(0,1,S) of the caller so that z is bit-identical to the reference on the same
 * device.  t_rand [n,S] (may be NULL) adds the stratified jitter.  pts [n,S,3] (may be NULL)
 * receives o + d*z.  Arithmetic is unfused and in the reference's operation order. */
int cnerf_stratified_z(const float* rays, int ray_stride, const float* t_vals, const float* t_rand,
                       int n_rays, int n_samples, int lindisp, float* z, float* pts, void* stream);

/* pts[n,S,3] = o + d * z  (NP/run_nerf.py:384,400). */
int cnerf_ray_points(const float* rays, int ray_stride, const float* z, int n_rays, int n_samples,
                     float* pts, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2 positional encoding -- Embedder NP/run_nerf_helpers.py:15-46.
 * out[i, col0 : col0 + C*(1+2L)] = [x, sin(2^0 x), cos(2^0 x), ...]; x is [n, C] with row
 * stride ldx, `repeat` > 1 reads row i/repeat (view directions shared by a ray's samples).
 * ---------------------------------------------------------------------------------------- */
int cnerf_posenc(const float* x, int ldx, int n, int C, int n_freqs, int repeat,
                 float* out, int ldo, int col0, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3 generic fp32 layer kernels (any D / W / skips; CUDA cores) -- nn.Linear + F.relu of
 * NeRF.forward NP/run_nerf_helpers.py:107-130.
 * ---------------------------------------------------------------------------------------- */

/* y[m,n] = act(x[m,k] w[n,k]^T + b[n]);  relu != 0 applies max(.,0). */
int cnerf_linear_fwd(const float* x, int ldx, const float* w, const float* b, int m, int n, int k,
                     int relu, float* y, int ldy, void* stream);
/* dx[m,k] (+)= (dy[m,n] * [y>0]) w[n,k];  y (may be NULL) is the layer's relu output. */
int cnerf_linear_bwd_data(const float* dy, int lddy, const float* y, int ldy, const float* w,
                          int m, int n, int k, float* dx, int lddx, int accumulate, void* stream);
/* dw[n,k] (+)= (dy*[y>0])^T x ; db[n] (+)= column sums.  Deterministic two-pass split over m;
 * workspace must hold cnerf_linear_bwd_weight_workspace(m,n,k) bytes. */
int64_t cnerf_linear_bwd_weight_workspace(int m, int n, int k);
int cnerf_linear_bwd_weight(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx,
                            int m, int n, int k, float* dw, float* db, int accumulate,
                            void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2+K3 fused: positional encoding + the 8x256 NeRF MLP on tcgen05 tensor cores
 * (run_network NP/run_nerf.py:37-52 + NeRF.forward).  Canonical architecture only:
 * D=8, W=256, skips=[4], use_viewdirs, multires=10, multires_views=4.
 *
 * Precision of the forward (inference and training), chosen by the caller with fwd_terms:
 *   fwd_terms    3: every fp32 operand is split into two fp16 terms (hi + lo); each product is
 *                   evaluated as hi*hi + hi*lo + lo*hi on the tensor cores with fp32 accumulation
 *                   in TMEM, i.e. ~2^-21 relative per product -- fp32-equivalent, 3 MMAs per
 *                   algorithmic MAC (measured: rendered maps within 7e-7 of the fp64 oracle);
 *                1: fp16 operands, ONE MMA per MAC, fp32 accumulation; the A operand is half as
 *                   large, so two 128-point tiles are in flight per SM (measured on workload A:
 *                   rendered maps within 2e-5 of the fp64 oracle -- bar 1e-4 -- raw within 1.1e-4).
 *                   In training it writes an fp16 record, i.e. it requires dw_terms == 1.
 *
 * Precision of the backward (K3b) is chosen by the caller with two arguments:
 *   chain_terms  3: the data-gradient chain G_{l-1} = (G_l W_l)[h>0] runs the same three-term split;
 *                1: fp16 operands, ONE MMA per MAC (fp32 accumulation), only the hi halves of the
 *                   weight blocks are fetched;
 *   dw_terms     3: dW_l = G_l^T X_l from fp16 hi + lo records of G and X (4 B per element, 3 MMAs);
 *                1: from the hi halves only (2 B per element -- half the record traffic of the
 *                   forward, the chain and the weight-gradient stage -- and 1 MMA).
 * The same dw_terms must be passed to cnerf_mlp_fwd_train and to every backward stage that reads
 * its record; (chain_terms, dw_terms) = (1, 3) is rejected.  Rounding is to nearest (unbiased):
 * per-element relative error 2^-11 in the one-term modes, averaged over the points of the batch.
 * ---------------------------------------------------------------------------------------- */
typedef struct cnerf_weights cnerf_weights;   /* opaque: packed fp16 hi/lo weight stream of ONE NeRF */

int cnerf_weights_create(cnerf_weights** out);
void cnerf_weights_destroy(cnerf_weights* w);
/* (Re)pack from live fp32 nn.Linear storage; call after every optimizer step.  All device
 * pointers, [out,in] row-major: pts_w[8]/pts_b[8], feature, alpha, views, rgb. */
int cnerf_weights_refresh(cnerf_weights* w, const float* const* pts_w, const float* const* pts_b,
                          const float* feature_w, const float* feature_b, const float* alpha_w,
                          const float* alpha_b, const float* views_w, const float* views_b,
                          const float* rgb_w, const float* rgb_b, void* stream);
/* raw[n_rays*n_samples, 4] = NeRF(embed(pts), embed(viewdirs[ray])). */
int cnerf_mlp_fwd(const cnerf_weights* w, const float* pts, const float* viewdirs, int n_rays,
                  int n_samples, float* raw, int fwd_terms, void* stream);
/* Training-mode forward: same result as cnerf_mlp_fwd, and every layer's A operand (encodings, post-activation
 * outputs; fp16 hi tiles, plus the lo tiles when dw_terms == 3) plus the ReLU sign bits are streamed into `acts`
 * (cnerf_mlp_acts_bytes(n_rays*n_samples) bytes, device: 1 345 536 per 128 points, opaque to the caller) for
 * cnerf_mlp_bwd.  With dw_terms == 1 the lo slots of the record are left untouched (never read). */
int64_t cnerf_mlp_acts_bytes(int64_t n_points);
int cnerf_mlp_fwd_train(const cnerf_weights* w, const float* pts, const float* viewdirs, int n_rays, int n_samples,
                        float* raw, void* acts, int fwd_terms, int dw_terms, void* stream);
/* K3b backward on tensor cores (autograd of run_network w.r.t. the parameters; loss.backward() of
 * NP/run_nerf_view.py:1982 for this module).  d_raw [n_points,4]; `acts` from cnerf_mlp_fwd_train with the SAME packed
 * weights; `grads_rec` scratch of cnerf_mlp_grads_bytes(n_points) bytes; `workspace` of
 * cnerf_mlp_bwd_workspace_bytes() bytes.  Outputs are the gradients of the twelve nn.Linear layers in their
 * [out,in] / [out] shapes (overwritten, or added to when accumulate != 0).  Deterministic. */
int64_t cnerf_mlp_grads_bytes(int64_t n_points);
int64_t cnerf_mlp_bwd_workspace_bytes(void);
int cnerf_mlp_bwd(const cnerf_weights* w, const float* d_raw, const void* acts, void* grads_rec, int n_points,
                  float* const* d_pts_w, float* const* d_pts_b, float* d_feature_w, float* d_feature_b,
                  float* d_alpha_w, float* d_alpha_b, float* d_views_w, float* d_views_b, float* d_rgb_w,
                  float* d_rgb_b, int accumulate, int chain_terms, int dw_terms, void* workspace, void* stream);
/* The three stages of cnerf_mlp_bwd as separate calls (same buffers and workspace; cnerf_mlp_bwd_data first):
 * the data-gradient chain, the weight/bias gradients of the ten GEMM layers, the two narrow heads. */
int cnerf_mlp_bwd_data(const cnerf_weights* w, const float* d_raw, const void* acts, void* grads_rec, int n_points,
                       int chain_terms, int dw_terms, void* workspace, void* stream);
int cnerf_mlp_bwd_weights(const void* acts, const void* grads_rec, int n_points, float* const* d_pts_w,
                          float* const* d_pts_b, float* d_feature_w, float* d_feature_b, float* d_views_w,
                          float* d_views_b, int accumulate, int dw_terms, void* workspace, void* stream);
int cnerf_mlp_bwd_heads(const float* d_raw, const void* acts, int n_points, float* d_alpha_w, float* d_alpha_b,
                        float* d_rgb_w, float* d_rgb_b, int accumulate, int dw_terms, void* workspace, void* stream);
/* ------------------------------------------------------------------------------------------
 * K4 alpha compositing -- raw2outputs NP/run_nerf.py:265-308 (depth_map as returned by
 * NP/run_nerf_view.py:392-439).
 * ---------------------------------------------------------------------------------------- */

/* raw [n,S,4], z [n,S], rays_d rows at d + i*d_stride, noise [n,S] already scaled (or NULL).
 * Outputs: rgb [n,3], disp [n], acc [n], depth [n], weights [n,S] (weights may be NULL). */
int cnerf_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                        const float* noise, int n_rays, int n_samples, int white_bkgd,
                        float* rgb, float* disp, float* acc, float* depth, float* weights, void* stream);
/* d_raw [n,S,4] from the output gradients (any of g_disp/g_acc/g_depth/g_weights may be NULL). */
int cnerf_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                        const float* noise, int n_rays, int n_samples, int white_bkgd,
                        const float* g_rgb, const float* g_disp, const float* g_acc, const float* g_depth,
                        const float* g_weights, float* d_raw, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5 hierarchical sampling -- sample_pdf NP/run_nerf_helpers.py:206-250 and the
 * detach/sort/std around it in render_rays NP/run_nerf.py:394-399,415.
 * ---------------------------------------------------------------------------------------- */

/* bins [n,B], weights [n,B-1], u [n,M] (NULL = det: linspace(0,1,M) taken from u_det[M]).
 * samples [n,M]; optional debug outputs cdf [n,B], below/above int32 [n,M] (may be NULL). */
int cnerf_sample_pdf(const float* bins, const float* weights, const float* u, const float* u_det,
                     int n_rays, int n_bins, int n_new, float* samples, float* cdf,
                     int32_t* below, int32_t* above, void* stream);
/* Fused fine-sample generation: bins = mid(z), weights[...,1:-1], inverse CDF, merge-sort with z,
 * population std of the new samples.  z/weights [n,S]; outputs z_samples [n,M], z_fine [n,S+M],
 * z_std [n] (any may be NULL except z_fine). */
int cnerf_sample_fine(const float* z, const float* weights, const float* u, const float* u_det,
                      int n_rays, int n_samples, int n_new, float* z_samples, float* z_fine,
                      float* z_std, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6 cross-view warp / gather / occlusion test -- get_ref_rays + get_test_label
 * NP/run_nerf_view.py:576-669 and the hard-mask loop :1014-1041.
 * ---------------------------------------------------------------------------------------- */

/* World points [R,3] -> rounded reference pixel (px,py as floats, torch.round semantics), strict
 * in-bounds mask, camera-space point [R,3]; optional integer gathers from img [C,H,W] / depth [H,W]
 * (zero where masked out) and the reference-view rays through the rounded pixel (c2w_host 3x4 or NULL).
 * w2c_host 4x4 (or 3x4) row-major, K_host 3x3: HOST pointers. */
int cnerf_project_gather(const float* pts_w, int n, const float* w2c_host, const float* K_host,
                         const float* c2w_host, const float* img, int C, const float* depth, int H, int W,
                         float* px, float* py, uint8_t* mask, float* cam, float* rgb_ref, float* depth_ref,
                         float* ref_rays_o, float* ref_rays_d, void* stream);
/* One (target, reference) pair of the hard-mask precompute: back-project target pixels with their
 * prior depth, project, compare against the reference prior depth with a threshold that doubles per
 * `chunk` pixels until the chunk has a hit.  mask [n] uint8 is OR-ed into (accumulate) or overwritten. */
int cnerf_hard_mask_pair(const float* rays_o, const float* rays_d, const float* depth_tgt, int n,
                         const float* w2c_host, const float* K_host, const float* depth_ref, int H, int W,
                         float thr0, int chunk, int accumulate, uint8_t* mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7 masked consistency losses -- NP/run_nerf_view.py:1645-1648,1737 and
 * NP/run_nerf_view_cal_correspondance.py:1516-1517,1550-1551.
 * ---------------------------------------------------------------------------------------- */

/* e = (pred/divisor - target/divisor)^2  (divisor = far for the depth term, 1 for rgb).
 * out[0] = mean_{mask==1}(e) + [use_unmasked && sum(mask) != n_ref] coef * mean_{mask==0}(e);
 * out[1] = #rows mask==1, out[2] = #rows mask==0, out[3] = plain mean over all rows (img2mse),
 * out[4] = sum(mask).  pred/target [n,C], mask [n] (NULL = all ones), out [5] floats.
 * global_counts (device, 4 floats, or NULL): {#mask==1, #mask==0, sum(mask), #rows} of the GLOBAL batch when the rows
 * are one rank's shard of it: the means are then taken over the global batch -- out[0] is this
 * rank's additive share of the single-GPU loss (the ranks' shares sum to it, and so do the gradients), out[1], out[2],
 * out[4] repeat the global counts, and n_ref must be the global reference count.
 * workspace: 8192 bytes of scratch (needs no initialisation).  Deterministic reduction order. */
int cnerf_masked_mse_fwd(const float* pred, const float* target, const float* mask, int n, int C,
                         float divisor, float coef, float n_ref, int use_unmasked, const float* global_counts,
                         float* out, void* workspace, void* stream);
/* d_pred [n,C] = g_loss[0] * d out[0] / d pred, using the counts in `out` from the forward. */
int cnerf_masked_mse_bwd(const float* pred, const float* target, const float* mask, int n, int C,
                         float divisor, float coef, float n_ref, int use_unmasked, const float* out,
                         const float* g_loss, float* d_pred, void* stream);

/* K7b soft-weighted MSE over n_elems values -- replaces img2mse_softmask / img2mse_depth_softmask (kind 0, param = temp)
 * NP/run_nerf_view.py:50,55 and img2mse_softLpmask (kind 1, param = coef) NP/run_nerf_view.py:58:
 *   d = pred / divisor - target / divisor,   w = exp(d^2 / temp)  |  |d|^coef + 1,   loss = sum(w d^2) / sum(w)
 * with sum(w) a constant for d (the reference detaches it) but, for kind 0, a function of temp.  The parameter is read from
 * param_dev[0] when that device pointer is non-NULL (temp = softplus(network_fine.temp_rgb) is a device scalar in the
 * reference, :1659) and from `param` otherwise.  out[5] = {loss, sum(w d^2), sum(w), d loss / d param (0 for kind 1), param}.
 * workspace: 8192 bytes of scratch.  Deterministic reduction order.  n_elems == 0 gives NaN like the reference (0 / 0). */
int cnerf_soft_mse_fwd(const float* pred, const float* target, int64_t n_elems, float divisor, int kind, float param,
                       const float* param_dev, float* out, void* workspace, void* stream);
/* d_pred [n_elems] = g_loss[0] * d out[0] / d pred, using sum(w) and the parameter stored in `out` by the forward. */
int cnerf_soft_mse_bwd(const float* pred, const float* target, int64_t n_elems, float divisor, int kind, const float* out,
                       const float* g_loss, float* d_pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNERF_H_ */
