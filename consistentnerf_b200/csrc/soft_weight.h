/* Per-element arithmetic of the soft-weighted MSE losses (K7b), shared by the CUDA kernels in crossview.cu and by a
 * host build of the same functions (tests/test_soft_weight_host.py compiles this header with gcc and checks it against
 * the reference's lambdas under torch autograd, so the math is pinned without a GPU).
 *
 *   kind 0, NP/run_nerf_view.py:50,55   img2mse_softmask / img2mse_depth_softmask(x, y, temp):
 *           sum(exp(d^2 / temp) * d^2) / sum(exp(d.detach()^2 / temp)),            d = x - y
 *   kind 1, NP/run_nerf_view.py:58      img2mse_softLpmask(x, y, coef):
 *           sum((|d|^coef + 1) * d^2) / sum(|d|^coef + 1).detach()
 *
 * Only the NUMERATOR's weight carries a gradient to d in both; the denominator is a constant for d, and (kind 0) a
 * function of temp. */
#ifndef CNERF_SOFT_WEIGHT_H_
#define CNERF_SOFT_WEIGHT_H_

#include <math.h>

#ifdef __CUDACC__
#define CNERF_HD __host__ __device__ __forceinline__
#else
#define CNERF_HD static inline
#endif

/* the residual the losses see: both arguments are scaled first (depth_pred / far, depth_cas / far), then subtracted */
CNERF_HD float cnerf_soft_residual(float pred, float target, float divisor) { return pred / divisor - target / divisor; }

/* weight of one element */
CNERF_HD float cnerf_soft_weight(float d, int kind, float param) {
    return kind == 0 ? expf(d * d / param) : powf(fabsf(d), param) + 1.f;
}

/* d (w(d) * d^2) / d d */
CNERF_HD float cnerf_soft_dnum(float d, int kind, float param) {
    if (kind == 0) return expf(d * d / param) * (2.f * d + 2.f * d * d * d / param);
    return ((param + 2.f) * powf(fabsf(d), param) + 2.f) * d;
}

/* loss and d loss / d temp from the three sums  num = sum w d^2,  den = sum w,  s4 = sum w d^4  (kind 0:
 * d num / d temp = -s4 / temp^2,  d den / d temp = -num / temp^2;  kind 1 has no learnable parameter) */
CNERF_HD void cnerf_soft_finish(double num, double den, double s4, int kind, double param, double* loss, double* dparam) {
    *loss = num / den;
    *dparam = kind == 0 ? (-s4 / den + num * num / (den * den)) / (param * param) : 0.0;
}

#endif /* CNERF_SOFT_WEIGHT_H_ */
