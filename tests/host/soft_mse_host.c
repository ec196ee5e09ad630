/* TEST INFRASTRUCTURE: a serial host build of the two K7b entry points, from the same element arithmetic the CUDA kernels use
 * (consistentnerf_b200/csrc/soft_weight.h) and against the same prototypes (include/cnerf.h -- a mismatch is a compile error).
 * tests/test_soft_mse_host.py loads it in place of libcnerf.so for these two names only, to check the math and the autograd glue
 * on a machine without a GPU.  Never loaded by the package. */
#include <stddef.h>
#include "cnerf.h"
#include "soft_weight.h"

int cnerf_soft_mse_fwd(const float* pred, const float* target, int64_t n_elems, float divisor, int kind, float param,
                       const float* param_dev, float* out, void* workspace, void* stream) {
    (void)workspace; (void)stream;
    if (param_dev) param = param_dev[0];
    double num = 0.0, den = 0.0, s4 = 0.0, loss, dparam;
    for (int64_t i = 0; i < n_elems; ++i) {
        float d = cnerf_soft_residual(pred[i], target[i], divisor);
        double w = (double)cnerf_soft_weight(d, kind, param), e = (double)(d * d);
        num += w * e; den += w; s4 += w * e * e;
    }
    cnerf_soft_finish(num, den, s4, kind, (double)param, &loss, &dparam);
    out[0] = (float)loss; out[1] = (float)num; out[2] = (float)den; out[3] = (float)dparam; out[4] = param;
    return 0;
}

int cnerf_soft_mse_bwd(const float* pred, const float* target, int64_t n_elems, float divisor, int kind, const float* out,
                       const float* g_loss, float* d_pred, void* stream) {
    (void)stream;
    for (int64_t i = 0; i < n_elems; ++i) {
        float d = cnerf_soft_residual(pred[i], target[i], divisor);
        d_pred[i] = g_loss[0] * cnerf_soft_dnum(d, kind, out[4]) / out[2] / divisor;
    }
    return 0;
}
