/*
 * cnerf_debug.h -- self-tests, microbenchmarks and profiling hooks of libcnerf.so.
 *
 * NOT part of the drop-in boundary (include/cnerf.h): nothing here is called by the product path.
 * The entry points exist so that the tcgen05 building blocks (descriptors, TMEM mapping, issue
 * rate) can be unit-tested and the fused kernels profiled from Python (tests/test_gpu_kernels.py,
 * scripts/prof_phases.py, scripts/umma_rate.py, scripts/mma_terms.py).  Same conventions as
 * cnerf.h (device pointers, explicit stream, int status).
 *
 * Section 1 is linked into every build of the library.  Section 2 exists only in builds made with
 * `python -m consistentnerf_b200.build --experiments` (-DCNERF_EXPERIMENTS): the CTA-pair
 * (cta_group::2) forward experiment csrc/experiments/mlp_fwd4.cu and its probes -- a measured
 * negative result (2.4x slower than the single-CTA kernel) kept for reference.
 */
#ifndef CNERF_DEBUG_H_
#define CNERF_DEBUG_H_

#include "cnerf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * 1. always built
 * ------------------------------------------------------------------------------------------- */

/* Unit self-test of the tcgen05 building blocks: d[128,n] = a[128,k] b[n,k]^T with the same
 * fp16 hi/lo split, descriptors and TMEM read-back the fused kernel uses (k%16==0, n%16==0, n<=256). */
int cnerf_umma_selftest(const float* a, const float* b, int n, int k, float* d, void* stream);
/* Same product with the A operand staged in tensor memory (tcgen05.st + TS-mode MMA); k <= 256. */
int cnerf_umma_selftest_ts(const float* a, const float* b, int n, int k, float* d, void* stream);

/* The K2+K3 forward (cnerf_mlp_fwd) with a SUBSET of the three partial products of the fp16 hi/lo split: terms bit 0
 * a_hi*w_hi (required), bit 1 a_hi*w_lo, bit 2 a_lo*w_hi.  Measurement only: the error / speed table of the reduced-MMA
 * variants in DESIGN.md section 3 (scripts/mma_terms.py); terms == 7 runs all three through the same instantiation. */
int cnerf_debug_mlp_fwd_terms(const cnerf_weights* w, const float* pts, const float* viewdirs, int n_rays,
                              int n_samples, float* raw, int terms, void* stream);

/* Enable/disable the in-kernel phase profile of the fused forward kernel (mlp_fwd3.cu) and read + clear its 16 cycle
 * counters (host pointer, may be NULL).  Synchronises the device. */
int cnerf_debug_profile3(int enable, unsigned long long* out16);
/* Same for the fp16 two-tile forward kernel (mlp_fwd5.cu). */
int cnerf_debug_profile5(int enable, unsigned long long* out16);
/* Same for the data-gradient chain kernel (mlp_bwd_tc.cu). */
int cnerf_debug_profile_chain(int enable, unsigned long long* out16);
/* Measured cycles per tcgen05.mma (M=128, N=n, K=16; mode 0 = SS, 1 = TS) on every SM; out: 148 device floats. */
int cnerf_debug_umma_rate(int mode, int n, int iters, int alt, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 2. experiment builds only (-DCNERF_EXPERIMENTS)
 * ------------------------------------------------------------------------------------------- */
#ifdef CNERF_EXPERIMENTS
/* CTA-pair variant (tcgen05 cta_group::2, one M=256 instruction stream for two SMs): d[256,n] = a[256,k] b[n,k]^T;
 * each CTA of the pair holds its 128 rows of a/d and n/2 rows of b.  n%32==0, n<=256, k<=128. */
int cnerf_umma_selftest_pair(const float* a, const float* b, int n, int k, float* d, void* stream);
/* Phase profile of the CTA-pair forward kernel (experiments/mlp_fwd4.cu). */
int cnerf_debug_profile4(int enable, unsigned long long* out16);
/* TMEM data layout of an M=128 cta_group::2 accumulator (64 rows of a[128,k] per CTA, fp16 hi parts only):
 * dump[2][128][256] = every CTA's TMEM window after d = a b^T. */
int cnerf_debug_pair_layout(const float* a, const float* b, int n, int k, float* dump, void* stream);
/* Cycles per N=256 K=16 SS tcgen05.mma, pair != 0: M=256 cta_group::2 on 74 CTA pairs, else M=128 on 148 CTAs; traffic
 * bit 0 adds concurrent st.shared traffic, bit 1 a bulk-copy ring fed from src (>= 1 MiB, device).  out: 148 device floats. */
int cnerf_debug_umma_rate_pair(int pair, int iters, int traffic, const void* src, float* out, void* stream);
#endif

#ifdef __cplusplus
}
#endif
#endif /* CNERF_DEBUG_H_ */
