// Layout shared by the tensor-core MLP kernels: shared-memory map, activation-record slots, the fp32
// side-parameter block and the weight handle.
#pragma once
#include "umma.cuh"

namespace cnerf {

constexpr int kNumLayers = 10;                  // 0-7 pts_linears, 8 feature_linear, 9 views_linears.0
// Copies of the packed weight streams (CTA b reads copy b % kWeightReplicas).  Measured with 8 copies (profiles/r2e_*): the wait for
// weight blocks did not change (28.3k cycles per tile in mlp_fwd5 either way), i.e. the 148 SMs streaming the same blocks at the same
// time is NOT an L2 hot-spot problem, and the larger L2 footprint cost the three-term kernels 5-10 %.  Hence one copy.
constexpr int kWeightReplicas = 1;
constexpr uint32_t kBlockBytes = 16384;
constexpr uint32_t kBlockHalfBytes = 8192;

// misc fp32 parameter block
constexpr int kMiscBias = 0;                    // [10][256]
constexpr int kMiscAlphaW = 2560;               // [256]
constexpr int kMiscAlphaB = 2816;               // [1]
constexpr int kMiscRgbW = 2820;                 // [3][128]
constexpr int kMiscRgbB = 3204;                 // [3]
constexpr int kMiscFloats = 3208;

struct RawParams {
    const float* w[kNumLayers];
    const float* b[kNumLayers];
    const float* alpha_w; const float* alpha_b; const float* rgb_w; const float* rgb_b;
    int ld[kNumLayers];
};

// ------------------------------------------------------------------------------------
// activation record (training): one record per 128-point tile
// ------------------------------------------------------------------------------------
constexpr size_t kSlotE = 0, kSlotH0 = 32768, kSlotF = kSlotH0 + 8 * 131072, kSlotV = kSlotF + 131072,
                 kSlotHV = kSlotV + 32768, kSlotM = kSlotHV + 65536, kTileBytes = kSlotM + 34816;      // 1 345 536 B per 128 points
// kSlotM: ReLU sign bits, one byte per (row, k-group): bit j = [feature 8 kg + j > 0].  Layout follows the epilogue thread
// mapping (thread = (row, p), k-groups kb * 4 + p for kb = 0..7) so that a thread stores / loads its eight bytes of a layer
// with one 8-byte access:  H_l (l = 0..7): l * 4096 + p * 1024 + row * 8 + kb;  views-layer output: 32768 + row * 16 + kg.
// The data-gradient chain reads these 34 KB per tile instead of the 544 KB of hi halves it used to scan for the same bits.


// Training mode: every A operand (encodings and post-activation layer outputs, fp16 hi/lo in the UMMA
// layout) is also streamed to HBM, one record per 128-point tile; the backward kernels (mlp_bwd_tc.cu)
// read ReLU masks and dW operands from it.  Slot = [hi | lo], each k-group 2048 B (128 rows x 16 B).
//   E  point encoding (K=64)   H0..H7 pts_linears outputs (K=256)   F feature_linear output (K=256)
//   V  direction encoding (K=32 used, stored as the 64-wide buffer)  HV views_linears output (K=128)   M  ReLU sign bits

}  // namespace cnerf

struct cnerf_weights {
    uint8_t* stream3 = nullptr;     // forward stream (mlp_fwd3.cu): [256 x 16] blocks, hi 8 KB | lo 8 KB, in consumption order
    uint8_t* stream_bwd3 = nullptr; // data-gradient chain stream (mlp_bwd_tc.cu): [256 x 16] transposed blocks
    uint8_t* stream4 = nullptr;     // experiments build: forward stream of the 64-row CTA-pair kernel (experiments/mlp_fwd4.cu)
    uint8_t* stream6 = nullptr;     // fp16 forward stream of the M = 256 CTA-pair kernel (mlp_fwd6.cu): per unit two 4 KB halves (one per CTA)
    float* misc = nullptr;          // biases + alpha/rgb heads (fp32)
    int device = -1;
    bool packed = false;
};
