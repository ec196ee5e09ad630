// Weight-block program of the N=256 forward kernels (mlp_fwd3.cu three-term, mlp_fwd5.cu fp16 two-tile, experiments/: CTA pairs) and the epilogue /
// encoding helpers they share.
#pragma once
#include "mlp_layout.cuh"

namespace cnerf {

constexpr int k3NumBlocks = 4 + 17 * 4 + 20 + 17 * 3 + 9;         // 152 (one bias block for every layer without an encoding block)

// ------------------------------------------------------------------------------------
// weight stream: blocks in consumption order
//   layers 0-8: [256 rows x 16 k] blocks, element (n, k) at (k/8)*4096 + n*16 + (k%8)*2 (hi), +8192 (lo)
//   layer 9   : [128 rows x 32 k] blocks, element (n, k) at (k/8)*2048 + n*16 + (k%8)*2 (hi), +8192 (lo)
// Biases ride on the tensor core: the padding column of the encoding tile (column 63 of the point encoding, column 31
// of the direction encoding) holds 1.0 and the matching weight column holds the bias.  Layers whose input has no
// encoding part get one extra block that multiplies encoding columns 48-63 with [0 ... 0, bias].
// ------------------------------------------------------------------------------------
struct Blk3 { int layer, src_k0, kvalid, bias_k; };       // bias_k: column of the block that carries the bias (-1: none)

__device__ __forceinline__ Blk3 block3_info(int b) {
    // layer 0: 4 blocks; 1-4: 16 + bias; 5: 4 + 16; 6-8: 16 + bias; 9: 8 + 1
    if (b < 4) return {0, 16 * b, b == 3 ? 15 : 16, b == 3 ? 15 : -1};
    b -= 4;
    if (b < 68) { int l = 1 + b / 17, j = b % 17; return j < 16 ? Blk3{l, 16 * j, 16, -1} : Blk3{l, 0, 0, 15}; }
    b -= 68;
    if (b < 4) return {5, 16 * b, b == 3 ? 15 : 16, b == 3 ? 15 : -1};
    b -= 4;
    if (b < 16) return {5, 63 + 16 * b, 16, -1};
    b -= 16;
    if (b < 51) { int l = 6 + b / 17, j = b % 17; return j < 16 ? Blk3{l, 16 * j, 16, -1} : Blk3{l, 0, 0, 15}; }
    b -= 51;
    if (b < 8) return {9, 32 * b, 32, -1};
    return {9, 256, 27, 31};
}

// ------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void st_global_v4_(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
// 8 fp32 of one k-group -> hi/lo words into the SMEM operand tile and (training) the global record
template <bool kSave>
__device__ __forceinline__ void emit_kgroup(uint32_t hi_base, uint32_t lo_base, uint8_t* rec_hi, size_t lo_off, uint32_t row,
                                            uint32_t kg, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    const uint32_t off = kg * kLBO + row * 16;
    st_shared_v4(hi_base + off, h[0], h[1], h[2], h[3]);
    st_shared_v4(lo_base + off, l[0], l[1], l[2], l[3]);
    if (kSave) {
        st_global_v4_(rec_hi + off, h[0], h[1], h[2], h[3]);
        st_global_v4_(rec_hi + lo_off + off, l[0], l[1], l[2], l[3]);
    }
}

// 8 fp32 of one k-group -> hi/lo words straight into the global record
__device__ __forceinline__ void emit_record(uint8_t* rec_hi, size_t lo_off, uint32_t row, uint32_t kg, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    const uint32_t off = kg * kLBO + row * 16;
    st_global_v4_(rec_hi + off, h[0], h[1], h[2], h[3]);
    st_global_v4_(rec_hi + lo_off + off, l[0], l[1], l[2], l[3]);
}

// bit j = [sign bit of w[j] clear] for eight floats: the top bytes are gathered with PRMT and the four sign bits of a word
// are compacted with one multiply (y * (2^24 + 2^17 + 2^10 + 2^3) puts bit 8 i of y at bit 24 + i, no two terms collide)
__device__ __forceinline__ uint32_t sign_clear_bits8(const float* w) {
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t a = __byte_perm(__float_as_uint(w[4 * q]), __float_as_uint(w[4 * q + 1]), 0x0073);
        const uint32_t b = __byte_perm(__float_as_uint(w[4 * q + 2]), __float_as_uint(w[4 * q + 3]), 0x0073);
        const uint32_t x = __byte_perm(a, b, 0x5410);
        const uint32_t y = (~x >> 7) & 0x01010101u;
        r |= ((y * 0x01020408u) >> 24) << (4 * q);
    }
    return r;
}

template <int J>
__device__ __forceinline__ float enc_col3(const float (&x)[3], int width) {
    if (J >= width) return 0.f;
    if (J < 3) return x[J];
    constexpr int b = (J - 3) / 3, c = (J - 3) % 3, oct = b / 2;
    float arg = x[c] * (float)(1 << oct);
    return (b & 1) ? cosf(arg) : sinf(arg);
}
template <int J0>
__device__ __forceinline__ void enc8(const float (&x)[3], int width, float* v) {
    v[0] = enc_col3<J0 + 0>(x, width); v[1] = enc_col3<J0 + 1>(x, width); v[2] = enc_col3<J0 + 2>(x, width);
    v[3] = enc_col3<J0 + 3>(x, width); v[4] = enc_col3<J0 + 4>(x, width); v[5] = enc_col3<J0 + 5>(x, width);
    v[6] = enc_col3<J0 + 6>(x, width); v[7] = enc_col3<J0 + 7>(x, width);
}


}  // namespace cnerf
