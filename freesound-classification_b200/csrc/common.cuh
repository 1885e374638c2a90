// Shared definitions for libfsb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fsb200.h"

namespace fsb {

// ---------------------------------------------------------------------------------------------
// error plumbing: every C-ABI function returns an int and records text for fsb_last_error()
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern long long g_launch_count;

#define FSB_CUDA(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            fsb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                           cudaGetErrorString(_e));                                     \
            return (int)_e;                                                             \
        }                                                                               \
    } while (0)

#define FSB_TRY(expr)                                                                   \
    do {                                                                                \
        int _r = (expr);                                                                \
        if (_r != 0) return _r;                                                         \
    } while (0)

#define FSB_REQUIRE(cond, ...)                                                          \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            fsb::set_error(__VA_ARGS__);                                                \
            return FSB_E_INVALID;                                                       \
        }                                                                               \
    } while (0)

// count + check a kernel launch (cudaGetLastError only reports launch-configuration errors; it
// does not synchronise)
#define FSB_LAUNCHED()                                                                  \
    do {                                                                                \
        ++fsb::g_launch_count;                                                          \
        FSB_CUDA(cudaGetLastError());                                                   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Padded-flat NHWC geometry ("PF"): a tensor of N images, H x W pixels, C channels is stored as
// rows = N * Hp * Wp pixels of Cs channels each (channels innermost, Cs = C rounded up to 16),
// with a zero border of padH rows / padW columns around every image.  A 3x3 "same" convolution
// then reads, for output row r and tap (dy,dx), input row r + (dy-1)*Wp + (dx-1): every conv is a
// sum of row-shifted GEMMs and the zero border supplies the padding.
// ---------------------------------------------------------------------------------------------
struct Geo {
    int N, H, W, C, Cs, padH, padW, Hp, Wp;
    long long rows;      // N * Hp * Wp
    long long pixels;    // N * H * W (interior)
    const unsigned char* mask;   // device: mask[row] = 1 for interior rows; nullptr = no border (all interior)
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

inline Geo make_geo(int N, int H, int W, int C, int padH, int padW) {
    Geo g;
    g.N = N; g.H = H; g.W = W; g.C = C; g.Cs = round_up(C, 16);
    g.padH = padH; g.padW = padW; g.Hp = H + 2 * padH; g.Wp = W + 2 * padW;
    g.rows = (long long)N * g.Hp * g.Wp;
    g.pixels = (long long)N * H * W;
    g.mask = nullptr;
    return g;
}

__host__ __device__ inline long long geo_row(const Geo& g, int n, int y, int x) {
    return ((long long)n * g.Hp + (y + g.padH)) * g.Wp + (x + g.padW);
}

// interior pixel index q in [0, N*H*W) -> padded row (32-bit divisions: N*H*W < 2^32 is checked on the host)
__device__ __forceinline__ long long geo_q_to_row(const Geo& g, long long q) {
    const unsigned uq = (unsigned)q;
    const unsigned t = uq / (unsigned)g.W;
    const unsigned x = uq - t * (unsigned)g.W;
    const unsigned n = t / (unsigned)g.H;
    const unsigned y = t - n * (unsigned)g.H;
    return ((long long)n * g.Hp + (y + g.padH)) * g.Wp + (x + g.padW);
}

// ---------------------------------------------------------------------------------------------
// activation storage formats consumed by the GEMM kernels
//   FMT_F32   : one float32 plane                          (precision 0, CUDA-core GEMM)
//   FMT_H16X2 : two IEEE-half planes hi, lo with x ~= hi+lo (three-product tcgen05 GEMM: lo*hi + hi*lo + hi*hi keeps
//               ~2^-22 per product -- float32-grade)
//   FMT_H16   : the hi plane only                           (single-pass tcgen05 GEMM, ~2^-12 per operand)
// A "plane" is rows*Cs elements; the lo plane follows the hi plane (buffers are always sized for both).
// Half (11-bit significand) rather than bfloat16 (8-bit): the single-pass backward of the mixed mode then keeps
// 2^-12 per operand instead of 2^-9.  The price is range: activations are clamped to +-65504 (post-BatchNorm values
// are O(1)), and gradient planes are written with a per-tensor power-of-two scale (GradScale below).
// ---------------------------------------------------------------------------------------------
enum { FMT_F32 = 0, FMT_H16X2 = 1, FMT_H16 = 2 };

__device__ __forceinline__ void split_h16(float x, __half& hi, __half& lo) {
    const float c = fminf(fmaxf(x, -65504.f), 65504.f);
    x = (x == x) ? c : x;                       // NaN stays NaN (a diverged run must not be masked)
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// Per-tensor gradient scale.  `absmax_bits` holds the float32 bit pattern of an upper bound B >= max |dz| of the
// tensor (0 = unknown / not used).  Writers multiply by 2^(13 - floor(log2 B)) so that the largest magnitude lands in
// [2^13, 2^14) at most -- a factor 4 below the half maximum, ~28 binades above the smallest normal half -- and the
// consumers (dgrad epilogue, wgrad finalize) multiply by the inverse.  Powers of two: scaling is exact and the
// backward pass stays exactly linear in the incoming gradient.
__host__ __device__ __forceinline__ int gs_exponent(unsigned absmax_bits) {
    const int e = (int)((absmax_bits >> 23) & 0xFFu);          // biased exponent of B
    if (absmax_bits == 0u || e == 0 || e == 255) return 0;      // unknown, denormal or non-finite bound: no scaling
    int k = 13 - (e - 127);
    return k < -100 ? -100 : (k > 100 ? 100 : k);
}
__device__ __forceinline__ float gs_scale(const unsigned* absmax_bits) {
    return absmax_bits ? __uint_as_float((unsigned)(127 + gs_exponent(*absmax_bits)) << 23) : 1.f;
}
__device__ __forceinline__ float gs_inv_scale(const unsigned* absmax_bits) {
    return absmax_bits ? __uint_as_float((unsigned)(127 - gs_exponent(*absmax_bits)) << 23) : 1.f;
}

// GradScale of a tensor whose bound is the PRODUCT of a stored bound and a float multiplier (dgrad outputs of the compact
// backward: |dA| <= max|dZ| * max_ci sum_{co,t} |W|, both already on the device): writer and readers derive the same
// exponent from the same two device words, so no extra slot or launch is needed.
__device__ __forceinline__ int gs_exponent2(const unsigned* absmax_bits, const float* mul) {
    if (!absmax_bits) return 0;
    unsigned b = *absmax_bits;
    if (mul) b = __float_as_uint(__uint_as_float(b) * (*mul));
    return gs_exponent(b);
}
__device__ __forceinline__ float gs_pow2(int k) {          // 2^k, k clamped to the normal float32 range
    k = k < -126 ? -126 : (k > 127 ? 127 : k);
    return __uint_as_float((unsigned)(127 + k) << 23);
}

// A gradient tensor as the element-wise backward kernels read it: a float32 plane, or (compact backward of the mixed
// mode) ONE half plane holding 2^k * gradient with k = gs_exponent2(bits, mul).
struct GradRef {
    const void* p;
    int half;                // 1 = half plane
    const unsigned* bits;    // half: bound slot (GradScale)
    const float* mul;        // half, optional: multiplier of the bound
};
inline GradRef grad_f32(const float* p) { return GradRef{p, 0, nullptr, nullptr}; }
inline GradRef grad_h16(const void* p, const unsigned* bits, const float* mul = nullptr) { return GradRef{p, 1, bits, mul}; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace fsb
