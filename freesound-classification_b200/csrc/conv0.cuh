// Block-0 entry convolution of the 2D network (networks/classifiers.py:524-532 for k = 0):
// BatchNorm2d(2) -> Conv2d(2 -> C0, 3x3, pad 1) -> MaxPool2d(2) fused into one CUDA-core kernel.
// K = 18 makes this layer bandwidth-bound, so it is a direct convolution that never materialises
// the un-pooled (N, C0, 128, frames) tensor; channel 1 (the frequency encoding,
// networks/classifiers.py:553-561) is synthesised analytically instead of being stored.
#pragma once
#include "common.cuh"

namespace fsb {

// sum / sum^2 of a contiguous float32 array -> channel 0 of a [nblk][2][16] double partial record
int plain_stats_blocks();
int plain_stats(const float* x, long long n, double* partials16, cudaStream_t s);

// feat (N, H, W) float32 ; scale/shift: BN_in coefficients (>= 2 entries) ; w (C0, 2, 3, 3), b (C0)
// zp: padded-flat pooled output, geometry gp = (N, H/2, W/2, C0)
// amax (optional, gp.rows * gp.Cs bytes): position 0..3 of the first maximum of every pool window (scan order), kept for
// the backward pass
int conv0_forward(const float* feat, int N, int H, int W, const float* scale, const float* shift, const float* w,
                  const float* b, float* zp, unsigned char* amax, const Geo& gp, cudaStream_t s);

// the same layer on the tensor cores (conv0_tc.cu): im2col operand built in shared memory, four TMEM accumulators = the
// four pool-window positions; precision 1 = three products, 2 = single pass.  Supported when Cs <= 128.
bool conv0_tc_supported(const Geo& gp);
int conv0_tc_forward(int precision, const float* feat, int N, int H, int W, const float* scale, const float* shift,
                     const float* w, const float* b, float* zp, unsigned char* amax, const Geo& gp, cudaStream_t s);

// weight / BatchNorm-input gradients on the tensor cores: dW-like reduction over the pooled pixels with the gradient
// routed by the stored arg-max (four masked gradient tiles x four patch tiles per 64 pixels, one TMEM accumulator per CTA);
// dz_absmax: GradScale slot of dzp (common.cuh), may be null.  Needs scale[0], scale[1] != 0 (checked on the device: the
// result is then NaN-free but the BatchNorm-input terms are computed by the CUDA-core kernel instead -- see net.cu).
int conv0_tc_backward(int precision, const float* feat, int N, int H, int W, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const float* w, const void* dzp, int dzp_half,
                      const unsigned char* amax, const unsigned* dz_absmax, const Geo& gp, float* dw, float* db,
                      float* dgamma_in, float* dbeta_in, void* scratch, cudaStream_t s);
// fixed-order sum of `nblk` per-CTA records [22][Cs] (18 dW taps, 2 d(gamma_in) terms, 2 d(beta_in) terms per channel)
int conv0_bwd_finalize(const float* partials, int nblk, const Geo& gp, float* dw, float* db, float* dgamma_in,
                       float* dbeta_in, cudaStream_t s);

int conv0_bwd_blocks();
size_t conv0_bwd_scratch_bytes(const Geo& gp);
// dzp: float32 padded-flat gradient wrt the pooled conv output.  Writes dw (C0,2,3,3), db (C0, zeros:
// the bias feeds a batch-statistics BN), dgamma_in / dbeta_in (2 entries each).
int conv0_backward(const float* feat, int N, int H, int W, const float* scale, const float* shift,
                   const float* mean, const float* invstd, const float* w, const float* b, const float* dzp,
                   const unsigned char* amax, const Geo& gp, float* dw, float* db, float* dgamma_in, float* dbeta_in,
                   void* scratch, cudaStream_t s);

}  // namespace fsb
