// Element-wise / reduction kernels over padded-flat NHWC tensors: BatchNorm statistics, BN-apply +
// PReLU (+ residual, + dropout), max-pool, global max heads, and all their backward passes.
// These replace the reference's nn.BatchNorm{1,2}d / nn.PReLU / nn.MaxPool{1,2}d /
// nn.AdaptiveMaxPool{1,2}d / nn.Dropout call sites (networks/classifiers.py:524-549, :72-104) and
// their autograd mirrors.
#pragma once
#include "common.cuh"

namespace fsb {

struct BnCoef {            // per-channel vectors of length Cs (padded entries are zero)
    const float* scale;    // gamma * invstd
    const float* shift;    // beta - mean * scale
    const float* slope;    // PReLU slope, nullptr = no activation
    const float* mean;     // batch mean      (backward only)
    const float* invstd;   // 1/sqrt(var+eps) (backward only)
};

struct Residual {          // identity branch r = prelu(zr*scale+shift, slope); nullptr zr = none
    const float* zr;
    const float* scale;
    const float* shift;
    const float* slope;
};

struct Dropout {           // keep mask = hash(seed, element) >= p ; scale 1/(1-p); p == 0 = off
    float p;
    unsigned long long seed;
    const unsigned long long* seed_ptr;   // optional DEVICE location of the seed (overrides `seed`): lets a captured
                                          // CUDA graph pick up a fresh seed on every replay
};

// number of pixel-blocks an element-wise reduction over `g` uses (partials are [nblk][K][Cs] doubles)
int ew_num_blocks(const Geo& g);

// fills mask[row] = 1 for interior rows of g, 0 for border rows (g.rows bytes)
int pf_build_mask(const Geo& g, unsigned char* mask, cudaStream_t s);
int pf_zero_border(void* buf, int fmt, const Geo& g, cudaStream_t s);
int pf_zero_all(void* buf, int fmt, const Geo& g, cudaStream_t s);

// sum / sum of squares per channel over interior pixels -> partials [nblk][2][Cs] (double)
int pf_stats(const float* x, const Geo& g, double* partials, cudaStream_t s);

// Finalize BN statistics.  training: batch statistics from `partials` (count = interior pixels),
// running stats updated in place (momentum 0.1, unbiased variance), *bn_count += 1.
// eval: running stats.  gamma/beta/run_* have C entries; outputs have Cs entries (tail zeroed).
int bn_finalize(const double* partials, int nblk, long long count, const float* gamma,
                const float* beta, float* run_mean, float* run_var, long long* bn_count, int training,
                int C, int Cs, float* scale, float* shift, float* mean, float* invstd, cudaStream_t s);

// analytic statistics of the frequency-encoding channel linspace(-1, 1, H) (2D block 0, channel 1):
// appends channel 1 to a 2-channel partial record [1][2][16] (count = N*H*W)
int freq_encoding_stats(int H, long long n_times_w, double* partials16, cudaStream_t s);

// a = act(BN(z) [+ residual]) ; writes any of: GEMM-format plane(s) `a_mma` (fmt), float32 `a_f32`;
// out_stats (optional): per-block partial sum / sum of squares of `a` ([ew_num_blocks(g)][2][Cs] doubles),
// i.e. the batch statistics of the NEXT BatchNorm, gathered while the data is in registers
// post (optional, eval): the operand planes receive post(a) -- the NEXT BatchNorm (+ PReLU), a fixed affine map when it runs
// on running statistics -- while a_f32 still receives a
int bn_act_forward(const float* z, const Geo& g, BnCoef bn, Residual res, Dropout dr, void* a_mma,
                   int fmt, float* a_f32, double* out_stats, cudaStream_t s, const BnCoef* post = nullptr);

// 2x2 (pool_h = 2) or 1x2 (pool_h = 1) max pool, floor mode: zf (gf) -> zp (gp)
// out_stats (optional): partial sum / sum of squares of zp, [ew_num_blocks(gp)][2][Cs] doubles
// amax (optional): position 0..3 of the first maximum of every window, one byte per pooled element (gp.rows * Cs bytes)
// post + a_mma (optional, eval): additionally writes act(BN(zp)) of the BatchNorm that follows as GEMM operand planes (fmt)
int maxpool_forward(const float* zf, const Geo& gf, float* zp, const Geo& gp, int pool_h, double* out_stats,
                    unsigned char* amax, cudaStream_t s, const BnCoef* post = nullptr, void* a_mma = nullptr,
                    int fmt = FMT_F32);
// dzf (fmt planes, full-res geometry) <- dzp routed to the first maximum of every window
// absmax (optional): GradScale of the half-precision destination (common.cuh)
int maxpool_backward(const float* dzp, const Geo& gp, const float* zf, const Geo& gf, int pool_h,
                     void* dzf, int fmt, const unsigned* absmax, cudaStream_t s);
// the same routing from the arg-max bytes of maxpool_forward (dzp: float32 or half + GradScale)
int maxpool_backward_amax(GradRef dzp, const Geo& gp, const unsigned char* amax, const Geo& gf, int pool_h, void* dzf,
                          int fmt, const unsigned* absmax, cudaStream_t s);

// global max over (H, W) per (n, c): feat[n*feat_stride + feat_off + c], argrow[n*C + c] (padded row)
// scratch: gmax_scratch_bytes(g) bytes (packed per-(n, c) atomic-max slots)
size_t gmax_scratch_bytes(const Geo& g);
int gmax_forward(const float* x, const Geo& g, float* feat, int feat_stride, int feat_off, int* argrow, void* scratch,
                 cudaStream_t s);
// dx[argrow, c] += dfeat[n*feat_stride + feat_off + c]
int gmax_backward(const float* dfeat, int feat_stride, int feat_off, const int* argrow, const Geo& g,
                  float* dx, cudaStream_t s);
// the same into a scaled half plane (GradScale of `bits`)
int gmax_backward_h16(const float* dfeat, int feat_stride, int feat_off, const int* argrow, const Geo& g, void* dx,
                      const unsigned* bits, cudaStream_t s);
// bits[0] = float32 bit pattern of max |x[i]|, i < n (one CTA)
int absmax_bits(const float* x, long long n, unsigned* bits, cudaStream_t s);

// backward of a = act(BN(z) [+ residual]) given dA = dA1 (+ dA2; dA2.p == nullptr: none):
//   reduce  : partials [nblk][5][Cs] doubles = sum dy, sum dy*zhat, sum dslope, max |dy|, max |zhat|
//   finalize: dgamma, dbeta, dslope (C entries, written) and c1 = mean dy, c2 = mean dy*zhat (Cs);
//             absmax (optional, zeroed by the caller): atomicMax of the float32 bits of a bound on |dz| (GradScale);
//             absmax_dy (optional): the same for |dy| (scale of a half-precision dres);
//             extra_bits (optional): bit pattern of a bound that is ADDED to the |dz| bound (a later scatter-add into dz)
//   apply   : dz = scale * (dy - c1 - zhat*c2) -> fmt planes (half formats: times the GradScale of `absmax`);
//             dres (optional) = dy, float32 or (dres_bits != nullptr) one half plane with that GradScale
// Gradients come as float32 planes or scaled half planes (GradRef, common.cuh).  When every incoming gradient is a half
// plane, there is no residual / dropout and dz is float32 or ONE half plane, the compact 8-channel kernels run
// (eltwise.cu): a_hi (optional) is then the hi half plane of the stored activation a = act(BN(z)), and channel groups
// whose inverse map a -> zhat is well conditioned read it instead of the float32 z (half the bytes).  The 4-channel
// kernels ignore a_hi.
int bn_act_bwd_reduce(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn,
                      Residual res, Dropout dr, double* partials, cudaStream_t s);
// number of partial records bn_act_bwd_reduce writes for these operands (the `nblk` of bn_bwd_finalize)
int bn_bwd_num_blocks(GradRef dA1, GradRef dA2, const Geo& g, Residual res, Dropout dr);
int bn_bwd_finalize(const double* partials, int nblk, long long count, int C, int Cs, const float* bn_scale,
                    float* dgamma, float* dbeta, float* dslope, float* c1, float* c2, unsigned* absmax,
                    unsigned* absmax_dy, const unsigned* extra_bits, cudaStream_t s);
int bn_act_bwd_apply(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn,
                     Residual res, Dropout dr, const float* c1, const float* c2, void* dz, int fmt, void* dres,
                     const unsigned* dres_bits, const unsigned* absmax, cudaStream_t s);

// Compact backward of a residual activation out = prelu(BN(z) + r), r = prelu(res.zr * res.scale + res.shift) (mixed
// mode): eight channels per thread.  Where the slope is well conditioned the reduce pass takes y and its sign from the
// float32 `out` (instead of recomputing r from res.zr) and leaves one sign byte per (row, 8 channels) in `smask`
// (rows * Cs / 8 bytes); the apply pass then reads only the gradient, z and that byte.  Other channel groups (or
// out / smask == nullptr) recompute y from z and res.zr.  Writes dz and dres = dy as ONE scaled half plane each.
// Records: bn_res_bwd_compact_blocks(g) x [5][Cs].
int bn_res_bwd_compact_reduce(GradRef dA, const float* z, const float* out, unsigned char* smask, const Geo& g, BnCoef bn,
                              Residual res, double* partials, cudaStream_t s);
int bn_res_bwd_compact_blocks(const Geo& g);
int bn_res_bwd_compact_apply(GradRef dA, const float* z, const float* out, const unsigned char* smask, const Geo& g,
                             BnCoef bn, Residual res, const float* c1, const float* c2, void* dz, const unsigned* absmax,
                             void* dres, const unsigned* dres_bits, cudaStream_t s);

// column sums of a dense (rows, C) matrix with row stride ld (final Linear bias gradient)
int colsum(const float* x, long long rows, int C, int ld, float* out, cudaStream_t s);
// dst[r*ldd + c] = src[r*lds + c]
int copy2d(const float* src, long long rows, int C, int lds, float* dst, int ldd, cudaStream_t s);

// layout converters for taps / unit tests
int nchw_to_pf(const float* x, const Geo& g, void* dst, int fmt, cudaStream_t s);
int pf_to_nchw(const float* src, const Geo& g, float* dst, cudaStream_t s);

}  // namespace fsb
