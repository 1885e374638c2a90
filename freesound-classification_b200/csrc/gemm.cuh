// Row-shifted GEMM interface: every convolution / linear layer of the network is
//     Z[r, :Cout] = bias + sum_t A[r + off_t, :Cin] * W_t          (forward)
//     dA[r, :Cin] = sum_t dZ[r - off_t, :Cout] * W_t^T             (dgrad)
//     dW_t        = sum_r A[r + off_t, :]^T dZ[r, :]               (wgrad)
// over padded-flat NHWC tensors (common.cuh).  Two back ends implement it:
//   precision 0 : float32 CUDA-core tiles            (gemm_simt.cu)
//   precision 1 : split-half (hi + lo) tcgen05 / TMEM / TMA, three products  (gemm_tc.cu)   -- fp32-grade accuracy
//   precision 2 : single-pass half tcgen05                                   (gemm_tc.cu)
// (the network-level "mixed" mode runs forward GEMMs at precision 1 and backward GEMMs at precision 2, net.cu)
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Conv1d / nn.Linear
// (networks/classifiers.py:526-531, :75-80, :544-549) and their autograd backward.
#pragma once
#include "common.cuh"

namespace fsb {

struct ConvGeom {
    long long rows;      // padded-flat rows of the output (== rows of the input)
    int ntaps;           // 1 (1x1 / linear), 3 (1x3) or 9 (3x3)
    int offs[9];         // input row offset per tap, torch tap order (dy major, dx minor)
    int Cin, CsIn, Cout, CsOut;
};

inline ConvGeom make_conv_geom(const Geo& g, int Cin, int Cout, int kh, int kw) {
    ConvGeom c;
    c.rows = g.rows;
    c.ntaps = kh * kw;
    for (int i = 0; i < 9; ++i) c.offs[i] = 0;
    for (int dy = 0; dy < kh; ++dy)
        for (int dx = 0; dx < kw; ++dx) c.offs[dy * kw + dx] = (dy - kh / 2) * g.Wp + (dx - kw / 2);
    c.Cin = Cin; c.CsIn = round_up(Cin, 16);
    c.Cout = Cout; c.CsOut = round_up(Cout, 16);
    return c;
}

inline int act_fmt(int precision) { return precision == 0 ? FMT_F32 : (precision == 1 ? FMT_H16X2 : FMT_H16); }

// bytes of the packed-weight record (forward pack + dgrad pack + padded bias)
size_t packed_weight_bytes(int precision, const ConvGeom& c);
// w: torch layout (Cout, Cin, taps) float32; bias: Cout or nullptr
int pack_weights(int precision, const float* w, const float* bias, const ConvGeom& c, void* packed, cudaStream_t s);

// Optional fused BatchNorm statistics of the forward output: per-CTA partial sum / sum of squares over
// interior pixels, partials[*nblk][2][CsOut] doubles (the tcgen05 epilogue reduces the tile while it is in
// registers; the float32 back end runs the stand-alone statistics pass).  g = geometry of Z.
struct FwdStats {
    double* partials;
    const Geo* g;
    int* nblk;
};
int conv_gemm_fwd(int precision, const void* A, const void* packed, float* Z, const ConvGeom& c, const FwdStats* st,
                  cudaStream_t s);
// Eval forward (tensor-core back ends) with the FOLLOWING BatchNorm + PReLU folded into the epilogue: with running
// statistics the BatchNorm is a fixed per-channel affine map, so a = prelu((conv + bias) * scale + shift) is written straight
// as the hi / lo operand planes of the next GEMM and the float32 pre-activation is never materialised.  scale / shift:
// CsOut entries (bn_finalize in eval mode); slope: C entries or nullptr.
struct FwdAct {
    const float* scale;
    const float* shift;
    const float* slope;
    int C;
    const unsigned char* mask;    // interior mask of the output geometry (border rows are written as zeros); nullptr = no border
};
int conv_gemm_fwd_act(int precision, const void* A, const void* packed, void* a_out, const ConvGeom& c, const FwdAct& act,
                      cudaStream_t s);
// dz_absmax (optional, tensor-core back ends): GradScale of dZ (common.cuh) -- dZ holds 2^k * gradient, the result is
// multiplied by 2^-k.
// out_half_mul (optional, tensor-core back ends; needs dz_absmax): dA is written as ONE half plane holding 2^j * dA,
// j = gs_exponent2(dz_absmax, out_half_mul); *out_half_mul = max column L1 norm of the weights (weight_l1_bounds), so
// that max|dZ| * (*out_half_mul) bounds |dA| (compact backward of the mixed mode).  nullptr: float32 dA.
int conv_gemm_dgrad(int precision, const void* dZ, const void* packed, void* dA, const ConvGeom& c,
                    const unsigned* dz_absmax, const float* out_half_mul, cudaStream_t s);

// out[l] = max over input channels ci of sum_{co, tap} |w_l[co][ci][tap]| for up to 32 layers in one launch
struct WeightL1Job { const float* w; int Cin, Cout, ntaps; };
int weight_l1_bounds(const WeightL1Job* jobs, int njobs, float* out, cudaStream_t s);

size_t wgrad_scratch_bytes(int precision, const ConvGeom& c);
// dw: torch layout (Cout, Cin, taps), fully overwritten
int conv_gemm_wgrad(int precision, const void* A, const void* dZ, float* dw, void* scratch, const ConvGeom& c,
                    const unsigned* dz_absmax, cudaStream_t s);

// --- per-backend entry points -------------------------------------------------------------------
size_t simt_packed_weight_bytes(const ConvGeom& c);
int simt_pack_weights(const float* w, const float* bias, const ConvGeom& c, void* packed, cudaStream_t s);
int simt_fwd(const float* A, const void* packed, float* Z, const ConvGeom& c, cudaStream_t s);
int simt_dgrad(const float* dZ, const void* packed, float* dA, const ConvGeom& c, cudaStream_t s);
// split-K variants for linear layers with few rows (the FC head); scratch >= simt_skinny_scratch_bytes(c)
size_t simt_skinny_scratch_bytes(const ConvGeom& c);
int simt_skinny_fwd(const float* A, const void* packed, float* Z, const ConvGeom& c, void* scratch, cudaStream_t s);
int simt_skinny_dgrad(const float* dZ, const void* packed, float* dA, const ConvGeom& c, void* scratch, cudaStream_t s);
size_t simt_wgrad_scratch_bytes(const ConvGeom& c);
int simt_wgrad(const float* A, const float* dZ, float* dw, void* scratch, const ConvGeom& c, cudaStream_t s);

size_t tc_packed_weight_bytes(const ConvGeom& c);
int tc_pack_weights(const float* w, const float* bias, const ConvGeom& c, void* packed, cudaStream_t s);
int tc_fwd(int precision, const void* A, const void* packed, float* Z, const ConvGeom& c, const FwdStats* st,
           cudaStream_t s);
int tc_fwd_act(int precision, const void* A, const void* packed, void* a_out, const ConvGeom& c, const FwdAct& act,
               cudaStream_t s);
int tc_max_ctas();
int tc_dgrad(int precision, const void* dZ, const void* packed, void* dA, const ConvGeom& c, const unsigned* dz_absmax,
             const float* out_half_mul, cudaStream_t s);
size_t tc_wgrad_scratch_bytes(const ConvGeom& c);
int tc_wgrad(int precision, const void* A, const void* dZ, float* dw, void* scratch, const ConvGeom& c,
             const unsigned* dz_absmax, cudaStream_t s);

}  // namespace fsb
