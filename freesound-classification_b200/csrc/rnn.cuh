// aggregation_type == "rnn" deep-supervision heads (networks/classifiers.py:514-522, 592-597): frequency mean ->
// LayerNorm -> bidirectional GRU(128) -> final hidden states.  See rnn.cu.
#pragma once
#include "common.cuh"
#include "eltwise.cuh"
#include "gemm.cuh"

namespace fsb {

struct RnnSave {            // gate values kept for the backward pass: [direction][r, z, candidate, W_hn h + b_hn, h_prev]
    float* p[10];
};

struct RnnHead {
    int N, H, W, C, Cs;                      // geometry of the block output this head reads
    ConvGeom g_ih, g_hh;                     // row GEMMs over rows = N * W: C -> 384 and 128 -> 384
    float *xhat, *xln, *rstd;                // LayerNorm: normalised input, affine output (rows, Cs), 1 / std (rows)
    float* gi[2];                            // W_ih x + b_ih per direction, (rows, 384)
    float* save[10];                         // see RnnSave (training only)
    float *dgi[2], *dgh[2], *dxln[2];        // backward: d(W_ih x + b_ih), d(W_hh h + b_hh) (rows, 384), d LN(x) (rows, Cs)
    float *dxl, *prod;                       // backward: summed d LN(x) and its product with xhat (rows, Cs)
};

size_t rnn_head_floats(int N, int W, int Cs, int training);
void rnn_head_carve(RnnHead& h, float* base, int N, int H, int W, int C, int Cs, int training);
size_t rnn_packed_bytes(int C);

// P / G: the head's 10 parameters / gradient destinations in named_parameters() order:
//   ln.weight, ln.bias, weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, then the four *_reverse tensors
int rnn_head_forward(RnnHead& h, const float* out_pf, const Geo& g, const float* const* P, void* const* pk_ih, float* feats,
                     int feat_stride, int feat_off, int training, cudaStream_t s);
int rnn_head_backward(RnnHead& h, const float* dfeats, int feat_stride, int feat_off, const float* const* P,
                      void* const* pk_ih, float* const* G, float* d_out_pf, const Geo& g, void* wgrad_scratch,
                      cudaStream_t s);

}  // namespace fsb
