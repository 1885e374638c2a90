// aggregation_type == "rnn" deep-supervision heads (networks/classifiers.py:514-522, 592-597):
//     block output (N, C, H, W) -> mean over the frequency axis H -> (N, W, C) -> LayerNorm(C)
//     -> bidirectional GRU(C -> 128, batch_first) over the W time steps -> final hidden states [forward | backward]
// forward and backward, float32 throughout (the heads hold < 1 % of the FLOPs; the recurrence is latency bound).
//
//   freqmean_ln_kernel   one warp per (n, w): mean over H of the padded-flat block output, LayerNorm over channels
//   input projections    gi = LN(x) W_ih^T + b_ih for all time steps at once: the CUDA-core row GEMM (gemm_simt.cu)
//   gru_fwd_kernel       one CTA per (4 samples, direction): thread j keeps row j of W_hh (3 x 128 rows of 128) in
//                        registers, the hidden state lives in shared memory; gate values are kept for the backward pass
//   gru_bwd_kernel       reverse-time recurrence: thread (gate, k) keeps column k of that gate's W_hh block in registers
//                        for W_hh^T dgh; per-step gate gradients go to memory, and the weight / bias gradients are GEMMs
//                        and column sums over all (sample, step) rows afterwards (deterministic)
//   ln_bwd_scatter_kernel LayerNorm backward + the gradient of the frequency mean added into the block gradient
#include "rnn.cuh"

namespace fsb {
namespace {

constexpr int RH = 128;            // GRU hidden size (rnn_size, networks/classifiers.py:509)
constexpr int RG = 3 * RH;         // gate rows r | z | n
constexpr int RSB = 4;             // samples per recurrence CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------
// x[n, w, c] = mean_h out[n, h, w, c];  xhat = (x - mu) * rstd over c < C;  xln = xhat * gamma + beta (pad channels 0)
__global__ void __launch_bounds__(256)
freqmean_ln_kernel(const float* __restrict__ out, Geo g, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ xhat, float* __restrict__ xln, float* __restrict__ rstd) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int rows = g.N * g.W;
    if (warp >= rows) return;
    const int n = warp / g.W, w = warp - n * g.W;
    constexpr int MAXV = 16;                           // Cs <= 512
    float v[MAXV];
    const float inv_h = 1.0f / (float)g.H;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        float s = 0.f;
        if (c < g.C)
            for (int h = 0; h < g.H; ++h) s += out[geo_row(g, n, h, w) * g.Cs + c];
        v[i] = s * inv_h;
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) sum += (lane + 32 * i < g.C) ? v[i] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mu = sum / (float)g.C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const float d = (lane + 32 * i < g.C) ? v[i] - mu : 0.f;
        sq = fmaf(d, d, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rs = 1.0f / sqrtf(sq / (float)g.C + 1e-5f);     // nn.LayerNorm: biased variance, eps 1e-5
    if (lane == 0) rstd[warp] = rs;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < g.Cs) {
            const float xh = c < g.C ? (v[i] - mu) * rs : 0.f;
            xhat[(long long)warp * g.Cs + c] = xh;
            xln[(long long)warp * g.Cs + c] = c < g.C ? fmaf(xh, gamma[c], beta[c]) : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward recurrence.  gi (rows, 384) holds W_ih x + b_ih; rows are (n * W + t).  dir 1 walks t = W-1 .. 0.
__global__ void __launch_bounds__(RG)
gru_fwd_kernel(const float* __restrict__ gi0, const float* __restrict__ gi1, const float* __restrict__ whh0,
               const float* __restrict__ whh1, const float* __restrict__ bhh0, const float* __restrict__ bhh1, int N, int W,
               float* __restrict__ feats, int feat_stride, int feat_off, RnnSave save) {
    const int dir = blockIdx.y, n0 = blockIdx.x * RSB;
    const float* gi = dir ? gi1 : gi0;
    const float* whh = dir ? whh1 : whh0;
    const float* bhh = dir ? bhh1 : bhh0;
    float* s_r = save.p[dir * 5 + 0];        // all null in eval mode
    float* s_z = save.p[dir * 5 + 1];
    float* s_c = save.p[dir * 5 + 2];
    float* s_hn = save.p[dir * 5 + 3];
    float* s_hp = save.p[dir * 5 + 4];
    __shared__ float h_s[RSB][RH];
    __shared__ float gh_s[RSB][RG];
    const int j = threadIdx.x;
    float wrow[RH];
#pragma unroll
    for (int k = 0; k < RH; ++k) wrow[k] = whh[j * RH + k];
    const float bj = bhh[j];
    for (int i = j; i < RSB * RH; i += RG) (&h_s[0][0])[i] = 0.f;
    __syncthreads();
    for (int step = 0; step < W; ++step) {
        const int t = dir ? W - 1 - step : step;
        float acc[RSB];
#pragma unroll
        for (int s = 0; s < RSB; ++s) acc[s] = bj;
#pragma unroll
        for (int k = 0; k < RH; ++k) {
#pragma unroll
            for (int s = 0; s < RSB; ++s) acc[s] = fmaf(wrow[k], h_s[s][k], acc[s]);
        }
#pragma unroll
        for (int s = 0; s < RSB; ++s) gh_s[s][j] = acc[s];
        __syncthreads();
        for (int item = j; item < RSB * RH; item += RG) {
            const int s = item / RH, k = item - s * RH;
            const int n = n0 + s;
            if (n < N) {
                const long long row = (long long)n * W + t;
                const float* g3 = gi + row * RG;
                const float r = sigmoidf_(g3[k] + gh_s[s][k]);
                const float z = sigmoidf_(g3[RH + k] + gh_s[s][RH + k]);
                const float hn = gh_s[s][2 * RH + k];
                const float cand = tanhf(fmaf(r, hn, g3[2 * RH + k]));
                const float hp = h_s[s][k];
                if (s_r) {
                    s_r[row * RH + k] = r; s_z[row * RH + k] = z; s_c[row * RH + k] = cand;
                    s_hn[row * RH + k] = hn; s_hp[row * RH + k] = hp;
                }
                h_s[s][k] = fmaf(z, hp - cand, cand);           // (1 - z) * cand + z * h
            }
        }
        __syncthreads();
    }
    for (int item = j; item < RSB * RH; item += RG) {
        const int s = item / RH, k = item - s * RH;
        if (n0 + s < N) feats[(long long)(n0 + s) * feat_stride + feat_off + dir * RH + k] = h_s[s][k];
    }
}

// backward recurrence: writes dgi (rows, 384) = d(W_ih x + b_ih) and dgh (rows, 384) = d(W_hh h + b_hh)
__global__ void __launch_bounds__(RG)
gru_bwd_kernel(const float* __restrict__ dfeats, int feat_stride, int feat_off, const float* __restrict__ whh0,
               const float* __restrict__ whh1, int N, int W, RnnSave save, float* __restrict__ dgi0,
               float* __restrict__ dgi1, float* __restrict__ dgh0, float* __restrict__ dgh1) {
    const int dir = blockIdx.y, n0 = blockIdx.x * RSB;
    const float* whh = dir ? whh1 : whh0;
    const float* s_r = save.p[dir * 5 + 0];
    const float* s_z = save.p[dir * 5 + 1];
    const float* s_c = save.p[dir * 5 + 2];
    const float* s_hn = save.p[dir * 5 + 3];
    const float* s_hp = save.p[dir * 5 + 4];
    float* dgi = dir ? dgi1 : dgi0;
    float* dgh = dir ? dgh1 : dgh0;
    __shared__ float dh_s[RSB][RH];
    __shared__ float dgh_s[RSB][RG];
    __shared__ float part_s[3][RSB][RH];
    const int j = threadIdx.x, gate = j / RH, kk = j - gate * RH;
    float wcol[RH];                                     // column kk of gate block `gate`: W_hh[gate * 128 + i][kk]
#pragma unroll
    for (int i = 0; i < RH; ++i) wcol[i] = whh[(gate * RH + i) * RH + kk];
    for (int item = j; item < RSB * RH; item += RG) {
        const int s = item / RH, k = item - s * RH;
        dh_s[s][k] = n0 + s < N ? dfeats[(long long)(n0 + s) * feat_stride + feat_off + dir * RH + k] : 0.f;
    }
    __syncthreads();
    for (int step = W - 1; step >= 0; --step) {
        const int t = dir ? W - 1 - step : step;
        for (int item = j; item < RSB * RH; item += RG) {
            const int s = item / RH, k = item - s * RH;
            const int n = n0 + s;
            float g_r = 0.f, g_z = 0.f, g_n = 0.f, g_hn = 0.f, dprev = 0.f;
            if (n < N) {
                const long long row = (long long)n * W + t;
                const float dh = dh_s[s][k];
                const float r = s_r[row * RH + k], z = s_z[row * RH + k], cand = s_c[row * RH + k];
                const float hn = s_hn[row * RH + k], hp = s_hp[row * RH + k];
                const float dcand = dh * (1.0f - z);
                const float dz = dh * (hp - cand);
                dprev = dh * z;
                g_n = dcand * (1.0f - cand * cand);
                g_hn = g_n * r;
                g_r = g_n * hn * r * (1.0f - r);
                g_z = dz * z * (1.0f - z);
                float* o = dgi + row * RG;
                o[k] = g_r; o[RH + k] = g_z; o[2 * RH + k] = g_n;
                float* oh = dgh + row * RG;
                oh[k] = g_r; oh[RH + k] = g_z; oh[2 * RH + k] = g_hn;
            }
            dgh_s[s][k] = g_r; dgh_s[s][RH + k] = g_z; dgh_s[s][2 * RH + k] = g_hn;
            dh_s[s][k] = dprev;
        }
        __syncthreads();
        float acc[RSB];
#pragma unroll
        for (int s = 0; s < RSB; ++s) acc[s] = 0.f;
#pragma unroll
        for (int i = 0; i < RH; ++i) {
#pragma unroll
            for (int s = 0; s < RSB; ++s) acc[s] = fmaf(wcol[i], dgh_s[s][gate * RH + i], acc[s]);
        }
#pragma unroll
        for (int s = 0; s < RSB; ++s) part_s[gate][s][kk] = acc[s];
        __syncthreads();
        for (int item = j; item < RSB * RH; item += RG) {
            const int s = item / RH, k = item - s * RH;
            dh_s[s][k] += (part_s[0][s][k] + part_s[1][s][k]) + part_s[2][s][k];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward for one (n, w) row per warp; dxl = dxln0 + dxln1.  Writes the per-row products needed for
// d(gamma) / d(beta) (column-summed afterwards) and adds dx / H into the block gradient at every frequency row.
__global__ void __launch_bounds__(256)
ln_bwd_scatter_kernel(const float* __restrict__ dx0, const float* __restrict__ dx1, const float* __restrict__ xhat,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, Geo g, float* __restrict__ dxl_out,
                      float* __restrict__ prod_out, float* __restrict__ d_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int rows = g.N * g.W;
    if (warp >= rows) return;
    const int n = warp / g.W, w = warp - n * g.W;
    constexpr int MAXV = 16;
    float gy[MAXV], xh[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        gy[i] = 0.f; xh[i] = 0.f;
        if (c < g.Cs) {
            const long long idx = (long long)warp * g.Cs + c;
            const float d = c < g.C ? dx0[idx] + dx1[idx] : 0.f;
            xh[i] = xhat[idx];
            dxl_out[idx] = d;
            prod_out[idx] = d * xh[i];
            gy[i] = c < g.C ? d * gamma[c] : 0.f;
            s1 += gy[i];
            s2 = fmaf(gy[i], xh[i], s2);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / (float)g.C, m2 = s2 / (float)g.C, rs = rstd[warp];
    const float inv_h = 1.0f / (float)g.H;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = lane + 32 * i;
        if (c < g.C) {
            const float dxm = rs * (gy[i] - m1 - xh[i] * m2) * inv_h;
            for (int h = 0; h < g.H; ++h) d_out[geo_row(g, n, h, w) * g.Cs + c] += dxm;
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
size_t rnn_head_floats(int N, int W, int Cs, int training) {
    const size_t rows = (size_t)N * W;
    size_t f = rows * Cs * 2 + rows + 2 * rows * RG;                      // xhat, xln, rstd, gi[2]
    if (training) f += 2 * 5 * rows * RH + 4 * rows * RG + 4 * rows * Cs; // saved gates, dgi/dgh[2], dxln[2], dxl, prod
    return f + 64;
}

void rnn_head_carve(RnnHead& h, float* base, int N, int H, int W, int C, int Cs, int training) {
    const size_t rows = (size_t)N * W;
    h.N = N; h.H = H; h.W = W; h.C = C; h.Cs = Cs;
    float* p = base;
    auto take = [&](size_t n) { float* r = p; p += n; return r; };
    h.xhat = take(rows * Cs); h.xln = take(rows * Cs); h.rstd = take(rows);
    for (int d = 0; d < 2; ++d) h.gi[d] = take(rows * RG);
    for (int i = 0; i < 10; ++i) h.save[i] = nullptr;
    for (int d = 0; d < 2; ++d) { h.dgi[d] = h.dgh[d] = h.dxln[d] = nullptr; }
    h.dxl = h.prod = nullptr;
    if (training) {
        for (int i = 0; i < 10; ++i) h.save[i] = take(rows * RH);
        for (int d = 0; d < 2; ++d) { h.dgi[d] = take(rows * RG); h.dgh[d] = take(rows * RG); h.dxln[d] = take(rows * Cs); }
        h.dxl = take(rows * Cs); h.prod = take(rows * Cs);
    }
    Geo gr = make_geo(1, 1, (int)rows, C, 0, 0);            // rows as a flat (1, 1, rows) geometry: no border
    h.g_ih = make_conv_geom(gr, C, RG, 1, 1);
    Geo gh = make_geo(1, 1, (int)rows, RH, 0, 0);
    h.g_hh = make_conv_geom(gh, RH, RG, 1, 1);
}

size_t rnn_packed_bytes(int C) {
    Geo gr = make_geo(1, 1, 16, C, 0, 0);
    return simt_packed_weight_bytes(make_conv_geom(gr, C, RG, 1, 1));
}

// P: ln.weight, ln.bias, then per direction weight_ih, weight_hh, bias_ih, bias_hh (torch named_parameters order)
int rnn_head_forward(RnnHead& h, const float* out_pf, const Geo& g, const float* const* P, void* const* pk_ih, float* feats,
                     int feat_stride, int feat_off, int training, cudaStream_t s) {
    FSB_REQUIRE(g.Cs <= 512, "rnn head: at most 512 channels");
    const int rows = h.N * h.W;
    freqmean_ln_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(out_pf, g, P[0], P[1], h.xhat, h.xln, h.rstd);
    FSB_LAUNCHED();
    for (int d = 0; d < 2; ++d) {
        FSB_TRY(simt_pack_weights(P[2 + 4 * d], P[4 + 4 * d], h.g_ih, pk_ih[d], s));
        FSB_TRY(simt_fwd(h.xln, pk_ih[d], h.gi[d], h.g_ih, s));
    }
    RnnSave save;
    for (int i = 0; i < 10; ++i) save.p[i] = training ? h.save[i] : nullptr;
    dim3 grid((h.N + RSB - 1) / RSB, 2);
    gru_fwd_kernel<<<grid, RG, 0, s>>>(h.gi[0], h.gi[1], P[3], P[7], P[5], P[9], h.N, h.W, feats, feat_stride, feat_off, save);
    FSB_LAUNCHED();
    return 0;
}

// G: gradient destinations in the same order as P.  d_out_pf receives += the gradient of the block output.
int rnn_head_backward(RnnHead& h, const float* dfeats, int feat_stride, int feat_off, const float* const* P,
                      void* const* pk_ih, float* const* G, float* d_out_pf, const Geo& g, void* wgrad_scratch,
                      cudaStream_t s) {
    const int rows = h.N * h.W;
    FSB_REQUIRE(h.save[0] != nullptr, "rnn head: backward needs a training-mode forward");
    RnnSave save;
    for (int i = 0; i < 10; ++i) save.p[i] = h.save[i];
    dim3 grid((h.N + RSB - 1) / RSB, 2);
    gru_bwd_kernel<<<grid, RG, 0, s>>>(dfeats, feat_stride, feat_off, P[3], P[7], h.N, h.W, save, h.dgi[0], h.dgi[1], h.dgh[0],
                                       h.dgh[1]);
    FSB_LAUNCHED();
    for (int d = 0; d < 2; ++d) {
        FSB_TRY(simt_wgrad(h.xln, h.dgi[d], G[2 + 4 * d], wgrad_scratch, h.g_ih, s));            // d W_ih
        FSB_TRY(simt_wgrad(h.save[d * 5 + 4], h.dgh[d], G[3 + 4 * d], wgrad_scratch, h.g_hh, s)); // d W_hh = dgh^T h_prev
        FSB_TRY(colsum(h.dgi[d], rows, RG, RG, G[4 + 4 * d], s));                                  // d b_ih
        FSB_TRY(colsum(h.dgh[d], rows, RG, RG, G[5 + 4 * d], s));                                  // d b_hh
        FSB_TRY(simt_dgrad(h.dgi[d], pk_ih[d], h.dxln[d], h.g_ih, s));                             // d LN(x)
    }
    ln_bwd_scatter_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(h.dxln[0], h.dxln[1], h.xhat, h.rstd, P[0], g, h.dxl, h.prod,
                                                                  d_out_pf);
    FSB_LAUNCHED();
    FSB_TRY(colsum(h.prod, rows, h.C, h.Cs, G[0], s));      // d gamma = sum dxl * xhat
    FSB_TRY(colsum(h.dxl, rows, h.C, h.Cs, G[1], s));       // d beta
    return 0;
}

}  // namespace fsb
