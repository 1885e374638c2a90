#include "conv0.cuh"

namespace fsb {

static const int PS_BLOCKS = 592;

__global__ void __launch_bounds__(256) plain_stats_kernel(const float* __restrict__ x, long long n, double* partials16) {
    double s = 0.0, ss = 0.0;
    float fs = 0.f, fss = 0.f;
    int cnt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = x[i];
        fs += v; fss += v * v;
        if (++cnt == 32) { s += fs; ss += fss; fs = fss = 0.f; cnt = 0; }
    }
    s += fs; ss += fss;
    __shared__ double r0[256], r1[256];
    r0[threadIdx.x] = s; r1[threadIdx.x] = ss;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) { r0[threadIdx.x] += r0[threadIdx.x + st]; r1[threadIdx.x] += r1[threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x < 16) {
        // channel 0 carries the sums; the remaining 15 channels of the record are zeroed
        partials16[((long long)blockIdx.x * 2 + 0) * 16 + threadIdx.x] = threadIdx.x == 0 ? r0[0] : 0.0;
        partials16[((long long)blockIdx.x * 2 + 1) * 16 + threadIdx.x] = threadIdx.x == 0 ? r1[0] : 0.0;
    }
}

int plain_stats_blocks() { return PS_BLOCKS; }

int plain_stats(const float* x, long long n, double* partials16, cudaStream_t s) {
    plain_stats_kernel<<<PS_BLOCKS, 256, 0, s>>>(x, n, partials16);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
static constexpr int C0_PX = 32;                 // pooled pixels per tile (one pooled row segment)
static constexpr int C0_PW = 2 * C0_PX + 2;      // input patch width
static constexpr int C0_THREADS = 128;

__device__ __forceinline__ float freq_enc(int h, int H) {
    float step = 2.0f / (float)(H - 1);
    return h < H / 2 ? -1.0f + step * (float)h : 1.0f - step * (float)(H - 1 - h);
}

// loads the BN-applied 2-channel input patch (rows 2*py-1 .. 2*py+2, cols 2*px0-1 .. 2*px0+2*PX) into
// shared memory; out-of-image positions are zero (conv padding) and flagged in `valid`.
__device__ __forceinline__ void load_patch(const float* __restrict__ feat, int n, int H, int W, int py, int px0,
                                           float sc0, float sh0, float sc1, float sh1, float mean0, float istd0,
                                           float mean1, float istd1, float (*u)[4][C0_PW], float (*xh)[4][C0_PW],
                                           float (*valid)[C0_PW]) {
    for (int e = threadIdx.x; e < 4 * C0_PW; e += blockDim.x) {
        int r = e / C0_PW, c = e - r * C0_PW;
        int y = 2 * py - 1 + r, x = 2 * px0 - 1 + c;
        bool ok = y >= 0 && y < H && x >= 0 && x < W;
        float v0 = 0.f, v1 = 0.f, h0 = 0.f, h1 = 0.f;
        if (ok) {
            float f = __ldg(feat + ((long long)n * H + y) * W + x);
            float e1 = freq_enc(y, H);
            v0 = fmaf(f, sc0, sh0);
            v1 = fmaf(e1, sc1, sh1);
            h0 = (f - mean0) * istd0;
            h1 = (e1 - mean1) * istd1;
        }
        u[0][r][c] = v0; u[1][r][c] = v1;
        if (xh) { xh[0][r][c] = h0; xh[1][r][c] = h1; }
        if (valid) valid[r][c] = ok ? 1.f : 0.f;
    }
}

__global__ void __launch_bounds__(C0_THREADS)
conv0_fwd_kernel(const float* __restrict__ feat, int N, int H, int W, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ w, const float* __restrict__ b,
                 float* __restrict__ zp, unsigned char* __restrict__ amax, Geo gp) {
    __shared__ float u[2][4][C0_PW];
    const int px0 = blockIdx.x * C0_PX, py = blockIdx.y, n = blockIdx.z;
    load_patch(feat, n, H, W, py, px0, scale[0], shift[0], scale[1], shift[1], 0.f, 0.f, 0.f, 0.f, u, nullptr, nullptr);
    __syncthreads();
    const int npx = min(C0_PX, gp.W - px0);
    for (int c = threadIdx.x; c < gp.Cs; c += blockDim.x) {
        float wr[2][3][3];
        float bias = 0.f;
        const bool real = c < gp.C;
        if (real) {
#pragma unroll
            for (int i = 0; i < 18; ++i) (&wr[0][0][0])[i] = w[c * 18 + i];
            bias = b[c];
        }
        for (int p = 0; p < npx; ++p) {
            float best = 0.f;
            int bpos = 0;
            if (real) {
                best = -INFINITY;
#pragma unroll
                for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                    for (int sx = 0; sx < 2; ++sx) {
                        float acc = bias;
#pragma unroll
                        for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx)
                                    acc = fmaf(wr[ci][dy][dx], u[ci][sy + dy][2 * p + sx + dx], acc);
                        if (acc > best) { best = acc; bpos = sy * 2 + sx; }      // first maximum in window scan order
                    }
            }
            const long long o = geo_row(gp, n, py, px0 + p) * gp.Cs + c;
            zp[o] = best;
            if (amax) amax[o] = (unsigned char)bpos;      // training: the backward pass routes the gradient by it
        }
    }
}

int conv0_forward(const float* feat, int N, int H, int W, const float* scale, const float* shift, const float* w,
                  const float* b, float* zp, unsigned char* amax, const Geo& gp, cudaStream_t s) {
    FSB_REQUIRE(gp.H == H / 2 && gp.W == W / 2 && gp.N == N, "conv0: geometry mismatch");
    FSB_REQUIRE(gp.H <= 65535 && N <= 65535, "conv0: grid too large");
    dim3 grid((gp.W + C0_PX - 1) / C0_PX, gp.H, N);
    conv0_fwd_kernel<<<grid, C0_THREADS, 0, s>>>(feat, N, H, W, scale, shift, w, b, zp, amax, gp);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
static const int C0B_BLOCKS = 1184;
static const int C0B_REC = 22;    // 18 dW + 2 dgamma-terms + 2 dbeta-terms per output channel

int conv0_bwd_blocks() { return C0B_BLOCKS; }
size_t conv0_bwd_scratch_bytes(const Geo& gp) { return (size_t)C0B_BLOCKS * C0B_REC * gp.Cs * sizeof(float); }

__global__ void __launch_bounds__(C0_THREADS)
conv0_bwd_kernel(const float* __restrict__ feat, int N, int H, int W, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                 const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ dzp,
                 const unsigned char* __restrict__ amax, Geo gp, float* __restrict__ partials) {
    __shared__ __align__(16) float u[2][4][C0_PW];
    __shared__ float xh[2][4][C0_PW];
    __shared__ float valid[4][C0_PW];
    const int tiles_x = (gp.W + C0_PX - 1) / C0_PX;
    const long long ntiles = (long long)tiles_x * gp.H * N;
    // each thread owns channels c = threadIdx.x + k*blockDim.x ; accumulators live across tiles, so
    // the channel loop is the OUTER loop and tiles are re-visited per channel chunk
    for (int c = threadIdx.x; c < gp.Cs + (int)blockDim.x; c += blockDim.x) {
        const bool real = c < gp.C;   // uniform participation in __syncthreads below
        if (c - (int)threadIdx.x >= gp.Cs) break;
        float wr[2][3][3];
        float bias = 0.f;
        float acc[C0B_REC];
#pragma unroll
        for (int i = 0; i < C0B_REC; ++i) acc[i] = 0.f;
        if (real) {
#pragma unroll
            for (int i = 0; i < 18; ++i) (&wr[0][0][0])[i] = w[c * 18 + i];
            bias = b[c];
        }
        // Fast path (BN_in scale != 0): the normalised input is an affine function of the BN output,
        // xhat = (u - shift) / gamma on valid positions, so sum_t w_t * xhat_t follows from the per-input-channel conv
        // partial sums S_ci that the arg-max recomputation produces anyway; the 4x4 patch lives in registers and the
        // weight gradient is accumulated with a one-hot gradient per window position (static indexing: no dynamic
        // shared-memory gathers, 144 FMAs per (pixel, channel) instead of ~130 FMAs + ~90 shared loads).
        const float sc0 = scale[0], sc1 = scale[1], sh0 = shift[0], sh1 = shift[1];
        const bool fast = sc0 != 0.f && sc1 != 0.f;
        float accV[9], gsum = 0.f;          // sum_q g valid_t over border windows, sum_q g over interior windows
#pragma unroll
        for (int i = 0; i < 9; ++i) accV[i] = 0.f;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int tx = (int)(tile % tiles_x);
            long long t2 = tile / tiles_x;
            int py = (int)(t2 % gp.H);
            int n = (int)(t2 / gp.H);
            int px0 = tx * C0_PX;
            __syncthreads();
            load_patch(feat, n, H, W, py, px0, sc0, sh0, sc1, sh1, mean[0], invstd[0], mean[1], invstd[1], u,
                       fast ? nullptr : xh, valid);
            __syncthreads();
            if (!real) continue;
            const int npx = min(C0_PX, gp.W - px0);
            if (fast) {
                // The forward pass stored the arg-max position of every pool window, so the gradient is routed with
                // 18 FMAs per (pixel, channel) (the previous version recomputed the four conv outputs: 72 FMAs, then
                // accumulated a one-hot gradient over all four positions: 144 more).  The BatchNorm-input terms are
                // linear in the accumulated weight gradient and need no per-pixel work:
                //   sum_q g S_ci = sum_t w[ci][t] acc[ci][t],   sum_q g V_ci = sum_t w[ci][t] (sum_q g valid_t)
                // where valid_t = 1 for every tap of a window that does not touch the image border.
                const bool row_border = 2 * py - 1 < 0 || 2 * py + 2 >= H;
                const long long o0 = geo_row(gp, n, py, px0) * gp.Cs + c;
                float g_next = dzp[o0];
                int pos_next = amax[o0];
                for (int p = 0; p < npx; ++p) {
                    const float g = g_next;
                    const int pos = pos_next;
                    if (p + 1 < npx) {                   // prefetch: the load latency hides behind this pixel's FMAs
                        g_next = dzp[o0 + (long long)(p + 1) * gp.Cs];
                        pos_next = amax[o0 + (long long)(p + 1) * gp.Cs];
                    }
                    const int sy = pos >> 1, sx = pos & 1;
                    const float* u0 = &u[0][sy][2 * p + sx];
                    const float* u1 = &u[1][sy][2 * p + sx];
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            acc[dy * 3 + dx] = fmaf(g, u0[dy * C0_PW + dx], acc[dy * 3 + dx]);
                            acc[9 + dy * 3 + dx] = fmaf(g, u1[dy * C0_PW + dx], acc[9 + dy * 3 + dx]);
                        }
                    const int x0 = 2 * (px0 + p) - 1;
                    if (row_border || x0 < 0 || x0 + 3 >= W) {      // window touches the image border: per-tap validity
                        const float* vv = &valid[sy][2 * p + sx];
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) accV[dy * 3 + dx] = fmaf(g, vv[dy * C0_PW + dx], accV[dy * 3 + dx]);
                    } else {
                        gsum += g;
                    }
                }
                continue;
            }
            for (int p = 0; p < npx; ++p) {
                float g = dzp[geo_row(gp, n, py, px0 + p) * gp.Cs + c];
                // recompute the four conv outputs to find the (first) arg-max of the pool window
                float best = -INFINITY;
                int bsy = 0, bsx = 0;
#pragma unroll
                for (int sy = 0; sy < 2; ++sy)
#pragma unroll
                    for (int sx = 0; sx < 2; ++sx) {
                        float a = bias;
#pragma unroll
                        for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx)
                                    a = fmaf(wr[ci][dy][dx], u[ci][sy + dy][2 * p + sx + dx], a);
                        if (a > best) { best = a; bsy = sy; bsx = sx; }
                    }
#pragma unroll
                for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            int r = bsy + dy, cc = 2 * p + bsx + dx;
                            acc[ci * 9 + dy * 3 + dx] = fmaf(g, u[ci][r][cc], acc[ci * 9 + dy * 3 + dx]);
                            float gw = g * wr[ci][dy][dx];
                            acc[18 + ci] = fmaf(gw, xh[ci][r][cc], acc[18 + ci]);
                            acc[20 + ci] = fmaf(gw, valid[r][cc], acc[20 + ci]);
                        }
            }
        }
        if (fast && real) {
            float sgs0 = 0.f, sgs1 = 0.f, accB0 = 0.f, accB1 = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float w0 = wr[0][t / 3][t % 3], w1 = wr[1][t / 3][t % 3], vg = gsum + accV[t];
                sgs0 = fmaf(w0, acc[t], sgs0);
                sgs1 = fmaf(w1, acc[9 + t], sgs1);
                accB0 = fmaf(w0, vg, accB0);
                accB1 = fmaf(w1, vg, accB1);
            }
            const float accG0 = sgs0 - sh0 * accB0, accG1 = sgs1 - sh1 * accB1;
            // xhat = (u - shift) * invstd / scale - mean * invstd   on valid positions
            acc[18] = accG0 * (invstd[0] / sc0) - mean[0] * invstd[0] * accB0;
            acc[19] = accG1 * (invstd[1] / sc1) - mean[1] * invstd[1] * accB1;
            acc[20] = accB0;
            acc[21] = accB1;
        }
        if (c < gp.Cs) {
            float* o = partials + (long long)blockIdx.x * C0B_REC * gp.Cs;
#pragma unroll
            for (int i = 0; i < C0B_REC; ++i) o[i * gp.Cs + c] = acc[i];
        }
    }
}

// One CTA per record k (18 dW taps, 2 dgamma terms, 2 dbeta terms): 32 channels x 32 slices of the per-block
// partials are summed at a time in a fixed order (deterministic); records 18..21 are then summed over channels.
__global__ void __launch_bounds__(1024)
conv0_bwd_finalize_kernel(const float* __restrict__ partials, int nblk, Geo gp, float* dw, float* db, float* dgamma_in,
                          float* dbeta_in) {
    __shared__ double red[32][33];
    __shared__ double tot[512];
    const int k = blockIdx.x;
    const int slice = threadIdx.x >> 5, cl = threadIdx.x & 31;
    for (int c0 = 0; c0 < gp.Cs; c0 += 32) {
        const int c = c0 + cl;
        double acc = 0.0;
        if (c < gp.C) {
#pragma unroll 8
            for (int b = slice; b < nblk; b += 32) acc += (double)partials[((long long)b * C0B_REC + k) * gp.Cs + c];
        }
        red[slice][cl] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) t += red[j][cl];
            if (c < 512) tot[c] = t;
            if (k < 18 && c < gp.C) dw[c * 18 + k] = (float)t;
            if (k == 0 && c < gp.C) db[c] = 0.f;   // conv bias feeds a batch-statistics BN: analytically zero gradient
        }
        __syncthreads();
    }
    if (k >= 18 && threadIdx.x == 0) {
        double t = 0.0;
        for (int c = 0; c < gp.C; ++c) t += tot[c];
        if (k < 20) dgamma_in[k - 18] = (float)t;
        else dbeta_in[k - 20] = (float)t;
    }
}

int conv0_bwd_finalize(const float* partials, int nblk, const Geo& gp, float* dw, float* db, float* dgamma_in,
                       float* dbeta_in, cudaStream_t s) {
    FSB_REQUIRE(gp.Cs <= 512, "conv0: at most 512 output channels");
    conv0_bwd_finalize_kernel<<<C0B_REC, 1024, 0, s>>>(partials, nblk, gp, dw, db, dgamma_in, dbeta_in);
    FSB_LAUNCHED();
    return 0;
}

int conv0_backward(const float* feat, int N, int H, int W, const float* scale, const float* shift, const float* mean,
                   const float* invstd, const float* w, const float* b, const float* dzp, const unsigned char* amax,
                   const Geo& gp, float* dw, float* db, float* dgamma_in, float* dbeta_in, void* scratch, cudaStream_t s) {
    FSB_REQUIRE(amax != nullptr, "conv0_backward: needs the arg-max map of the forward pass");
    conv0_bwd_kernel<<<C0B_BLOCKS, C0_THREADS, 0, s>>>(feat, N, H, W, scale, shift, mean, invstd, w, b, dzp, amax, gp,
                                                       (float*)scratch);
    FSB_LAUNCHED();
    FSB_REQUIRE(gp.Cs <= 512, "conv0: at most 512 output channels");
    conv0_bwd_finalize_kernel<<<C0B_REC, 1024, 0, s>>>((const float*)scratch, C0B_BLOCKS, gp, dw, db, dgamma_in, dbeta_in);
    FSB_LAUNCHED();
    return 0;
}

}  // namespace fsb
