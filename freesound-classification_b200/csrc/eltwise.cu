#include <stdlib.h>

#include "eltwise.cuh"

namespace fsb {

// ---------------------------------------------------------------------------------------------
// launch shape shared by all kernels here: threadIdx.x <-> a float4 of channels (fixed for the
// thread, so per-channel reductions stay in registers), threadIdx.y <-> pixel; blockIdx.y covers
// channel chunks when Cs/4 exceeds blockDim.x; blockIdx.x strides over interior pixels.
// Consecutive x-threads touch consecutive 16 B and consecutive pixels are consecutive rows, so a
// warp's accesses are contiguous.
// ---------------------------------------------------------------------------------------------
struct EwShape {
    dim3 grid, block;
};

static int ew_max_blocks() {            // 8 x 148 SMs by default; FSB200_EW_BLOCKS overrides (tuning)
    static int v = 0;
    if (!v) {
        const char* e = getenv("FSB200_EW_BLOCKS");
        v = e ? atoi(e) : 1184;
        if (v < 1) v = 1184;
    }
    return v;
}
#define EW_MAX_BLOCKS ew_max_blocks()

static EwShape ew_shape(const Geo& g) {
    int cv = g.Cs / 4;
    int bx = cv < 64 ? cv : 64;
    int by = 256 / bx;
    if (by < 1) by = 1;
    long long need = (g.rows + by * 4 - 1) / (by * 4);
    int gx = (int)(need < 1 ? 1 : (need > EW_MAX_BLOCKS ? EW_MAX_BLOCKS : need));
    EwShape s;
    s.block = dim3(bx, by);
    s.grid = dim3(gx, (cv + bx - 1) / bx);
    return s;
}

int ew_num_blocks(const Geo& g) { return (int)ew_shape(g).grid.x; }

#define EW_CHECK(g) \
    FSB_REQUIRE(((g).padH == 0 && (g).padW == 0) || (g).mask != nullptr, "element-wise kernel: geometry has a border but no interior mask")

#define EW_PROLOGUE                                                        \
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;                  \
    const bool cok = cv < g.Cs / 4;                                        \
    const int c0 = cv * 4;

// Grid-stride loop over the PADDED rows (consecutive rows -> consecutive memory, no divisions); border rows
// are skipped through the geometry's byte mask (1 = interior pixel; nullptr = every row is interior).
#define EW_PIXEL_LOOP                                                                        \
    for (long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y; row < g.rows;     \
         row += (long long)gridDim.x * blockDim.y)                                            \
        if (g.mask == nullptr || g.mask[row])

// Same walk, two rows per iteration: the loads of both rows are issued before either is consumed (twice the
// bytes in flight per thread; the streaming kernels are latency-bound otherwise).  okA / okB: row is interior.
#define EW_PIXEL_LOOP2                                                                                  \
    const long long ew_stride = (long long)gridDim.x * blockDim.y;                                      \
    for (long long rowA = (long long)blockIdx.x * blockDim.y + threadIdx.y; rowA < g.rows;              \
         rowA += 2 * ew_stride)                                                                          \
        for (bool ew_once = true; ew_once;)                                                              \
            for (const long long rowB = rowA + ew_stride; ew_once;)                                      \
                for (const bool okA = g.mask == nullptr || g.mask[rowA],                                 \
                                okB = rowB < g.rows && (g.mask == nullptr || g.mask[rowB]);              \
                     ew_once; ew_once = false)

// interior mask of a padded-flat geometry
__global__ void interior_mask_kernel(Geo g, unsigned char* mask) {
    for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < g.rows;
         row += (long long)gridDim.x * blockDim.x) {
        int rr = (int)(row % ((long long)g.Hp * g.Wp));
        int yy = rr / g.Wp, xx = rr - yy * g.Wp;
        mask[row] = (yy >= g.padH && yy < g.padH + g.H && xx >= g.padW && xx < g.padW + g.W) ? 1 : 0;
    }
}

int pf_build_mask(const Geo& g, unsigned char* mask, cudaStream_t s) {
    int blocks = (int)((g.rows + 255) / 256 > 1184 ? 1184 : (g.rows + 255) / 256);
    interior_mask_kernel<<<blocks, 256, 0, s>>>(g, mask);
    FSB_LAUNCHED();
    return 0;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldh4(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ld_grad(const GradRef& r, long long idx) {
    return r.half ? ldh4(reinterpret_cast<const __half*>(r.p) + idx) : ld4(reinterpret_cast<const float*>(r.p) + idx);
}

// writes four channels in the destination format; half formats multiply by `scale` first (1 for activations, the
// tensor's power-of-two GradScale for gradients, common.cuh)
__device__ __forceinline__ void store_fmt(void* base, int fmt, long long plane_elems, long long idx, float4 v,
                                          float scale = 1.f) {
    if (fmt == FMT_F32) {
        st4(reinterpret_cast<float*>(base) + idx, v);
    } else {
        __half h[4], l[4];
        split_h16(v.x * scale, h[0], l[0]);
        split_h16(v.y * scale, h[1], l[1]);
        split_h16(v.z * scale, h[2], l[2]);
        split_h16(v.w * scale, h[3], l[3]);
        __half* hp = reinterpret_cast<__half*>(base) + idx;
        *reinterpret_cast<uint2*>(hp) = *reinterpret_cast<uint2*>(h);
        if (fmt == FMT_H16X2) *reinterpret_cast<uint2*>(hp + plane_elems) = *reinterpret_cast<uint2*>(l);
    }
}

// accumulate float4 streams: float32 inside runs of 32 pixels, double across runs
// Per-thread partial sums of float4 streams.  A thread visits rows / (gridDim.x * blockDim.y) ~ 10^2 pixels, so it sums
// in float32 (the previous version flushed to double every 32 values, which bought a factor ~3 in the worst-case
// rounding bound for 20 more registers per stream); the cross-thread and cross-block reductions run in double.
struct Acc4 {
    float4 f;
    double d[4];
    __device__ __forceinline__ void init() { f = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void add(float4 v) { f.x += v.x; f.y += v.y; f.z += v.z; f.w += v.w; }
    __device__ __forceinline__ void flush() { d[0] = f.x; d[1] = f.y; d[2] = f.z; d[3] = f.w; }
};

// reduce K Acc4 records across threadIdx.y and write partials[(blockIdx.x*K + k)*Cs + c]; the last NMAX records are
// combined with max instead of + (gradient magnitude bounds)
template <int K, int NMAX = 0>
__device__ __forceinline__ void block_reduce_store(Acc4 (&acc)[K], double* partials, int Cs, int c0, bool cok) {
    __shared__ double red[256 * 4];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc[k].flush();
        int t = threadIdx.y * blockDim.x + threadIdx.x;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) red[t * 4 + i] = acc[k].d[i];
        __syncthreads();
        if (threadIdx.y == 0 && cok) {
            double s[4] = {0, 0, 0, 0};
            for (int y = 0; y < (int)blockDim.y; ++y) {
                int tt = y * blockDim.x + threadIdx.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) s[i] = k >= K - NMAX ? fmax(s[i], red[tt * 4 + i]) : s[i] + red[tt * 4 + i];
            }
            double* o = partials + ((long long)blockIdx.x * K + k) * Cs + c0;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = s[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void zero_border_kernel(float4* buf, Geo g, long long plane_vec, int nplanes, int vec_per_row) {
    // one thread per (border row, vec); border rows enumerated per image
    int border_per_img = g.Hp * g.Wp - g.H * g.W;
    long long total = (long long)g.N * border_per_img * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int v = (int)(i % vec_per_row);
        long long t = i / vec_per_row;
        int b = (int)(t % border_per_img);
        int n = (int)(t / border_per_img);
        // map border index -> (y, x) in padded coords
        int y, x;
        int top = g.padH * g.Wp;
        if (b < top) {
            y = b / g.Wp; x = b % g.Wp;
        } else if (b < 2 * top) {
            int bb = b - top;
            y = g.Hp - g.padH + bb / g.Wp; x = bb % g.Wp;
        } else {
            int bb = b - 2 * top;             // side columns: H rows x 2*padW
            y = g.padH + bb / (2 * g.padW);
            int sx = bb % (2 * g.padW);
            x = sx < g.padW ? sx : g.Wp - 2 * g.padW + sx;
        }
        long long row = ((long long)n * g.Hp + y) * g.Wp + x;
        for (int p = 0; p < nplanes; ++p) buf[p * plane_vec + row * vec_per_row + v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

int pf_zero_border(void* buf, int fmt, const Geo& g, cudaStream_t s) {
    int border_per_img = g.Hp * g.Wp - g.H * g.W;
    if (border_per_img == 0) return 0;
    // a float4 is 4 float32 or 8 halves: vec_per_row counts 16-byte units per row per plane
    int vec_per_row = fmt == FMT_F32 ? g.Cs / 4 : g.Cs / 8;
    int nplanes = fmt == FMT_H16X2 ? 2 : 1;
    long long plane_vec = g.rows * vec_per_row;
    long long total = (long long)g.N * border_per_img * vec_per_row;
    int blocks = (int)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
    zero_border_kernel<<<blocks, 256, 0, s>>>((float4*)buf, g, plane_vec, nplanes, vec_per_row);
    FSB_LAUNCHED();
    return 0;
}

int pf_zero_all(void* buf, int fmt, const Geo& g, cudaStream_t s) {
    size_t bytes = (size_t)g.rows * g.Cs * 4;   // f32: 4 B/elem ; half: 2 planes x 2 B/elem
    (void)fmt;
    FSB_CUDA(cudaMemsetAsync(buf, 0, bytes, s));
    return 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_kernel(const float* __restrict__ xin, Geo g, double* partials) {
    EW_PROLOGUE
    Acc4 acc[2];
    acc[0].init(); acc[1].init();
    if (cok) {
        EW_PIXEL_LOOP {
            float4 v = ld4(xin + row * g.Cs + c0);
            acc[0].add(v);
            acc[1].add(make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w));
        }
    }
    block_reduce_store<2>(acc, partials, g.Cs, c0, cok);
}

int pf_stats(const float* x, const Geo& g, double* partials, cudaStream_t s) {
    EW_CHECK(g);
    EwShape sh = ew_shape(g);
    stats_kernel<<<sh.grid, sh.block, 0, s>>>(x, g, partials);
    FSB_LAUNCHED();
    return 0;
}

// Fixed-order parallel reduction of per-block partial sums: FIN_CH channels x FIN_SLICES slices per CTA (1024
// threads); every slice walks the blocks b = slice, slice + FIN_SLICES, ... (at most ~10 loads per thread for the 1184
// records of an element-wise pass), then two tree levels add the slice totals in a fixed order (deterministic).
constexpr int FIN_CH = 8;
constexpr int FIN_SLICES = 128;
constexpr int FIN_THREADS = FIN_CH * FIN_SLICES;

template <int K, int NMAX = 0>
__device__ __forceinline__ void reduce_partials(const double* __restrict__ partials, int nblk, int Cs, int c,
                                                double (&out)[K]) {
    __shared__ double red[K][FIN_SLICES][FIN_CH];
    __shared__ double red2[K][8][FIN_CH];
    const int slice = threadIdx.x / FIN_CH, cl = threadIdx.x % FIN_CH;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    if (c < Cs) {
#pragma unroll 4
        for (int b = slice; b < nblk; b += FIN_SLICES) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double v = partials[((long long)b * K + k) * Cs + c];
                acc[k] = k >= K - NMAX ? fmax(acc[k], v) : acc[k] + v;      // trailing records: max (magnitude bounds)
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) red[k][slice][cl] = acc[k];
    __syncthreads();
    if (threadIdx.x < 8 * FIN_CH) {            // 8 groups of 16 slices
        const int g = threadIdx.x / FIN_CH;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < FIN_SLICES / 8; ++j) {
                const double v = red[k][g * (FIN_SLICES / 8) + j][cl];
                t = k >= K - NMAX ? fmax(t, v) : t + v;
            }
            red2[k][g][cl] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = 0.0;
        if (threadIdx.x < FIN_CH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) t = k >= K - NMAX ? fmax(t, red2[k][j][cl]) : t + red2[k][j][cl];
        }
        out[k] = t;
    }
}

__global__ void __launch_bounds__(FIN_THREADS)
bn_finalize_kernel(const double* partials, int nblk, long long count, const float* gamma,
                   const float* beta, float* run_mean, float* run_var, long long* bn_count,
                   int training, int C, int Cs, float* scale, float* shift, float* mean,
                   float* invstd) {
    const int c = blockIdx.x * FIN_CH + (threadIdx.x % FIN_CH);
    double tot[2] = {0.0, 0.0};
    if (training) reduce_partials<2>(partials, nblk, Cs, c, tot);
    if (threadIdx.x >= FIN_CH || c >= Cs) return;
    if (c >= C) {
        scale[c] = 0.f; shift[c] = 0.f;
        if (mean) mean[c] = 0.f;
        if (invstd) invstd[c] = 0.f;
        return;
    }
    const float eps = 1e-5f, momentum = 0.1f;
    float m, istd;
    if (training) {
        double mu = tot[0] / (double)count;
        double var = tot[1] / (double)count - mu * mu;
        if (var < 0.0) var = 0.0;
        m = (float)mu;
        istd = (float)(1.0 / sqrt(var + (double)eps));
        if (run_mean) {
            double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
            run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * m;
            run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unbiased;
        }
        if (c == 0 && bn_count) *bn_count += 1;
    } else {
        m = run_mean[c];
        istd = 1.0f / sqrtf(run_var[c] + eps);
    }
    float sc = gamma[c] * istd;
    scale[c] = sc;
    shift[c] = beta[c] - m * sc;
    if (mean) mean[c] = m;
    if (invstd) invstd[c] = istd;
}

int bn_finalize(const double* partials, int nblk, long long count, const float* gamma, const float* beta,
                float* run_mean, float* run_var, long long* bn_count, int training, int C, int Cs,
                float* scale, float* shift, float* mean, float* invstd, cudaStream_t s) {
    bn_finalize_kernel<<<(Cs + FIN_CH - 1) / FIN_CH, FIN_THREADS, 0, s>>>(partials, nblk, count, gamma, beta, run_mean,
                                                     run_var, bn_count, training, C, Cs, scale, shift,
                                                     mean, invstd);
    FSB_LAUNCHED();
    return 0;
}

__global__ void freq_stats_kernel(int H, long long n_times_w, double* partials16) {
    // channel 1 of the block-0 input: v_h = linspace(-1, 1, H)[h] (float32, as torch.linspace)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0, ss = 0.0;
        float step = 2.0f / (float)(H - 1);
        for (int h = 0; h < H; ++h) {
            float v = h < H / 2 ? -1.0f + step * (float)h : 1.0f - step * (float)(H - 1 - h);
            s += v; ss += (double)v * v;
        }
        partials16[0 * 16 + 1] = s * (double)n_times_w;
        partials16[1 * 16 + 1] = ss * (double)n_times_w;
    }
}

int freq_encoding_stats(int H, long long n_times_w, double* partials16, cudaStream_t s) {
    freq_stats_kernel<<<1, 32, 0, s>>>(H, n_times_w, partials16);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float prelu1(float y, float a) { return y > 0.f ? y : a * y; }

__device__ __forceinline__ float4 affine4(float4 z, float4 sc, float4 sh) {
    return make_float4(fmaf(z.x, sc.x, sh.x), fmaf(z.y, sc.y, sh.y), fmaf(z.z, sc.z, sh.z), fmaf(z.w, sc.w, sh.w));
}
__device__ __forceinline__ float4 prelu4(float4 y, float4 a) {
    return make_float4(prelu1(y.x, a.x), prelu1(y.y, a.y), prelu1(y.z, a.z), prelu1(y.w, a.w));
}

__device__ __forceinline__ Dropout resolve_seed(Dropout dr) {
    if (dr.p > 0.f && dr.seed_ptr) dr.seed = *dr.seed_ptr;
    return dr;
}

__device__ __forceinline__ float keep_scale(Dropout dr, long long elem) {
    // counter-based hash (splitmix64 finaliser) -> uniform in [0,1)
    unsigned long long h = dr.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(elem + 1);
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
    h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
    h ^= h >> 31;
    float u = (float)(h >> 40) * (1.0f / 16777216.0f);
    return u >= dr.p ? 1.0f / (1.0f - dr.p) : 0.f;
}

struct Coef4 {
    float4 sc, sh, sl;
    bool has_sl;
};

__device__ __forceinline__ Coef4 load_coef(const float* scale, const float* shift, const float* slope, int c0) {
    Coef4 c;
    c.sc = ld4(scale + c0);
    c.sh = ld4(shift + c0);
    c.has_sl = slope != nullptr;
    c.sl = c.has_sl ? ld4(slope + c0) : make_float4(1.f, 1.f, 1.f, 1.f);
    return c;
}

struct FwdIn {
    float4 z, r;
};

__device__ __forceinline__ FwdIn fwd_load(const float* z, const float* zr, long long idx) {
    FwdIn in;
    in.z = ld4(z + idx);
    in.r = zr ? ld4(zr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    return in;
}

__device__ __forceinline__ float4 fwd_compute(const FwdIn& in, const Coef4& cb, const Coef4& cr, bool has_res,
                                              const Dropout& dr, long long idx) {
    float4 v = affine4(in.z, cb.sc, cb.sh);
    if (has_res) {
        float4 r = affine4(in.r, cr.sc, cr.sh);
        if (cr.has_sl) r = prelu4(r, cr.sl);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (cb.has_sl) v = prelu4(v, cb.sl);
    if (dr.p > 0.f) {
        v.x *= keep_scale(dr, idx); v.y *= keep_scale(dr, idx + 1);
        v.z *= keep_scale(dr, idx + 2); v.w *= keep_scale(dr, idx + 3);
    }
    return v;
}

// POST (eval only): a compile-time switch, so that the training instantiations keep their register allocation
template <bool STATS, bool POST>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ z, Geo g, BnCoef bn, Residual res, Dropout dr_in, void* a_mma,
                  int fmt, float* a_f32, double* out_stats, BnCoef post) {
    const Dropout dr = resolve_seed(dr_in);
    EW_PROLOGUE
    if (!STATS && !cok) return;
    Acc4 acc[2];
    if (STATS) { acc[0].init(); acc[1].init(); }
    if (cok) {
        Coef4 cb = load_coef(bn.scale, bn.shift, bn.slope, c0);
        Coef4 cr = cb;
        const bool has_res = res.zr != nullptr;
        if (has_res) cr = load_coef(res.scale, res.shift, res.slope, c0);
        const long long plane = g.rows * g.Cs;
        // post (eval): the operand planes receive post(v) = the NEXT BatchNorm (+ PReLU) applied to this output
        Coef4 cp = cb;
        if (POST) cp = load_coef(post.scale, post.shift, post.slope, c0);
        auto post_apply = [&](float4 v) {
            if (!POST) return v;
            float4 y = affine4(v, cp.sc, cp.sh);
            return cp.has_sl ? prelu4(y, cp.sl) : y;
        };
        EW_PIXEL_LOOP2 {
            const long long idxA = rowA * g.Cs + c0, idxB = rowB * g.Cs + c0;
            FwdIn inA, inB;
            if (okA) inA = fwd_load(z, res.zr, idxA);
            if (okB) inB = fwd_load(z, res.zr, idxB);
            if (okA) {
                float4 v = fwd_compute(inA, cb, cr, has_res, dr, idxA);
                if (a_f32) st4(a_f32 + idxA, v);
                if (a_mma) store_fmt(a_mma, fmt, plane, idxA, post_apply(v));
                if (STATS) {
                    acc[0].add(v);
                    acc[1].add(make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w));
                }
            }
            if (okB) {
                float4 v = fwd_compute(inB, cb, cr, has_res, dr, idxB);
                if (a_f32) st4(a_f32 + idxB, v);
                if (a_mma) store_fmt(a_mma, fmt, plane, idxB, post_apply(v));
                if (STATS) {
                    acc[0].add(v);
                    acc[1].add(make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w));
                }
            }
        }
    }
    if (STATS) block_reduce_store<2>(acc, out_stats, g.Cs, c0, cok);
}

// The common forward case -- a = act(BN(z)) written as GEMM operand planes only, no residual / dropout / float32 copy /
// statistics -- with FOUR rows in flight per thread: at 70 registers three CTAs fit per SM and two rows of one 16-byte
// load each keep only ~24 KB per SM in flight (68 % of HBM bandwidth measured); four rows double that.
__global__ void __launch_bounds__(256)
bn_act_fwd_simple_kernel(const float* __restrict__ z, Geo g, BnCoef bn, void* a_mma, int fmt) {
    EW_PROLOGUE
    if (!cok) return;
    const Coef4 cb = load_coef(bn.scale, bn.shift, bn.slope, c0);
    const long long plane = g.rows * g.Cs;
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += 4 * stride) {
        long long idx[4];
        bool ok[4];
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long row = row0 + k * stride;
            ok[k] = row < g.rows && (g.mask == nullptr || g.mask[row]);
            idx[k] = row * g.Cs + c0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (ok[k]) v[k] = ld4(z + idx[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (ok[k]) {
                float4 y = affine4(v[k], cb.sc, cb.sh);
                if (cb.has_sl) y = prelu4(y, cb.sl);
                store_fmt(a_mma, fmt, plane, idx[k], y);
            }
    }
}

int bn_act_forward(const float* z, const Geo& g, BnCoef bn, Residual res, Dropout dr, void* a_mma, int fmt,
                   float* a_f32, double* out_stats, cudaStream_t s, const BnCoef* post) {
    EW_CHECK(g);
    EwShape sh = ew_shape(g);
    const BnCoef no_post = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const BnCoef pc = post ? *post : no_post;
    if (!out_stats && !res.zr && dr.p == 0.f && !a_f32 && a_mma && !post) {
        bn_act_fwd_simple_kernel<<<sh.grid, sh.block, 0, s>>>(z, g, bn, a_mma, fmt);
        FSB_LAUNCHED();
        return 0;
    }
    if (post) {
        FSB_REQUIRE(!out_stats, "bn_act_forward: the post map is an eval option (no statistics)");
        bn_act_fwd_kernel<false, true><<<sh.grid, sh.block, 0, s>>>(z, g, bn, res, dr, a_mma, fmt, a_f32, nullptr, pc);
    } else if (out_stats) {
        bn_act_fwd_kernel<true, false><<<sh.grid, sh.block, 0, s>>>(z, g, bn, res, dr, a_mma, fmt, a_f32, out_stats, pc);
    } else {
        bn_act_fwd_kernel<false, false><<<sh.grid, sh.block, 0, s>>>(z, g, bn, res, dr, a_mma, fmt, a_f32, nullptr, pc);
    }
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// amax (optional, training): position 0..3 of the first maximum of every window in scan order, one byte per pooled
// element -- the backward pass routes by it instead of re-reading the four full-resolution planes
template <bool STATS, bool POST>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const float* __restrict__ zf, Geo gf, float* zp, Geo g, int pool_h, double* out_stats,
                   unsigned char* amax, BnCoef post, void* a_mma, int fmt) {
    EW_PROLOGUE
    if (!STATS && !cok) return;
    Acc4 acc[2];
    if (STATS) { acc[0].init(); acc[1].init(); }
    if (cok) {
        Coef4 cp;
        if (POST) cp = load_coef(post.scale, post.shift, post.slope, c0);
        const long long plane = g.rows * g.Cs;
        EW_PIXEL_LOOP {
            const unsigned img = (unsigned)(g.Hp * g.Wp);
            const int n = (int)((unsigned long long)row / img);
            const unsigned rr = (unsigned)(row - (long long)n * img);
            const int y = (int)(rr / (unsigned)g.Wp) - g.padH, x = (int)(rr % (unsigned)g.Wp) - g.padW;
            long long r00 = geo_row(gf, n, y * pool_h, 2 * x);
            const float4 v0 = ld4(zf + r00 * gf.Cs + c0), v1 = ld4(zf + (r00 + 1) * gf.Cs + c0);
            float4 v2 = v0, v3 = v0;
            float4 m = max4(v0, v1);
            if (pool_h == 2) {
                long long r10 = r00 + gf.Wp;
                v2 = ld4(zf + r10 * gf.Cs + c0);
                v3 = ld4(zf + (r10 + 1) * gf.Cs + c0);
                m = max4(m, max4(v2, v3));
            }
            st4(zp + row * g.Cs + c0, m);
            if (POST) {                  // eval: the BatchNorm + PReLU that follows, straight into the operand planes
                float4 y = affine4(m, cp.sc, cp.sh);
                if (cp.has_sl) y = prelu4(y, cp.sl);
                store_fmt(a_mma, fmt, plane, row * g.Cs + c0, y);
            }
            if (amax) {
                const float a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w};
                const float a2[4] = {v2.x, v2.y, v2.z, v2.w}, a3[4] = {v3.x, v3.y, v3.z, v3.w};
                unsigned packed = 0u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    unsigned best = 0u;
                    float bv = a0[i];
                    if (a1[i] > bv) { bv = a1[i]; best = 1u; }
                    if (pool_h == 2) {
                        if (a2[i] > bv) { bv = a2[i]; best = 2u; }
                        if (a3[i] > bv) { bv = a3[i]; best = 3u; }
                    }
                    packed |= best << (8 * i);
                }
                *reinterpret_cast<unsigned*>(amax + row * g.Cs + c0) = packed;
            }
            if (STATS) {
                acc[0].add(m);
                acc[1].add(make_float4(m.x * m.x, m.y * m.y, m.z * m.z, m.w * m.w));
            }
        }
    }
    if (STATS) block_reduce_store<2>(acc, out_stats, g.Cs, c0, cok);
}

int maxpool_forward(const float* zf, const Geo& gf, float* zp, const Geo& gp, int pool_h, double* out_stats,
                    unsigned char* amax, cudaStream_t s, const BnCoef* post, void* a_mma, int fmt) {
    EW_CHECK(gp);
    FSB_REQUIRE(!post == !a_mma, "maxpool_forward: the fused BatchNorm output needs both its coefficients and its planes");
    EwShape sh = ew_shape(gp);
    const BnCoef no_post = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const BnCoef pc = post ? *post : no_post;
    if (post) {
        FSB_REQUIRE(!out_stats, "maxpool_forward: the fused BatchNorm output is an eval option (no statistics)");
        maxpool_fwd_kernel<false, true><<<sh.grid, sh.block, 0, s>>>(zf, gf, zp, gp, pool_h, nullptr, amax, pc, a_mma, fmt);
    } else if (out_stats) {
        maxpool_fwd_kernel<true, false><<<sh.grid, sh.block, 0, s>>>(zf, gf, zp, gp, pool_h, out_stats, amax, pc, a_mma, fmt);
    } else {
        maxpool_fwd_kernel<false, false><<<sh.grid, sh.block, 0, s>>>(zf, gf, zp, gp, pool_h, nullptr, amax, pc, a_mma, fmt);
    }
    FSB_LAUNCHED();
    return 0;
}

// One thread per POOLED pixel and float4 of channels: it reads its window of zf once (2 or 4 pixels), routes the
// gradient to the first maximum in scan order and writes all window positions of dzf (zeros elsewhere).  Full-res
// pixels that no window covers (odd trailing row / column, floor mode) get zeros from the thread of the adjacent
// window.  (The previous version walked full-res pixels and re-read the window four times: 3.4 TB/s.)
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ dzp, Geo gp, const float* __restrict__ zf, Geo g, int pool_h,
                   void* dzf, int fmt, const unsigned* absmax) {
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv >= g.Cs / 4) return;
    const float gscale = gs_scale(absmax);
    const int c0 = cv * 4;
    const long long plane = g.rows * g.Cs;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool odd_w = g.W > 2 * gp.W, odd_h = pool_h == 2 && g.H > 2 * gp.H;
    for (long long prow = (long long)blockIdx.x * blockDim.y + threadIdx.y; prow < gp.rows;
         prow += (long long)gridDim.x * blockDim.y) {
        if (gp.mask != nullptr && !gp.mask[prow]) continue;
        const unsigned img = (unsigned)(gp.Hp * gp.Wp);
        const int n = (int)((unsigned long long)prow / img);
        const unsigned rr = (unsigned)(prow - (long long)n * img);
        const int py = (int)(rr / (unsigned)gp.Wp) - gp.padH, px = (int)(rr % (unsigned)gp.Wp) - gp.padW;
        const long long r00 = geo_row(g, n, py * pool_h, 2 * px);
        const float4 gsrc = ld4(dzp + prow * gp.Cs + c0);
        const float4 v0 = ld4(zf + r00 * g.Cs + c0), v1 = ld4(zf + (r00 + 1) * g.Cs + c0);
        float4 v2 = zero, v3 = zero;
        if (pool_h == 2) {
            v2 = ld4(zf + (r00 + g.Wp) * g.Cs + c0);
            v3 = ld4(zf + (r00 + g.Wp + 1) * g.Cs + c0);
        }
        const float a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w};
        const float a2[4] = {v2.x, v2.y, v2.z, v2.w}, a3[4] = {v3.x, v3.y, v3.z, v3.w};
        const float gs[4] = {gsrc.x, gsrc.y, gsrc.z, gsrc.w};
        float o[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int best = 0;
            float bv = a0[i];
            if (a1[i] > bv) { bv = a1[i]; best = 1; }
            if (pool_h == 2) {
                if (a2[i] > bv) { bv = a2[i]; best = 2; }
                if (a3[i] > bv) { bv = a3[i]; best = 3; }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][i] = best == k ? gs[i] : 0.f;
        }
        store_fmt(dzf, fmt, plane, r00 * g.Cs + c0, make_float4(o[0][0], o[0][1], o[0][2], o[0][3]), gscale);
        store_fmt(dzf, fmt, plane, (r00 + 1) * g.Cs + c0, make_float4(o[1][0], o[1][1], o[1][2], o[1][3]), gscale);
        if (pool_h == 2) {
            store_fmt(dzf, fmt, plane, (r00 + g.Wp) * g.Cs + c0, make_float4(o[2][0], o[2][1], o[2][2], o[2][3]), gscale);
            store_fmt(dzf, fmt, plane, (r00 + g.Wp + 1) * g.Cs + c0, make_float4(o[3][0], o[3][1], o[3][2], o[3][3]), gscale);
        }
        // uncovered trailing column / row / corner
        const bool last_x = odd_w && px == gp.W - 1, last_y = odd_h && py == gp.H - 1;
        if (last_x) {
            store_fmt(dzf, fmt, plane, (r00 + 2) * g.Cs + c0, zero);
            if (pool_h == 2) store_fmt(dzf, fmt, plane, (r00 + g.Wp + 2) * g.Cs + c0, zero);
        }
        if (last_y) {
            store_fmt(dzf, fmt, plane, (r00 + 2 * g.Wp) * g.Cs + c0, zero);
            store_fmt(dzf, fmt, plane, (r00 + 2 * g.Wp + 1) * g.Cs + c0, zero);
            if (last_x) store_fmt(dzf, fmt, plane, (r00 + 2 * g.Wp + 2) * g.Cs + c0, zero);
        }
    }
}

int maxpool_backward(const float* dzp, const Geo& gp, const float* zf, const Geo& gf, int pool_h, void* dzf,
                     int fmt, const unsigned* absmax, cudaStream_t s) {
    EW_CHECK(gf);
    EW_CHECK(gp);
    FSB_REQUIRE(gf.W - 2 * gp.W <= 1 && gf.W >= 2 * gp.W && (pool_h == 1 ? gf.H == gp.H : (gf.H - 2 * gp.H <= 1 && gf.H >= 2 * gp.H)),
                "maxpool_backward: geometries are not a floor-mode 2x pooling pair");
    EwShape sh = ew_shape(gp);
    maxpool_bwd_kernel<<<sh.grid, sh.block, 0, s>>>(dzp, gp, zf, gf, pool_h, dzf, fmt, absmax);
    FSB_LAUNCHED();
    return 0;
}

// Same routing from the stored arg-max bytes: reads dzp (float32, or a half plane that already carries the GradScale
// of `absmax`) and one byte per pooled element instead of the four full-resolution planes of zf.
__global__ void __launch_bounds__(256)
maxpool_bwd_amax_kernel(const GradRef dzp, Geo gp, const unsigned char* __restrict__ amax, Geo g, int pool_h,
                        void* dzf, int fmt, const unsigned* absmax) {
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv >= g.Cs / 4) return;
    // half source: stored 2^k1 g, destination wants 2^k2 g (k1 == k2 when both use the same slot: exact copy)
    const float gscale = gs_scale(absmax) * (dzp.half ? gs_pow2(-gs_exponent2(dzp.bits, dzp.mul)) : 1.f);
    const int c0 = cv * 4;
    const long long plane = g.rows * g.Cs;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool odd_w = g.W > 2 * gp.W, odd_h = pool_h == 2 && g.H > 2 * gp.H;
    for (long long prow = (long long)blockIdx.x * blockDim.y + threadIdx.y; prow < gp.rows;
         prow += (long long)gridDim.x * blockDim.y) {
        if (gp.mask != nullptr && !gp.mask[prow]) continue;
        const unsigned img = (unsigned)(gp.Hp * gp.Wp);
        const int n = (int)((unsigned long long)prow / img);
        const unsigned rr = (unsigned)(prow - (long long)n * img);
        const int py = (int)(rr / (unsigned)gp.Wp) - gp.padH, px = (int)(rr % (unsigned)gp.Wp) - gp.padW;
        const long long r00 = geo_row(g, n, py * pool_h, 2 * px);
        const float4 gsrc = ld_grad(dzp, prow * gp.Cs + c0);
        const unsigned am = *reinterpret_cast<const unsigned*>(amax + prow * gp.Cs + c0);
        const float gs[4] = {gsrc.x, gsrc.y, gsrc.z, gsrc.w};
        float o[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int best = (int)((am >> (8 * i)) & 0xFFu);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][i] = best == k ? gs[i] : 0.f;
        }
        store_fmt(dzf, fmt, plane, r00 * g.Cs + c0, make_float4(o[0][0], o[0][1], o[0][2], o[0][3]), gscale);
        store_fmt(dzf, fmt, plane, (r00 + 1) * g.Cs + c0, make_float4(o[1][0], o[1][1], o[1][2], o[1][3]), gscale);
        if (pool_h == 2) {
            store_fmt(dzf, fmt, plane, (r00 + g.Wp) * g.Cs + c0, make_float4(o[2][0], o[2][1], o[2][2], o[2][3]), gscale);
            store_fmt(dzf, fmt, plane, (r00 + g.Wp + 1) * g.Cs + c0, make_float4(o[3][0], o[3][1], o[3][2], o[3][3]), gscale);
        }
        const bool last_x = odd_w && px == gp.W - 1, last_y = odd_h && py == gp.H - 1;
        if (last_x) {
            store_fmt(dzf, fmt, plane, (r00 + 2) * g.Cs + c0, zero);
            if (pool_h == 2) store_fmt(dzf, fmt, plane, (r00 + g.Wp + 2) * g.Cs + c0, zero);
        }
        if (last_y) {
            store_fmt(dzf, fmt, plane, (r00 + 2 * g.Wp) * g.Cs + c0, zero);
            store_fmt(dzf, fmt, plane, (r00 + 2 * g.Wp + 1) * g.Cs + c0, zero);
            if (last_x) store_fmt(dzf, fmt, plane, (r00 + 2 * g.Wp + 2) * g.Cs + c0, zero);
        }
    }
}

// Half plane in, half plane out, same GradScale: eight channels per thread, the routing is a byte compare and a mask
// (no conversion); every access is 16 bytes (8 for the arg-max bytes).
__global__ void __launch_bounds__(256)
maxpool_bwd_amax8_kernel(const __half* __restrict__ dzp, Geo gp, const unsigned char* __restrict__ amax, Geo g, int pool_h,
                         __half* dzf) {
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv >= g.Cs / 8) return;
    const int c0 = cv * 8;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const bool odd_w = g.W > 2 * gp.W, odd_h = pool_h == 2 && g.H > 2 * gp.H;
    for (long long prow = (long long)blockIdx.x * blockDim.y + threadIdx.y; prow < gp.rows;
         prow += (long long)gridDim.x * blockDim.y) {
        if (gp.mask != nullptr && !gp.mask[prow]) continue;
        const unsigned img = (unsigned)(gp.Hp * gp.Wp);
        const int n = (int)((unsigned long long)prow / img);
        const unsigned rr = (unsigned)(prow - (long long)n * img);
        const int py = (int)(rr / (unsigned)gp.Wp) - gp.padH, px = (int)(rr % (unsigned)gp.Wp) - gp.padW;
        const long long r00 = geo_row(g, n, py * pool_h, 2 * px);
        const uint4 gsrc = *reinterpret_cast<const uint4*>(dzp + prow * gp.Cs + c0);
        const uint2 am = *reinterpret_cast<const uint2*>(amax + prow * gp.Cs + c0);
        uint4 o[4];
#pragma unroll
        for (unsigned pos = 0; pos < 4; ++pos) {
            const unsigned m_lo = __vcmpeq4(am.x, pos * 0x01010101u), m_hi = __vcmpeq4(am.y, pos * 0x01010101u);
            o[pos] = make_uint4(gsrc.x & __byte_perm(m_lo, 0u, 0x1100u), gsrc.y & __byte_perm(m_lo, 0u, 0x3322u),
                                gsrc.z & __byte_perm(m_hi, 0u, 0x1100u), gsrc.w & __byte_perm(m_hi, 0u, 0x3322u));
        }
        *reinterpret_cast<uint4*>(dzf + r00 * g.Cs + c0) = o[0];
        *reinterpret_cast<uint4*>(dzf + (r00 + 1) * g.Cs + c0) = o[1];
        if (pool_h == 2) {
            *reinterpret_cast<uint4*>(dzf + (r00 + g.Wp) * g.Cs + c0) = o[2];
            *reinterpret_cast<uint4*>(dzf + (r00 + g.Wp + 1) * g.Cs + c0) = o[3];
        }
        const bool last_x = odd_w && px == gp.W - 1, last_y = odd_h && py == gp.H - 1;
        if (last_x) {
            *reinterpret_cast<uint4*>(dzf + (r00 + 2) * g.Cs + c0) = zero;
            if (pool_h == 2) *reinterpret_cast<uint4*>(dzf + (r00 + g.Wp + 2) * g.Cs + c0) = zero;
        }
        if (last_y) {
            *reinterpret_cast<uint4*>(dzf + (r00 + 2 * g.Wp) * g.Cs + c0) = zero;
            *reinterpret_cast<uint4*>(dzf + (r00 + 2 * g.Wp + 1) * g.Cs + c0) = zero;
            if (last_x) *reinterpret_cast<uint4*>(dzf + (r00 + 2 * g.Wp + 2) * g.Cs + c0) = zero;
        }
    }
}

int maxpool_backward_amax(GradRef dzp, const Geo& gp, const unsigned char* amax, const Geo& gf, int pool_h, void* dzf,
                          int fmt, const unsigned* absmax, cudaStream_t s) {
    EW_CHECK(gf);
    EW_CHECK(gp);
    if (dzp.half && fmt == FMT_H16 && dzp.bits == absmax && dzp.mul == nullptr && gf.Cs % 8 == 0 && gf.Cs == gp.Cs &&
        gf.W - 2 * gp.W <= 1 && gf.W >= 2 * gp.W && (pool_h == 1 ? gf.H == gp.H : (gf.H - 2 * gp.H <= 1 && gf.H >= 2 * gp.H))) {
        const int cv = gf.Cs / 8, bx = cv < 32 ? cv : 32, by = 256 / bx;
        dim3 grid(ew_shape(gp).grid.x, (cv + bx - 1) / bx);
        maxpool_bwd_amax8_kernel<<<grid, dim3(bx, by), 0, s>>>((const __half*)dzp.p, gp, amax, gf, pool_h, (__half*)dzf);
        FSB_LAUNCHED();
        return 0;
    }
    FSB_REQUIRE(gf.W - 2 * gp.W <= 1 && gf.W >= 2 * gp.W && (pool_h == 1 ? gf.H == gp.H : (gf.H - 2 * gp.H <= 1 && gf.H >= 2 * gp.H)),
                "maxpool_backward: geometries are not a floor-mode 2x pooling pair");
    EwShape sh = ew_shape(gp);
    maxpool_bwd_amax_kernel<<<sh.grid, sh.block, 0, s>>>(dzp, gp, amax, gf, pool_h, dzf, fmt, absmax);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// global max: grid (channel chunks, N, pixel slices); block (bx cvec, by pixels).  Every CTA reduces its slice of
// the image and merges into packed[n][c] with a 64-bit atomicMax of (ordered value bits << 32 | ~row): the winner is
// the largest value and, among equal values, the smallest row -- the first maximum in scan order, independent
// of scheduling.  A second tiny kernel unpacks value and row.
__device__ __forceinline__ unsigned long long gmax_pack(float v, int row) {
    unsigned u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    if (v != v) u = 0xFFFFFFFFu;          // NaN is the maximum (torch.max propagates it: a diverged run stays visible)
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)row);
}

__global__ void __launch_bounds__(256)
gmax_fwd_kernel(const float* __restrict__ x, Geo g, unsigned long long* packed) {
    const int cv = blockIdx.x * blockDim.x + threadIdx.x;
    const bool cok = cv < g.Cs / 4;
    const int c0 = cv * 4;
    const int n = blockIdx.y;
    float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bi[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    const int hw = g.H * g.W;
    const int per = (hw + gridDim.z - 1) / gridDim.z;
    const int p_begin = blockIdx.z * per, p_end = min(hw, p_begin + per);
    if (cok) {
        for (int p = p_begin + threadIdx.y; p < p_end; p += blockDim.y) {
            int yy = p / g.W, xx = p - yy * g.W;
            long long row = geo_row(g, n, yy, xx);
            float4 v4 = ld4(x + row * g.Cs + c0);
            float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (v[i] > bv[i] || (v[i] != v[i] && bv[i] == bv[i]) || bi[i] == 0x7fffffff) { bv[i] = v[i]; bi[i] = (int)row; }
        }
    }
    __shared__ float sv[256 * 4];
    __shared__ int si[256 * 4];
    int t = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) { sv[t * 4 + i] = bv[i]; si[t * 4 + i] = bi[i]; }
    __syncthreads();
    if (threadIdx.y == 0 && cok) {
        for (int i = 0; i < 4; ++i) {
            float b = -INFINITY;
            int r = 0x7fffffff;
            for (int y = 0; y < (int)blockDim.y; ++y) {
                int tt = (y * blockDim.x + threadIdx.x) * 4 + i;
                // first maximum in scan order: larger value wins, ties go to the smaller row
                const bool nan_new = sv[tt] != sv[tt], nan_old = b != b;
                if (si[tt] != 0x7fffffff &&
                    (r == 0x7fffffff || (nan_new && (!nan_old || si[tt] < r)) ||
                     (!nan_old && (sv[tt] > b || (sv[tt] == b && si[tt] < r))))) { b = sv[tt]; r = si[tt]; }
            }
            int c = c0 + i;
            if (c < g.C && r != 0x7fffffff) atomicMax(packed + (long long)n * g.Cs + c, gmax_pack(b, r));
        }
    }
}

__global__ void gmax_unpack_kernel(const unsigned long long* __restrict__ packed, Geo g, float* feat, int feat_stride,
                                   int feat_off, int* argrow) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)g.N * g.C) return;
    int n = (int)(i / g.C), c = (int)(i % g.C);
    unsigned long long pk = packed[(long long)n * g.Cs + c];
    unsigned u = (unsigned)(pk >> 32);
    const bool is_nan = u == 0xFFFFFFFFu;
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    feat[(long long)n * feat_stride + feat_off + c] = is_nan ? __int_as_float(0x7FC00000) : __uint_as_float(u);
    argrow[i] = (int)(0xFFFFFFFFu - (unsigned)(pk & 0xFFFFFFFFull));
}

size_t gmax_scratch_bytes(const Geo& g) { return (size_t)g.N * g.Cs * sizeof(unsigned long long); }

int gmax_forward(const float* x, const Geo& g, float* feat, int feat_stride, int feat_off, int* argrow, void* scratch,
                 cudaStream_t s) {
    int cv = g.Cs / 4;
    int bx = cv < 32 ? cv : 32;
    int by = 256 / bx;
    int chunks = (cv + bx - 1) / bx;
    // enough pixel slices for ~4 CTAs per SM, at least `by` pixels per slice
    int hw = g.H * g.W;
    int slices = (4 * 148 + chunks * g.N - 1) / (chunks * g.N);
    if (slices > (hw + 4 * by - 1) / (4 * by)) slices = (hw + 4 * by - 1) / (4 * by);
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    FSB_CUDA(cudaMemsetAsync(scratch, 0, gmax_scratch_bytes(g), s));
    dim3 grid(chunks, g.N, slices);
    gmax_fwd_kernel<<<grid, dim3(bx, by), 0, s>>>(x, g, (unsigned long long*)scratch);
    FSB_LAUNCHED();
    long long total = (long long)g.N * g.C;
    gmax_unpack_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>((const unsigned long long*)scratch, g, feat, feat_stride,
                                                                  feat_off, argrow);
    FSB_LAUNCHED();
    return 0;
}

__global__ void gmax_bwd_kernel(const float* dfeat, int feat_stride, int feat_off, const int* argrow, Geo g,
                                float* dx) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)g.N * g.C) return;
    int n = (int)(i / g.C), c = (int)(i % g.C);
    long long row = argrow[i];
    dx[row * g.Cs + c] += dfeat[(long long)n * feat_stride + feat_off + c];
}

int gmax_backward(const float* dfeat, int feat_stride, int feat_off, const int* argrow, const Geo& g,
                  float* dx, cudaStream_t s) {
    long long total = (long long)g.N * g.C;
    gmax_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(dfeat, feat_stride, feat_off, argrow, g, dx);
    FSB_LAUNCHED();
    return 0;
}

// the same scatter-add into a scaled half plane (compact backward: dx holds 2^k * gradient, k from `bits`)
__global__ void gmax_bwd_h16_kernel(const float* dfeat, int feat_stride, int feat_off, const int* argrow, Geo g,
                                    __half* dx, const unsigned* bits) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)g.N * g.C) return;
    int n = (int)(i / g.C), c = (int)(i % g.C);
    long long row = argrow[i];
    const float scale = gs_scale(bits);
    __half* d = dx + row * g.Cs + c;
    *d = __float2half_rn(fmaf(dfeat[(long long)n * feat_stride + feat_off + c], scale, __half2float(*d)));
}

int gmax_backward_h16(const float* dfeat, int feat_stride, int feat_off, const int* argrow, const Geo& g, void* dx,
                      const unsigned* bits, cudaStream_t s) {
    long long total = (long long)g.N * g.C;
    gmax_bwd_h16_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(dfeat, feat_stride, feat_off, argrow, g, (__half*)dx, bits);
    FSB_LAUNCHED();
    return 0;
}

// bits[0] = float32 bit pattern of max |x| over n values (single CTA; a few thousand head gradients)
__global__ void __launch_bounds__(256) absmax_bits_kernel(const float* x, long long n, unsigned* bits) {
    float m = 0.f;
    for (long long i = threadIdx.x; i < n; i += 256) m = fmaxf(m, fabsf(x[i]));
    __shared__ float red[256];
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) bits[0] = __float_as_uint(red[0]);
}

int absmax_bits(const float* x, long long n, unsigned* bits, cudaStream_t s) {
    absmax_bits_kernel<<<1, 256, 0, s>>>(x, n, bits);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// backward of a = act(BN(z) [+ r]) : shared recomputation of dy (and the PReLU slope term).
// These 4-channel kernels serve the float32-plane backward (strict / float32 modes, FSB200_COMPACT_BWD=0, the FC head);
// incoming gradients may be float32 planes or scaled half planes (GradRef).  The compact backward of the mixed mode runs
// the 8-channel kernels further down.
struct BwdCoef {
    Coef4 cb, cr;
    float4 mean, invstd;
    float inv1, inv2;         // inverse GradScale of the two incoming gradient planes (1 for float32 planes)
    bool has_res;
};

struct BwdIn {
    float4 z, r, g, g2;
};


// RES: the activation has a residual branch (res.zr); DA2: the incoming gradient is dA1 + dA2
template <bool RES, bool DA2>
__device__ __forceinline__ BwdIn bwd_load(const GradRef& dA1, const GradRef& dA2, const float* z, const float* zr,
                                          long long idx) {
    BwdIn in;
    in.z = ld4(z + idx);
    in.g = ld_grad(dA1, idx);
    if (RES) in.r = ld4(zr + idx);
    if (DA2) in.g2 = ld_grad(dA2, idx);
    return in;
}

template <bool RES, bool DA2>
__device__ __forceinline__ void bwd_compute(const BwdIn& in, const BwdCoef& k, const Dropout& dr, long long idx,
                                            float4& dy, float4& zhat, float4& dsl) {
    const float4 zz = in.z;
    float4 y = affine4(zz, k.cb.sc, k.cb.sh);
    if (RES) {
        float4 r = affine4(in.r, k.cr.sc, k.cr.sh);
        if (k.cr.has_sl) r = prelu4(r, k.cr.sl);
        y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    float4 g = make_float4(in.g.x * k.inv1, in.g.y * k.inv1, in.g.z * k.inv1, in.g.w * k.inv1);
    if (DA2) { g.x = fmaf(in.g2.x, k.inv2, g.x); g.y = fmaf(in.g2.y, k.inv2, g.y); g.z = fmaf(in.g2.z, k.inv2, g.z); g.w = fmaf(in.g2.w, k.inv2, g.w); }
    if (dr.p > 0.f) {
        g.x *= keep_scale(dr, idx); g.y *= keep_scale(dr, idx + 1);
        g.z *= keep_scale(dr, idx + 2); g.w *= keep_scale(dr, idx + 3);
    }
    if (k.cb.has_sl) {
        dsl = make_float4(y.x > 0.f ? 0.f : y.x * g.x, y.y > 0.f ? 0.f : y.y * g.y,
                          y.z > 0.f ? 0.f : y.z * g.z, y.w > 0.f ? 0.f : y.w * g.w);
        dy = make_float4(y.x > 0.f ? g.x : k.cb.sl.x * g.x, y.y > 0.f ? g.y : k.cb.sl.y * g.y,
                         y.z > 0.f ? g.z : k.cb.sl.z * g.z, y.w > 0.f ? g.w : k.cb.sl.w * g.w);
    } else {
        dsl = make_float4(0.f, 0.f, 0.f, 0.f);
        dy = g;
    }
    zhat = make_float4((zz.x - k.mean.x) * k.invstd.x, (zz.y - k.mean.y) * k.invstd.y,
                       (zz.z - k.mean.z) * k.invstd.z, (zz.w - k.mean.w) * k.invstd.w);
}

template <bool RES>
__device__ __forceinline__ BwdCoef load_bwd(const BnCoef& bn, const Residual& res, const GradRef& dA1, const GradRef& dA2,
                                            int c0) {
    BwdCoef k;
    k.cb = load_coef(bn.scale, bn.shift, bn.slope, c0);
    k.has_res = RES;
    if (RES) k.cr = load_coef(res.scale, res.shift, res.slope, c0);
    k.mean = ld4(bn.mean + c0);
    k.invstd = ld4(bn.invstd + c0);
    k.inv1 = dA1.half ? gs_pow2(-gs_exponent2(dA1.bits, dA1.mul)) : 1.f;
    k.inv2 = (dA2.p && dA2.half) ? gs_pow2(-gs_exponent2(dA2.bits, dA2.mul)) : 1.f;
    return k;
}

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void fma4(float4& a, const float4& b, const float4& c) {
    a.x = fmaf(b.x, c.x, a.x); a.y = fmaf(b.y, c.y, a.y); a.z = fmaf(b.z, c.z, a.z); a.w = fmaf(b.w, c.w, a.w);
}

bool bn_bwd_compact_ok(const GradRef& dA1, const GradRef& dA2, const Geo& g, const Residual& res, const Dropout& dr);
static int bn_bwd_c8_reduce(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn,
                            double* partials, cudaStream_t s);
static int bn_bwd_c8_apply(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn,
                           const float* c1, const float* c2, void* dz, int fmt, const unsigned* absmax, cudaStream_t s);

// Per-thread float32 partial sums (a thread visits rows / (gridDim.x * blockDim.y) ~ 10^2 pixels); the cross-thread
// and cross-block reductions run in double.
template <bool RES, bool DA2>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_reduce_kernel(const GradRef dA1, const GradRef dA2, const float* __restrict__ z, Geo g,
                         BnCoef bn, Residual res, Dropout dr_in, double* partials) {
    const Dropout dr = resolve_seed(dr_in);
    EW_PROLOGUE
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0;
    float4 mx = s0, zx = s0;              // max |dy|, max |zhat|: bound of |dz| for the half-precision gradient scale
    if (cok) {
        const BwdCoef k = load_bwd<RES>(bn, res, dA1, dA2, c0);
        // ROWS rows in flight per thread (two loads each in the plain variant): see bn_act_fwd_simple_kernel
        constexpr int ROWS = (RES || DA2) ? 2 : 4;
        const long long stride = (long long)gridDim.x * blockDim.y;
        for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += ROWS * stride) {
            long long idx[ROWS];
            bool ok[ROWS];
            BwdIn in[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                const long long row = row0 + j * stride;
                ok[j] = row < g.rows && (g.mask == nullptr || g.mask[row]);
                idx[j] = row * g.Cs + c0;
            }
#pragma unroll
            for (int j = 0; j < ROWS; ++j)
                if (ok[j]) in[j] = bwd_load<RES, DA2>(dA1, dA2, z, res.zr, idx[j]);
#pragma unroll
            for (int j = 0; j < ROWS; ++j)
                if (ok[j]) {
                    float4 dy, zh, dsl;
                    bwd_compute<RES, DA2>(in[j], k, dr, idx[j], dy, zh, dsl);
                    add4(s0, dy); fma4(s1, dy, zh); add4(s2, dsl);
                    mx = make_float4(fmaxf(mx.x, fabsf(dy.x)), fmaxf(mx.y, fabsf(dy.y)), fmaxf(mx.z, fabsf(dy.z)),
                                     fmaxf(mx.w, fabsf(dy.w)));
                    zx = make_float4(fmaxf(zx.x, fabsf(zh.x)), fmaxf(zx.y, fabsf(zh.y)), fmaxf(zx.z, fabsf(zh.z)),
                                     fmaxf(zx.w, fabsf(zh.w)));
                }
        }
    }
    Acc4 acc[5];
    acc[0].f = s0; acc[1].f = s1; acc[2].f = s2; acc[3].f = mx; acc[4].f = zx;
    block_reduce_store<5, 2>(acc, partials, g.Cs, c0, cok);
}

int bn_act_bwd_reduce(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn,
                      Residual res, Dropout dr, double* partials, cudaStream_t s) {
    EW_CHECK(g);
    FSB_REQUIRE(!(res.zr && dA2.p), "bn_act_bwd: residual and second gradient are mutually exclusive");
    if (bn_bwd_compact_ok(dA1, dA2, g, res, dr)) return bn_bwd_c8_reduce(dA1, dA2, z, a_hi, g, bn, partials, s);
    EwShape sh = ew_shape(g);
    if (res.zr)
        bn_act_bwd_reduce_kernel<true, false><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, g, bn, res, dr, partials);
    else if (dA2.p)
        bn_act_bwd_reduce_kernel<false, true><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, g, bn, res, dr, partials);
    else
        bn_act_bwd_reduce_kernel<false, false><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, g, bn, res, dr, partials);
    FSB_LAUNCHED();
    return 0;
}

__global__ void __launch_bounds__(FIN_THREADS)
bn_bwd_finalize_kernel(const double* partials, int nblk, long long count, int C, int Cs, const float* bn_scale,
                       float* dgamma, float* dbeta, float* dslope, float* c1, float* c2, unsigned* absmax,
                       unsigned* absmax_dy, const unsigned* extra_bits) {
    const int c = blockIdx.x * FIN_CH + (threadIdx.x % FIN_CH);
    double tot[5];
    reduce_partials<5, 2>(partials, nblk, Cs, c, tot);
    if (threadIdx.x >= FIN_CH || c >= Cs) return;
    if (c >= C) { c1[c] = 0.f; c2[c] = 0.f; return; }
    if (dbeta) dbeta[c] = (float)tot[0];
    if (dgamma) dgamma[c] = (float)tot[1];
    if (dslope) dslope[c] = (float)tot[2];
    const float m1 = (float)(tot[0] / (double)count), m2 = (float)(tot[1] / (double)count);
    c1[c] = m1;
    c2[c] = m2;
    if (absmax) {
        // |dz| = |scale (dy - c1 - zhat c2)| <= |scale| (max|dy| + |c1| + max|zhat| |c2|); 1.001: float32 rounding of
        // the apply pass.  Non-negative floats order like their bit patterns, and max is order independent.
        // extra_bits (optional): bound of a term added to dz afterwards (global-max scatter into the same plane)
        const float bound = 1.001f * fabsf(bn_scale[c]) * ((float)tot[3] + fabsf(m1) + (float)tot[4] * fabsf(m2)) +
                            (extra_bits ? 1.001f * __uint_as_float(*extra_bits) : 0.f);
        if (bound > 0.f) atomicMax(absmax, __float_as_uint(bound));
    }
    if (absmax_dy) {                      // bound of the residual-branch gradient dres = dy (written as a half plane)
        const float bound = 1.001f * (float)tot[3];
        if (bound > 0.f) atomicMax(absmax_dy, __float_as_uint(bound));
    }
}

int bn_bwd_finalize(const double* partials, int nblk, long long count, int C, int Cs, const float* bn_scale,
                    float* dgamma, float* dbeta, float* dslope, float* c1, float* c2, unsigned* absmax,
                    unsigned* absmax_dy, const unsigned* extra_bits, cudaStream_t s) {
    bn_bwd_finalize_kernel<<<(Cs + FIN_CH - 1) / FIN_CH, FIN_THREADS, 0, s>>>(partials, nblk, count, C, Cs, bn_scale, dgamma,
                                                                              dbeta, dslope, c1, c2, absmax, absmax_dy,
                                                                              extra_bits);
    FSB_LAUNCHED();
    return 0;
}

// dres: float32 plane, or (dres_bits != nullptr) one half plane with the GradScale of *dres_bits
__device__ __forceinline__ void store_dres(void* dres, const unsigned* dres_bits, float dscale, long long idx, float4 dy) {
    if (dres_bits) {
        __half2 a = __floats2half2_rn(dy.x * dscale, dy.y * dscale), b = __floats2half2_rn(dy.z * dscale, dy.w * dscale);
        uint2 u;
        u.x = *reinterpret_cast<unsigned*>(&a);
        u.y = *reinterpret_cast<unsigned*>(&b);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(dres) + idx) = u;
    } else {
        st4(reinterpret_cast<float*>(dres) + idx, dy);
    }
}

template <bool RES, bool DA2>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_apply_kernel(const GradRef dA1, const GradRef dA2, const float* __restrict__ z, Geo g,
                        BnCoef bn, Residual res, Dropout dr_in, const float* c1, const float* c2, void* dz, int fmt,
                        void* dres, const unsigned* dres_bits, const unsigned* absmax) {
    const Dropout dr = resolve_seed(dr_in);
    EW_PROLOGUE
    if (!cok) return;
    const float gscale = gs_scale(absmax);
    const float dscale = gs_scale(dres_bits);
    const BwdCoef k = load_bwd<RES>(bn, res, dA1, dA2, c0);
    float4 m1 = ld4(c1 + c0), m2 = ld4(c2 + c0);
    const long long plane = g.rows * g.Cs;
    EW_PIXEL_LOOP2 {
        const long long idxA = rowA * g.Cs + c0, idxB = rowB * g.Cs + c0;
        BwdIn inA, inB;
        if (okA) inA = bwd_load<RES, DA2>(dA1, dA2, z, res.zr, idxA);
        if (okB) inB = bwd_load<RES, DA2>(dA1, dA2, z, res.zr, idxB);
        float4 dy, zh, dsl;
        if (okA) {
            bwd_compute<RES, DA2>(inA, k, dr, idxA, dy, zh, dsl);
            float4 o = make_float4(k.cb.sc.x * (dy.x - m1.x - zh.x * m2.x), k.cb.sc.y * (dy.y - m1.y - zh.y * m2.y),
                                   k.cb.sc.z * (dy.z - m1.z - zh.z * m2.z), k.cb.sc.w * (dy.w - m1.w - zh.w * m2.w));
            store_fmt(dz, fmt, plane, idxA, o, gscale);
            if (RES && dres) store_dres(dres, dres_bits, dscale, idxA, dy);
        }
        if (okB) {
            bwd_compute<RES, DA2>(inB, k, dr, idxB, dy, zh, dsl);
            float4 o = make_float4(k.cb.sc.x * (dy.x - m1.x - zh.x * m2.x), k.cb.sc.y * (dy.y - m1.y - zh.y * m2.y),
                                   k.cb.sc.z * (dy.z - m1.z - zh.z * m2.z), k.cb.sc.w * (dy.w - m1.w - zh.w * m2.w));
            store_fmt(dz, fmt, plane, idxB, o, gscale);
            if (RES && dres) store_dres(dres, dres_bits, dscale, idxB, dy);
        }
    }
}

int bn_act_bwd_apply(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn, Residual res,
                     Dropout dr, const float* c1, const float* c2, void* dz, int fmt, void* dres,
                     const unsigned* dres_bits, const unsigned* absmax, cudaStream_t s) {
    EW_CHECK(g);
    FSB_REQUIRE(!(res.zr && dA2.p), "bn_act_bwd: residual and second gradient are mutually exclusive");
    FSB_REQUIRE(res.zr || !dres, "bn_act_bwd: dres needs a residual branch");
    if (bn_bwd_compact_ok(dA1, dA2, g, res, dr) && (fmt == FMT_F32 || fmt == FMT_H16))
        return bn_bwd_c8_apply(dA1, dA2, z, a_hi, g, bn, c1, c2, dz, fmt, absmax, s);
    EwShape sh = ew_shape(g);
    if (res.zr)
        bn_act_bwd_apply_kernel<true, false><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, g, bn, res, dr, c1, c2, dz, fmt, dres, dres_bits, absmax);
    else if (dA2.p)
        bn_act_bwd_apply_kernel<false, true><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, g, bn, res, dr, c1, c2, dz, fmt, dres, dres_bits, absmax);
    else
        bn_act_bwd_apply_kernel<false, false><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, g, bn, res, dr, c1, c2, dz, fmt, dres, dres_bits, absmax);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Compact BatchNorm backward (mixed mode, no residual branch): every operand is a half plane, so a thread owns EIGHT
// channels and every access is 16 bytes (the 4-channel kernels above would move 8 bytes per load and, being latency
// bound, lose what the smaller planes save).  Gradients arrive as scaled half planes (GradRef), zhat and the PReLU branch
// come from the hi plane of the stored activation where the inverse map is well conditioned (see BwdCoef above) and from
// the float32 z otherwise -- decided per thread.  Partials / finalize are those of the 4-channel kernels.
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8_scaled(const float (&f)[8], float scale) {
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // clamp to the half range; NaN stays NaN (fminf / fmaxf would drop it)
        const float x0 = f[2 * i] * scale, x1 = f[2 * i + 1] * scale;
        const float lo = x0 == x0 ? fminf(fmaxf(x0, -65504.f), 65504.f) : x0, hi = x1 == x1 ? fminf(fmaxf(x1, -65504.f), 65504.f) : x1;
        const __half2 h = __floats2half2_rn(lo, hi);
        w[i] = *reinterpret_cast<const unsigned*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// may this thread's eight channels work from the stored activation?
__device__ __forceinline__ bool c8_from_a(const BnCoef& bn, const void* a_hi, int c0, int C) {
    bool ok = a_hi != nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (c0 + i >= C) continue;                        // padded channel (the slope vector has C entries only)
        const float sc = bn.scale[c0 + i], is = bn.invstd[c0 + i];
        const float gamma = sc / is, beta = fmaf(bn.mean[c0 + i], sc, bn.shift[c0 + i]);
        const float sl = bn.slope ? bn.slope[c0 + i] : 1.f;
        ok = ok && fabsf(gamma) > 0.f && fabsf(beta) <= 8.f * fabsf(gamma) && fabsf(1.f / gamma) < 3.0e38f &&
             sl >= 0.015625f && sl <= 16.f;
    }
    return ok;
}

// General path (from the float32 pre-activation): y = z * e + f (e = scale, f = shift), zhat = z * b + a (b = invstd,
// a = -mean * invstd).  The fast path (c8_reduce_fast / c8_apply_fast below) folds its coefficients differently.
struct C8Coef {
    float e[8], f[8], sl[8], a[8], b[8];
    bool has_sl;
    __device__ __forceinline__ void load(const BnCoef& bn, int c0, int C) {
        has_sl = bn.slope != nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float is = bn.invstd[c0 + i];
            sl[i] = (has_sl && c0 + i < C) ? bn.slope[c0 + i] : 1.f;     // padded channels: scale = shift = 0, everything stays 0
            e[i] = bn.scale[c0 + i];
            f[i] = bn.shift[c0 + i];
            b[i] = is;
            a[i] = -bn.mean[c0 + i] * is;
        }
    }
};

// raw operands of one row: the source is the stored activation's hi plane (FROM_A, s0 only) or z (two float4)
template <bool DA2, bool FROM_A>
struct C8In {
    uint4 s0, s1;
    uint4 g1, g2;
    __device__ __forceinline__ void load(const GradRef& dA1, const GradRef& dA2, const float* z, const __half* a, long long idx) {
        if (FROM_A) {
            s0 = *reinterpret_cast<const uint4*>(a + idx);
        } else {
            s0 = *reinterpret_cast<const uint4*>(z + idx);
            s1 = *reinterpret_cast<const uint4*>(z + idx + 4);
        }
        g1 = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(dA1.p) + idx);
        if (DA2) g2 = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(dA2.p) + idx);
    }
};

// dy, zhat and the PReLU-slope gradient term of the eight channels (general path)
template <bool DA2>
__device__ __forceinline__ void c8_compute(const C8In<DA2, false>& in, const C8Coef& k, float inv1, float inv2,
                                           float (&dy)[8], float (&zh)[8], float (&dsl)[8]) {
    float g[8];
    const unsigned w[8] = {in.s0.x, in.s0.y, in.s0.z, in.s0.w, in.s1.x, in.s1.y, in.s1.z, in.s1.w};
    unpack8(in.g1, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] *= inv1;
    if (DA2) {
        float g2[8];
        unpack8(in.g2, g2);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(g2[i], inv2, g[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float v = __uint_as_float(w[i]);
        const float y = fmaf(v, k.e[i], k.f[i]);
        const bool pos = y > 0.f || !k.has_sl;
        dy[i] = pos ? g[i] : k.sl[i] * g[i];
        dsl[i] = pos ? 0.f : y * g[i];
        zh[i] = fmaf(v, k.b[i], k.a[i]);
    }
}

struct C8Sums {
    float s0[8], s1[8], s2[8];
    float mx, zx;          // max |dy|, max |zhat| over the thread's eight channels (a common bound for all eight)
};

template <int ROWS, typename In>
__device__ __forceinline__ void c8_load_rows(In (&in)[ROWS], bool (&ok)[ROWS], const GradRef& dA1, const GradRef& dA2,
                                             const float* z, const __half* ah, const Geo& g, long long row0,
                                             long long stride, int c0) {
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
        const long long row = row0 + j * stride;
        ok[j] = row < g.rows && (g.mask == nullptr || g.mask[row]);
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j)
        if (ok[j]) in[j].load(dA1, dA2, z, ah, (row0 + j * stride) * g.Cs + c0);
}

// ---- fast path (from the stored activation): everything folded into per-channel select-and-FMA coefficients.
//   pos = a > 0;  dy = g (pos ? inv1 : sl inv1);  zhat = a (pos ? 1/gamma : 1/(slope gamma)) - beta/gamma;
//   slope term = (1/slope) sum_{!pos} a g inv1   (no activation: sl = 1 and the slope term is dropped)
template <bool DA2>
__device__ __forceinline__ void c8_reduce_fast(const GradRef& dA1, const GradRef& dA2, const __half* ah, const Geo& g,
                                               const BnCoef& bn, int c0, C8Sums& S) {
    const float inv1 = gs_pow2(-gs_exponent2(dA1.bits, dA1.mul));
    const float ratio = DA2 ? gs_pow2(gs_exponent2(dA1.bits, dA1.mul) - gs_exponent2(dA2.bits, dA2.mul)) : 0.f;   // inv2 / inv1
    const bool has_sl = bn.slope != nullptr;
    float slI[8], bp[8], bn_[8], a0[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float sc = bn.scale[c0 + i], sh = bn.shift[c0 + i], mu = bn.mean[c0 + i], is = bn.invstd[c0 + i];
        const bool pad = c0 + i >= g.C;                  // padded channel: a = g = 0, every coefficient 0
        const float sl = (has_sl && !pad) ? bn.slope[c0 + i] : 1.f;
        const float rg = pad ? 0.f : is / sc;
        slI[i] = pad ? 0.f : sl * inv1;
        bp[i] = rg;
        bn_[i] = rg / sl;
        a0[i] = pad ? 0.f : -fmaf(mu, sc, sh) * rg;
    }
    constexpr int ROWS = DA2 ? 2 : 4;
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += ROWS * stride) {
        bool ok[ROWS];
        C8In<DA2, true> in[ROWS];
        c8_load_rows<ROWS>(in, ok, dA1, dA2, nullptr, ah, g, row0, stride, c0);
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) {
                float av[8], gv[8];
                unpack8(in[j].s0, av);
                unpack8(in[j].g1, gv);
                if (DA2) {
                    float g2[8];
                    unpack8(in[j].g2, g2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) gv[i] = fmaf(g2[i], ratio, gv[i]);
                }
                float ady[8], azh[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool pos = av[i] > 0.f;
                    const float dy = gv[i] * (pos ? inv1 : slI[i]);
                    const float zh = fmaf(av[i], pos ? bp[i] : bn_[i], a0[i]);
                    S.s2[i] = fmaf(av[i], pos ? 0.f : gv[i], S.s2[i]);
                    S.s0[i] += dy;
                    S.s1[i] = fmaf(dy, zh, S.s1[i]);
                    ady[i] = fabsf(dy);
                    azh[i] = fabsf(zh);
                }
                // tree maxima: one dependent step per row on the running bounds instead of eight
                const float m0 = fmaxf(fmaxf(ady[0], ady[1]), fmaxf(ady[2], ady[3])), m1 = fmaxf(fmaxf(ady[4], ady[5]), fmaxf(ady[6], ady[7]));
                const float z0 = fmaxf(fmaxf(azh[0], azh[1]), fmaxf(azh[2], azh[3])), z1 = fmaxf(fmaxf(azh[4], azh[5]), fmaxf(azh[6], azh[7]));
                S.mx = fmaxf(S.mx, fmaxf(m0, m1));
                S.zx = fmaxf(S.zx, fmaxf(z0, z1));
            }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) S.s2[i] = (has_sl && c0 + i < g.C) ? S.s2[i] * inv1 / bn.slope[c0 + i] : 0.f;
}

template <bool DA2, bool OUT_F32>
__device__ __forceinline__ void c8_apply_fast(const GradRef& dA1, const GradRef& dA2, const __half* ah, const Geo& g,
                                              const BnCoef& bn, const float* c1, const float* c2, void* dz, float gscale,
                                              int c0) {
    const float inv1 = gs_pow2(-gs_exponent2(dA1.bits, dA1.mul));
    const float ratio = DA2 ? gs_pow2(gs_exponent2(dA1.bits, dA1.mul) - gs_exponent2(dA2.bits, dA2.mul)) : 0.f;
    const bool has_sl = bn.slope != nullptr;
    // dz = scale (dy - c1 - zhat c2) gscale = g (pos ? P1 : P2) + a (pos ? Q1 : Q2) + R
    float P1[8], P2[8], Q1[8], Q2[8], R[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float sc = bn.scale[c0 + i], sh = bn.shift[c0 + i], mu = bn.mean[c0 + i], is = bn.invstd[c0 + i];
        const bool pad = c0 + i >= g.C;                  // padded channel: every coefficient 0, so dz = 0 exactly
        const float sl = (has_sl && !pad) ? bn.slope[c0 + i] : 1.f;
        const float rg = pad ? 0.f : is / sc;
        const float p = pad ? 0.f : sc * gscale, q = pad ? 0.f : -p * c2[c0 + i];
        P1[i] = p * inv1;
        P2[i] = P1[i] * sl;
        Q1[i] = q * rg;
        Q2[i] = Q1[i] / sl;
        R[i] = pad ? 0.f : -p * c1[c0 + i] - q * fmaf(mu, sc, sh) * rg;
    }
    constexpr int ROWS = DA2 ? 2 : 4;
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += ROWS * stride) {
        bool ok[ROWS];
        C8In<DA2, true> in[ROWS];
        c8_load_rows<ROWS>(in, ok, dA1, dA2, nullptr, ah, g, row0, stride, c0);
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) {
                const long long idx = (row0 + j * stride) * g.Cs + c0;
                float av[8], gv[8], o[8];
                unpack8(in[j].s0, av);
                unpack8(in[j].g1, gv);
                if (DA2) {
                    float g2[8];
                    unpack8(in[j].g2, g2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) gv[i] = fmaf(g2[i], ratio, gv[i]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool pos = av[i] > 0.f;
                    o[i] = fmaf(gv[i], pos ? P1[i] : P2[i], fmaf(av[i], pos ? Q1[i] : Q2[i], R[i]));
                }
                if (OUT_F32) {
                    float* d = reinterpret_cast<float*>(dz) + idx;
                    st4(d, make_float4(o[0], o[1], o[2], o[3]));
                    st4(d + 4, make_float4(o[4], o[5], o[6], o[7]));
                } else {
                    // |o| <= the GradScale bound (2^14): no clamp needed
                    unsigned w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __half2 h = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
                        w[i] = *reinterpret_cast<const unsigned*>(&h);
                    }
                    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(dz) + idx) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
    }
}

// ---- general path (from the float32 pre-activation): channel groups whose inverse map is badly conditioned
template <bool DA2>
__device__ __forceinline__ void c8_reduce_general(const GradRef& dA1, const GradRef& dA2, const float* __restrict__ z,
                                                  const Geo& g, const BnCoef& bn, int c0, C8Sums& S) {
    C8Coef k;
    k.load(bn, c0, g.C);
    const float inv1 = gs_pow2(-gs_exponent2(dA1.bits, dA1.mul));
    const float inv2 = DA2 ? gs_pow2(-gs_exponent2(dA2.bits, dA2.mul)) : 1.f;
    constexpr int ROWS = 2;
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += ROWS * stride) {
        bool ok[ROWS];
        C8In<DA2, false> in[ROWS];
        c8_load_rows<ROWS>(in, ok, dA1, dA2, z, nullptr, g, row0, stride, c0);
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) {
                float dy[8], zh[8], dsl[8];
                c8_compute<DA2>(in[j], k, inv1, inv2, dy, zh, dsl);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    S.s0[i] += dy[i];
                    S.s1[i] = fmaf(dy[i], zh[i], S.s1[i]);
                    S.s2[i] += dsl[i];
                    S.mx = fmaxf(S.mx, fabsf(dy[i]));
                    S.zx = fmaxf(S.zx, fabsf(zh[i]));
                }
            }
    }
}

__device__ __forceinline__ void c8_publish(const C8Sums& S, const Geo& g, double* partials) {
    // one-pass block reduction: every thread parks its 26 partials in shared memory (float32), then thread (k, x, i)
    // sums one (record, channel) column over threadIdx.y in double.  Record layout [blockIdx.x][5][Cs] as read by
    // bn_bwd_finalize; the two bounds are common to a thread's eight channels.
    __shared__ float red[26][256];
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[i][t] = S.s0[i]; red[8 + i][t] = S.s1[i]; red[16 + i][t] = S.s2[i]; }
    red[24][t] = S.mx;
    red[25][t] = S.zx;
    __syncthreads();
    const int nthr = blockDim.x * blockDim.y, nout = 5 * 8 * (int)blockDim.x;
    for (int o = t; o < nout; o += nthr) {
        const int k = o / (8 * (int)blockDim.x), rem = o - k * 8 * (int)blockDim.x;
        const int x = rem >> 3, i = rem & 7;
        const int c = (blockIdx.y * blockDim.x + x) * 8 + i;
        if (c >= g.Cs) continue;
        const int slot = k < 3 ? k * 8 + i : 21 + k;
        double acc = 0.0;
        for (int y = 0; y < (int)blockDim.y; ++y) {
            const double v = (double)red[slot][y * blockDim.x + x];
            acc = k < 3 ? acc + v : fmax(acc, v);
        }
        partials[((long long)blockIdx.x * 5 + k) * g.Cs + c] = acc;
    }
}

template <bool DA2>
__global__ void __launch_bounds__(256, 2)
bn_bwd_c8_reduce_kernel(const GradRef dA1, const GradRef dA2, const float* __restrict__ z, const void* a_hi, Geo g,
                        BnCoef bn, double* partials) {
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    const bool cok = cv < g.Cs / 8;
    const int c0 = cv * 8;
    C8Sums S;
#pragma unroll
    for (int i = 0; i < 8; ++i) { S.s0[i] = 0.f; S.s1[i] = 0.f; S.s2[i] = 0.f; }
    S.mx = 0.f; S.zx = 0.f;
    if (cok) {
        if (c8_from_a(bn, a_hi, c0, g.C)) c8_reduce_fast<DA2>(dA1, dA2, reinterpret_cast<const __half*>(a_hi), g, bn, c0, S);
        else c8_reduce_general<DA2>(dA1, dA2, z, g, bn, c0, S);
    }
    c8_publish(S, g, partials);
}

template <bool DA2, bool OUT_F32>
__device__ __forceinline__ void c8_apply_general(const GradRef& dA1, const GradRef& dA2, const float* __restrict__ z,
                                                 const Geo& g, const BnCoef& bn, const float* c1, const float* c2, void* dz,
                                                 float gscale, int c0) {
    C8Coef k;
    k.load(bn, c0, g.C);
    const float inv1 = gs_pow2(-gs_exponent2(dA1.bits, dA1.mul));
    const float inv2 = DA2 ? gs_pow2(-gs_exponent2(dA2.bits, dA2.mul)) : 1.f;
    // dz = scale (dy - c1 - zhat c2) = p dy + q zhat + r   (output GradScale folded in)
    float pq[8], qq[8], rr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float sc = bn.scale[c0 + i] * gscale;
        pq[i] = sc;
        qq[i] = -sc * c2[c0 + i];
        rr[i] = -sc * c1[c0 + i];
    }
    constexpr int ROWS = 2;
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += ROWS * stride) {
        bool ok[ROWS];
        C8In<DA2, false> in[ROWS];
        c8_load_rows<ROWS>(in, ok, dA1, dA2, z, nullptr, g, row0, stride, c0);
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) {
                const long long idx = (row0 + j * stride) * g.Cs + c0;
                float dy[8], zh[8], dsl[8], o[8];
                c8_compute<DA2>(in[j], k, inv1, inv2, dy, zh, dsl);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(pq[i], dy[i], fmaf(qq[i], zh[i], rr[i]));
                if (OUT_F32) {
                    float* d = reinterpret_cast<float*>(dz) + idx;
                    st4(d, make_float4(o[0], o[1], o[2], o[3]));
                    st4(d + 4, make_float4(o[4], o[5], o[6], o[7]));
                } else {
                    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(dz) + idx) = pack8_scaled(o, 1.f);
                }
            }
    }
}

template <bool DA2, bool OUT_F32>
__global__ void __launch_bounds__(256, 2)
bn_bwd_c8_apply_kernel(const GradRef dA1, const GradRef dA2, const float* __restrict__ z, const void* a_hi, Geo g,
                       BnCoef bn, const float* c1, const float* c2, void* dz, const unsigned* absmax) {
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv >= g.Cs / 8) return;
    const int c0 = cv * 8;
    const float gscale = OUT_F32 ? 1.f : gs_scale(absmax);
    if (c8_from_a(bn, a_hi, c0, g.C))
        c8_apply_fast<DA2, OUT_F32>(dA1, dA2, reinterpret_cast<const __half*>(a_hi), g, bn, c1, c2, dz, gscale, c0);
    else
        c8_apply_general<DA2, OUT_F32>(dA1, dA2, z, g, bn, c1, c2, dz, gscale, c0);
}

// ---------------------------------------------------------------------------------------------
// Compact backward of the block output  out = prelu3(bn3(z3) + r0)  (residual branch), eight channels per thread.
// Fast path: the pre-activation y3 and its sign come from the float32 block output `out` (y = out > 0 ? out : out / slope,
// exact to float32) instead of being recomputed from z3 AND zp; zhat3 still comes from the float32 z3 (taking it from
// out - r0 with the half-precision r0 plane was measurably noisier on small networks).  8 bytes per element and pass
// when the incoming gradient is a half plane, 12 for the 4-channel kernel.  Used where 1/64 <= slope3 <= 16; other channel
// groups recompute y3 from z3 and zp.  Writes dz3 and the residual-branch gradient dres = dy as scaled half planes.
struct C8ResIn {
    uint4 g0, g1;       // incoming gradient: half (g0) or float32 (g0, g1)
    uint4 o0, o1;       // z3 (float32)
    uint4 r0, r1;       // fast: out (float32) ; general: zp (float32)
};

template <bool FAST>
__device__ __forceinline__ void c8res_load(C8ResIn& in, const GradRef& dA, const float* z, const float* out,
                                           const float* zr, long long idx) {
    if (dA.half) {
        in.g0 = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(dA.p) + idx);
    } else {
        in.g0 = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dA.p) + idx);
        in.g1 = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dA.p) + idx + 4);
    }
    in.o0 = *reinterpret_cast<const uint4*>(z + idx);
    in.o1 = *reinterpret_cast<const uint4*>(z + idx + 4);
    const float* src = FAST ? out : zr;
    in.r0 = *reinterpret_cast<const uint4*>(src + idx);
    in.r1 = *reinterpret_cast<const uint4*>(src + idx + 4);
}

__device__ __forceinline__ void f8_from(const uint4& a, const uint4& b, float (&f)[8]) {
    const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(w[i]);
}

__device__ __forceinline__ bool c8res_fast_ok(const BnCoef& bn, const void* out, const void* smask, int c0, int C) {
    bool ok = out != nullptr && smask != nullptr && bn.slope != nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (c0 + i >= C) continue;
        const float sl = bn.slope[c0 + i];
        ok = ok && sl >= 0.015625f && sl <= 16.f;
    }
    return ok;
}

// dy (unscaled), zhat and the slope-gradient term of eight channels.  FAST: from out / r0; else from z3 / zp.
template <bool FAST>
struct C8ResCoef {
    float sl[8], e[8], b[8], a[8];           // slope, (FAST: 1/slope ; else scale), invstd, -mean invstd
    float sh[FAST ? 1 : 8], rsc[FAST ? 1 : 8], rsh[FAST ? 1 : 8], rsl[FAST ? 1 : 8];
    bool r_has_sl;
    __device__ __forceinline__ void load(const BnCoef& bn, const Residual& res, int c0, int C) {
        r_has_sl = res.slope != nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool pad = c0 + i >= C;
            const float sc = bn.scale[c0 + i], shv = bn.shift[c0 + i], mu = bn.mean[c0 + i], is = bn.invstd[c0 + i];
            sl[i] = (bn.slope && !pad) ? bn.slope[c0 + i] : 1.f;
            b[i] = is;
            a[i] = -mu * is;
            if (FAST) {
                e[i] = 1.f / sl[i];
            } else {
                e[i] = sc; sh[i] = shv;
                rsc[i] = res.scale[c0 + i]; rsh[i] = res.shift[c0 + i];
                rsl[i] = (r_has_sl && !pad) ? res.slope[c0 + i] : 1.f;
            }
        }
    }
    __device__ __forceinline__ void compute(const C8ResIn& in, bool g_half, float inv, float (&dy)[8], float (&zh)[8],
                                            float (&dsl)[8], unsigned& posbits) const {
        posbits = 0u;
        float g[8], o[8];
        if (g_half) unpack8(in.g0, g);
        else f8_from(in.g0, in.g1, g);
        f8_from(in.o0, in.o1, o);
        if (FAST) {
            float ov[8];
            f8_from(in.r0, in.r1, ov);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool pos = ov[i] > 0.f;
                posbits |= pos ? (1u << i) : 0u;
                const float gi = g[i] * inv;
                const float y = pos ? ov[i] : ov[i] * e[i];
                dy[i] = pos ? gi : sl[i] * gi;
                dsl[i] = pos ? 0.f : y * gi;
                zh[i] = fmaf(o[i], b[i], a[i]);
            }
        } else {
            float zp[8];
            f8_from(in.r0, in.r1, zp);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float rr = fmaf(zp[i], rsc[i], rsh[i]);
                if (r_has_sl) rr = rr > 0.f ? rr : rsl[i] * rr;
                const float y = fmaf(o[i], e[i], sh[i]) + rr;
                const bool pos = y > 0.f;
                const float gi = g[i] * inv;
                dy[i] = pos ? gi : sl[i] * gi;
                dsl[i] = pos ? 0.f : y * gi;
                zh[i] = fmaf(o[i], b[i], a[i]);
            }
        }
    }
};

template <bool FAST>
__device__ __forceinline__ void c8res_reduce_body(const GradRef& dA, const float* z, const float* out, unsigned char* smask,
                                                  const Geo& g, const BnCoef& bn, const Residual& res, int c0, C8Sums& S) {
    C8ResCoef<FAST> k;
    k.load(bn, res, c0, g.C);
    const float inv = dA.half ? gs_pow2(-gs_exponent2(dA.bits, dA.mul)) : 1.f;
    constexpr int ROWS = FAST ? 2 : 1;       // the general path carries twice the coefficients
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += ROWS * stride) {
        bool ok[ROWS];
        C8ResIn in[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const long long row = row0 + j * stride;
            ok[j] = row < g.rows && (g.mask == nullptr || g.mask[row]);
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) c8res_load<FAST>(in[j], dA, z, out, res.zr, (row0 + j * stride) * g.Cs + c0);
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) {
                float dy[8], zh[8], dsl[8];
                unsigned posbits;
                k.compute(in[j], dA.half != 0, inv, dy, zh, dsl, posbits);
                // the apply pass takes the PReLU branch from this byte instead of re-reading the float32 output
                if (FAST) smask[(row0 + j * stride) * (g.Cs >> 3) + (c0 >> 3)] = (unsigned char)posbits;
                float m = 0.f, zm = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    S.s0[i] += dy[i];
                    S.s1[i] = fmaf(dy[i], zh[i], S.s1[i]);
                    S.s2[i] += dsl[i];
                    m = fmaxf(m, fabsf(dy[i]));
                    zm = fmaxf(zm, fabsf(zh[i]));
                }
                S.mx = fmaxf(S.mx, m);
                S.zx = fmaxf(S.zx, zm);
            }
    }
}

__global__ void __launch_bounds__(256, 2)
bn_bwd_c8res_reduce_kernel(const GradRef dA, const float* __restrict__ z, const float* __restrict__ out, unsigned char* smask,
                           Geo g, BnCoef bn, Residual res, double* partials) {
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    const bool cok = cv < g.Cs / 8;
    const int c0 = cv * 8;
    C8Sums S;
#pragma unroll
    for (int i = 0; i < 8; ++i) { S.s0[i] = 0.f; S.s1[i] = 0.f; S.s2[i] = 0.f; }
    S.mx = 0.f; S.zx = 0.f;
    if (cok) {
        if (c8res_fast_ok(bn, out, smask, c0, g.C)) c8res_reduce_body<true>(dA, z, out, smask, g, bn, res, c0, S);
        else c8res_reduce_body<false>(dA, z, out, smask, g, bn, res, c0, S);
    }
    c8_publish(S, g, partials);
}

// general path of the apply pass: y3 recomputed from z3 and zp
__device__ __forceinline__ void c8res_apply_general(const GradRef& dA, const float* z, const Geo& g, const BnCoef& bn,
                                                    const Residual& res, const float* c1, const float* c2, __half* dz,
                                                    float gscale, __half* dres, float dscale, int c0) {
    C8ResCoef<false> k;
    k.load(bn, res, c0, g.C);
    const float inv = dA.half ? gs_pow2(-gs_exponent2(dA.bits, dA.mul)) : 1.f;
    float pq[8], qq[8], rr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool pad = c0 + i >= g.C;
        const float sc = pad ? 0.f : bn.scale[c0 + i] * gscale;
        pq[i] = sc;
        qq[i] = pad ? 0.f : -sc * c2[c0 + i];
        rr[i] = pad ? 0.f : -sc * c1[c0 + i];
    }
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row = (long long)blockIdx.x * blockDim.y + threadIdx.y; row < g.rows; row += stride) {
        if (g.mask != nullptr && !g.mask[row]) continue;
        const long long idx = row * g.Cs + c0;
        C8ResIn in;
        c8res_load<false>(in, dA, z, nullptr, res.zr, idx);
        float dy[8], zh[8], dsl[8];
        unsigned posbits;
        k.compute(in, dA.half != 0, inv, dy, zh, dsl, posbits);
        unsigned wz[4], wr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float o0 = fmaf(pq[2 * i], dy[2 * i], fmaf(qq[2 * i], zh[2 * i], rr[2 * i]));
            const float o1 = fmaf(pq[2 * i + 1], dy[2 * i + 1], fmaf(qq[2 * i + 1], zh[2 * i + 1], rr[2 * i + 1]));
            const __half2 hz = __floats2half2_rn(o0, o1);
            const __half2 hr = __floats2half2_rn(dy[2 * i] * dscale, dy[2 * i + 1] * dscale);
            wz[i] = *reinterpret_cast<const unsigned*>(&hz);
            wr[i] = *reinterpret_cast<const unsigned*>(&hr);
        }
        *reinterpret_cast<uint4*>(dz + idx) = make_uint4(wz[0], wz[1], wz[2], wz[3]);
        if (dres) *reinterpret_cast<uint4*>(dres + idx) = make_uint4(wr[0], wr[1], wr[2], wr[3]);
    }
}

// fast path of the apply pass: the PReLU branch comes from the sign byte the reduce pass left, so only the incoming
// gradient and z3 are read:   dz = g (pos ? P1 : P2) + z Q + R ,  dres = g (pos ? D1 : D2)
__device__ __forceinline__ void c8res_apply_fast(const GradRef& dA, const float* __restrict__ z,
                                                 const unsigned char* __restrict__ smask, const Geo& g, const BnCoef& bn,
                                                 const float* c1, const float* c2, __half* dz, float gscale, __half* dres,
                                                 float dscale, int c0) {
    const float inv = dA.half ? gs_pow2(-gs_exponent2(dA.bits, dA.mul)) : 1.f;
    float P1[8], P2[8], Q[8], R[8], D2[8];
    const float D1 = inv * dscale;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool pad = c0 + i >= g.C;
        const float sl = pad ? 1.f : bn.slope[c0 + i];
        const float sc = pad ? 0.f : bn.scale[c0 + i] * gscale, is = bn.invstd[c0 + i], mu = bn.mean[c0 + i];
        const float q = pad ? 0.f : -sc * c2[c0 + i];
        P1[i] = sc * inv;
        P2[i] = P1[i] * sl;
        Q[i] = q * is;
        R[i] = pad ? 0.f : -sc * c1[c0 + i] - q * mu * is;
        D2[i] = pad ? 0.f : D1 * sl;
    }
    const bool g_half = dA.half != 0;
    constexpr int ROWS = 2;
    const long long stride = (long long)gridDim.x * blockDim.y;
    for (long long row0 = (long long)blockIdx.x * blockDim.y + threadIdx.y; row0 < g.rows; row0 += ROWS * stride) {
        bool ok[ROWS];
        uint4 g0[ROWS], g1[ROWS], z0[ROWS], z1[ROWS];
        unsigned mk[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const long long row = row0 + j * stride;
            ok[j] = row < g.rows && (g.mask == nullptr || g.mask[row]);
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) {
                const long long row = row0 + j * stride, idx = row * g.Cs + c0;
                if (g_half) {
                    g0[j] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(dA.p) + idx);
                } else {
                    g0[j] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dA.p) + idx);
                    g1[j] = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(dA.p) + idx + 4);
                }
                z0[j] = *reinterpret_cast<const uint4*>(z + idx);
                z1[j] = *reinterpret_cast<const uint4*>(z + idx + 4);
                mk[j] = smask[row * (g.Cs >> 3) + (c0 >> 3)];
            }
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            if (ok[j]) {
                const long long idx = (row0 + j * stride) * g.Cs + c0;
                float gv[8], zv[8];
                if (g_half) unpack8(g0[j], gv);
                else f8_from(g0[j], g1[j], gv);
                f8_from(z0[j], z1[j], zv);
                unsigned wz[4], wr[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool p0 = (mk[j] >> (2 * i)) & 1u, p1 = (mk[j] >> (2 * i + 1)) & 1u;
                    const float o0 = fmaf(gv[2 * i], p0 ? P1[2 * i] : P2[2 * i], fmaf(zv[2 * i], Q[2 * i], R[2 * i]));
                    const float o1 = fmaf(gv[2 * i + 1], p1 ? P1[2 * i + 1] : P2[2 * i + 1], fmaf(zv[2 * i + 1], Q[2 * i + 1], R[2 * i + 1]));
                    const __half2 hz = __floats2half2_rn(o0, o1);
                    const __half2 hr = __floats2half2_rn(gv[2 * i] * (p0 ? D1 : D2[2 * i]), gv[2 * i + 1] * (p1 ? D1 : D2[2 * i + 1]));
                    wz[i] = *reinterpret_cast<const unsigned*>(&hz);
                    wr[i] = *reinterpret_cast<const unsigned*>(&hr);
                }
                *reinterpret_cast<uint4*>(dz + idx) = make_uint4(wz[0], wz[1], wz[2], wz[3]);
                if (dres) *reinterpret_cast<uint4*>(dres + idx) = make_uint4(wr[0], wr[1], wr[2], wr[3]);
            }
    }
}

__global__ void __launch_bounds__(256, 2)
bn_bwd_c8res_apply_kernel(const GradRef dA, const float* __restrict__ z, const float* __restrict__ out,
                          const unsigned char* smask, Geo g, BnCoef bn, Residual res, const float* c1, const float* c2,
                          void* dz, const unsigned* absmax, void* dres, const unsigned* dres_bits) {
    const int cv = blockIdx.y * blockDim.x + threadIdx.x;
    if (cv >= g.Cs / 8) return;
    const int c0 = cv * 8;
    const float gscale = gs_scale(absmax), dscale = gs_scale(dres_bits);
    if (c8res_fast_ok(bn, out, smask, c0, g.C))
        c8res_apply_fast(dA, z, smask, g, bn, c1, c2, (__half*)dz, gscale, (__half*)dres, dscale, c0);
    else
        c8res_apply_general(dA, z, g, bn, res, c1, c2, (__half*)dz, gscale, (__half*)dres, dscale, c0);
}

static const int C8_MAX_BLOCKS = 296;   // 2 x 148 SMs: one resident wave (128 registers, 2 CTAs per SM)

static EwShape c8_shape(const Geo& g) {
    const int cv = g.Cs / 8;
    const int bx = cv < 32 ? cv : 32;
    const int by = 256 / bx;
    EwShape s;
    s.block = dim3(bx, by);
    // fewer, longer-lived CTAs than the 4-channel kernels: the block-level reduction tail and the finalize pass scale
    // with the CTA count
    const int gx = ew_shape(g).grid.x;
    s.grid = dim3(gx < C8_MAX_BLOCKS ? gx : C8_MAX_BLOCKS, (cv + bx - 1) / bx);
    return s;
}

int bn_res_bwd_compact_reduce(GradRef dA, const float* z, const float* out, unsigned char* smask, const Geo& g, BnCoef bn,
                              Residual res, double* partials, cudaStream_t s) {
    EW_CHECK(g);
    FSB_REQUIRE(g.Cs % 8 == 0 && res.zr && z, "compact residual BatchNorm backward: bad operands");
    EwShape sh = c8_shape(g);
    bn_bwd_c8res_reduce_kernel<<<sh.grid, sh.block, 0, s>>>(dA, z, out, smask, g, bn, res, partials);
    FSB_LAUNCHED();
    return 0;
}

int bn_res_bwd_compact_blocks(const Geo& g) { return (int)c8_shape(g).grid.x; }

int bn_res_bwd_compact_apply(GradRef dA, const float* z, const float* out, const unsigned char* smask, const Geo& g, BnCoef bn,
                             Residual res, const float* c1, const float* c2, void* dz, const unsigned* absmax, void* dres,
                             const unsigned* dres_bits, cudaStream_t s) {
    EW_CHECK(g);
    FSB_REQUIRE(g.Cs % 8 == 0 && res.zr && z && absmax && (!dres || dres_bits), "compact residual BatchNorm backward: bad operands");
    EwShape sh = c8_shape(g);
    bn_bwd_c8res_apply_kernel<<<sh.grid, sh.block, 0, s>>>(dA, z, out, smask, g, bn, res, c1, c2, dz, absmax, dres, dres_bits);
    FSB_LAUNCHED();
    return 0;
}

bool bn_bwd_compact_ok(const GradRef& dA1, const GradRef& dA2, const Geo& g, const Residual& res, const Dropout& dr) {
    return dA1.half && (!dA2.p || dA2.half) && !res.zr && dr.p == 0.f && g.Cs % 8 == 0;
}

int bn_bwd_num_blocks(GradRef dA1, GradRef dA2, const Geo& g, Residual res, Dropout dr) {
    return bn_bwd_compact_ok(dA1, dA2, g, res, dr) ? (int)c8_shape(g).grid.x : ew_num_blocks(g);
}

static int bn_bwd_c8_reduce(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn,
                            double* partials, cudaStream_t s) {
    EwShape sh = c8_shape(g);
    if (dA2.p) bn_bwd_c8_reduce_kernel<true><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, a_hi, g, bn, partials);
    else bn_bwd_c8_reduce_kernel<false><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, a_hi, g, bn, partials);
    FSB_LAUNCHED();
    return 0;
}

static int bn_bwd_c8_apply(GradRef dA1, GradRef dA2, const float* z, const void* a_hi, const Geo& g, BnCoef bn,
                           const float* c1, const float* c2, void* dz, int fmt, const unsigned* absmax, cudaStream_t s) {
    FSB_REQUIRE(fmt == FMT_F32 || fmt == FMT_H16, "compact BatchNorm backward writes float32 or one half plane");
    EwShape sh = c8_shape(g);
    if (fmt == FMT_F32) {
        if (dA2.p) bn_bwd_c8_apply_kernel<true, true><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, a_hi, g, bn, c1, c2, dz, absmax);
        else bn_bwd_c8_apply_kernel<false, true><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, a_hi, g, bn, c1, c2, dz, absmax);
    } else {
        if (dA2.p) bn_bwd_c8_apply_kernel<true, false><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, a_hi, g, bn, c1, c2, dz, absmax);
        else bn_bwd_c8_apply_kernel<false, false><<<sh.grid, sh.block, 0, s>>>(dA1, dA2, z, a_hi, g, bn, c1, c2, dz, absmax);
    }
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const float* x, long long rows, int C, int ld, float* out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (long long r = 0; r < rows; ++r) s += x[r * ld + c];
    out[c] = (float)s;
}

int colsum(const float* x, long long rows, int C, int ld, float* out, cudaStream_t s) {
    colsum_kernel<<<(C + 127) / 128, 128, 0, s>>>(x, rows, C, ld, out);
    FSB_LAUNCHED();
    return 0;
}

__global__ void copy2d_kernel(const float* src, long long rows, int C, int lds, float* dst, int ldd) {
    long long total = rows * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i / C;
        int c = (int)(i % C);
        dst[r * ldd + c] = src[r * lds + c];
    }
}

int copy2d(const float* src, long long rows, int C, int lds, float* dst, int ldd, cudaStream_t s) {
    long long total = rows * C;
    int blocks = (int)((total + 255) / 256 > 2048 ? 2048 : (total + 255) / 256);
    if (blocks < 1) blocks = 1;
    copy2d_kernel<<<blocks, 256, 0, s>>>(src, rows, C, lds, dst, ldd);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void nchw_to_pf_kernel(const float* x, Geo g, void* dst, int fmt) {
    long long total = g.pixels * g.Cs;
    long long plane = g.rows * g.Cs;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % g.Cs);
        long long q = i / g.Cs;
        int xx = (int)(q % g.W);
        long long t = q / g.W;
        int yy = (int)(t % g.H);
        int n = (int)(t / g.H);
        float v = c < g.C ? x[(((long long)n * g.C + c) * g.H + yy) * g.W + xx] : 0.f;
        long long idx = geo_row(g, n, yy, xx) * g.Cs + c;
        if (fmt == FMT_F32) {
            reinterpret_cast<float*>(dst)[idx] = v;
        } else {
            __half h, l;
            split_h16(v, h, l);
            reinterpret_cast<__half*>(dst)[idx] = h;
            if (fmt == FMT_H16X2) reinterpret_cast<__half*>(dst)[plane + idx] = l;
        }
    }
}

int nchw_to_pf(const float* x, const Geo& g, void* dst, int fmt, cudaStream_t s) {
    nchw_to_pf_kernel<<<1184, 256, 0, s>>>(x, g, dst, fmt);
    FSB_LAUNCHED();
    return 0;
}

__global__ void pf_to_nchw_kernel(const float* src, Geo g, float* dst) {
    long long total = g.pixels * g.C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int xx = (int)(i % g.W);
        long long t = i / g.W;
        int yy = (int)(t % g.H);
        t /= g.H;
        int c = (int)(t % g.C);
        int n = (int)(t / g.C);
        dst[i] = src[geo_row(g, n, yy, xx) * g.Cs + c];
    }
}

int pf_to_nchw(const float* src, const Geo& g, float* dst, cudaStream_t s) {
    pf_to_nchw_kernel<<<1184, 256, 0, s>>>(src, g, dst);
    FSB_LAUNCHED();
    return 0;
}

}  // namespace fsb
