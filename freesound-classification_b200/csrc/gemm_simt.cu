// float32 CUDA-core back end of the row-shifted GEMM (precision 0).  It is the bit-faithful
// float32 path (used for the FC head in every precision mode, and as the on-device cross-check of
// the tcgen05 back end at sizes the CPU oracle cannot reach).
#include "gemm.cuh"
#include "eltwise.cuh"

namespace fsb {

// packed record: [fwd: ntaps][CsIn][CsOut] | [dgrad: ntaps][CsOut][CsIn] | [bias: CsOut]
size_t simt_packed_weight_bytes(const ConvGeom& c) {
    return ((size_t)2 * c.ntaps * c.CsIn * c.CsOut + c.CsOut) * sizeof(float);
}

__global__ void simt_pack_kernel(const float* w, const float* bias, ConvGeom c, float* fwd, float* dgr, float* pb) {
    long long total = (long long)c.ntaps * c.CsIn * c.CsOut;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int co = (int)(i % c.CsOut);
        long long t2 = i / c.CsOut;
        int ci = (int)(t2 % c.CsIn);
        int t = (int)(t2 / c.CsIn);
        float v = (co < c.Cout && ci < c.Cin) ? w[((long long)co * c.Cin + ci) * c.ntaps + t] : 0.f;
        fwd[i] = v;
        // dgrad consumes taps in the same index order but with negated offsets
        dgr[((long long)t * c.CsOut + co) * c.CsIn + ci] = v;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.CsOut; i += gridDim.x * blockDim.x)
        pb[i] = (bias && i < c.Cout) ? bias[i] : 0.f;
}

int simt_pack_weights(const float* w, const float* bias, const ConvGeom& c, void* packed, cudaStream_t s) {
    float* fwd = (float*)packed;
    float* dgr = fwd + (size_t)c.ntaps * c.CsIn * c.CsOut;
    float* pb = dgr + (size_t)c.ntaps * c.CsIn * c.CsOut;
    long long total = (long long)c.ntaps * c.CsIn * c.CsOut;
    int blocks = (int)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
    simt_pack_kernel<<<blocks, 256, 0, s>>>(w, bias, c, fwd, dgr, pb);
    FSB_LAUNCHED();
    return 0;
}

struct Taps {
    int n;
    int off[9];
};

// Z[r, n] = bias[n] + sum_t sum_k A[r + off_t, k] * W[t][k][n];   128 x 64 tile, 8 x 4 per thread
__global__ void __launch_bounds__(256)
simt_gemm_kernel(const float* __restrict__ A, long long rows, int K, Taps taps, const float* __restrict__ W,
                 const float* __restrict__ bias, float* __restrict__ Z, int Nn, int k_per_split) {
    __shared__ __align__(16) float As[16][132];
    __shared__ __align__(16) float Bs[16][64];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long r0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * 64;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < taps.n; ++t) {
        const long long off = taps.off[t];
        const float* Wt = W + (long long)t * K * Nn;
        // split-K (skinny problems): slice blockIdx.z of the reduction goes to plane blockIdx.z of Z
        const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
        for (int k0 = k_begin; k0 < k_end; k0 += 16) {
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                int idx = tid + l * 256;
                int row = idx >> 2, kv = idx & 3;
                long long r = r0 + row + off;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r >= 0 && r < rows) v = *reinterpret_cast<const float4*>(A + r * K + k0 + kv * 4);
                As[kv * 4 + 0][row] = v.x; As[kv * 4 + 1][row] = v.y;
                As[kv * 4 + 2][row] = v.z; As[kv * 4 + 3][row] = v.w;
            }
            {
                int kk = tid >> 4, nv = tid & 15;
                int n = n0 + nv * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < Nn) v = *reinterpret_cast<const float4*>(Wt + (long long)(k0 + kk) * Nn + n);
                *reinterpret_cast<float4*>(&Bs[kk][nv * 4]) = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                float a[8];
                float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
                float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                    acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                    acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                    acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
                }
            }
            __syncthreads();
        }
    }
    const int n = n0 + tx * 4;
    Z += (long long)blockIdx.z * rows * Nn;
    if (n < Nn) {
        float4 bv = bias ? *reinterpret_cast<const float4*>(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            long long r = r0 + ty * 8 + i;
            if (r < rows)
                *reinterpret_cast<float4*>(Z + r * Nn + n) =
                    make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        }
    }
}

int simt_fwd(const float* A, const void* packed, float* Z, const ConvGeom& c, cudaStream_t s) {
    const float* fwd = (const float*)packed;
    const float* pb = fwd + (size_t)2 * c.ntaps * c.CsIn * c.CsOut;
    Taps t;
    t.n = c.ntaps;
    for (int i = 0; i < 9; ++i) t.off[i] = c.offs[i];
    dim3 grid((unsigned)((c.rows + 127) / 128), (c.CsOut + 63) / 64);
    simt_gemm_kernel<<<grid, 256, 0, s>>>(A, c.rows, c.CsIn, t, fwd, pb, Z, c.CsOut, c.CsIn);
    FSB_LAUNCHED();
    return 0;
}

int simt_dgrad(const float* dZ, const void* packed, float* dA, const ConvGeom& c, cudaStream_t s) {
    const float* dgr = (const float*)packed + (size_t)c.ntaps * c.CsIn * c.CsOut;
    Taps t;
    t.n = c.ntaps;
    for (int i = 0; i < 9; ++i) t.off[i] = -c.offs[i];
    dim3 grid((unsigned)((c.rows + 127) / 128), (c.CsIn + 63) / 64);
    simt_gemm_kernel<<<grid, 256, 0, s>>>(dZ, c.rows, c.CsOut, t, dgr, nullptr, dA, c.CsIn, c.CsOut);
    FSB_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// skinny (few rows, e.g. the FC head: rows = batch) linear layers: split-K over gridDim.z, partial planes in
// scratch, summed in split order together with the bias (deterministic)
static const int SKINNY_SPLITS = 16;

size_t simt_skinny_scratch_bytes(const ConvGeom& c) {
    int cs = c.CsIn > c.CsOut ? c.CsIn : c.CsOut;
    return (size_t)SKINNY_SPLITS * c.rows * cs * sizeof(float);
}

__global__ void __launch_bounds__(256)
skinny_finalize_kernel(const float* __restrict__ P, int splits, long long plane, int Nn, const float* __restrict__ bias,
                       float* __restrict__ Z) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) {
        float s = bias ? bias[i % Nn] : 0.f;
#pragma unroll 4
        for (int sp = 0; sp < splits; ++sp) s += P[sp * plane + i];
        Z[i] = s;
    }
}

static int skinny_gemm(const float* A, long long rows, int K, const float* W, const float* bias, float* Z, int Nn,
                       float* scratch, cudaStream_t s) {
    FSB_REQUIRE(K % 16 == 0, "skinny_gemm: K must be a multiple of 16");
    int k_per_split = ((K / 16 + SKINNY_SPLITS - 1) / SKINNY_SPLITS) * 16;
    int splits = (K + k_per_split - 1) / k_per_split;
    Taps t;
    t.n = 1;
    for (int i = 0; i < 9; ++i) t.off[i] = 0;
    dim3 grid((unsigned)((rows + 127) / 128), (Nn + 63) / 64, splits);
    simt_gemm_kernel<<<grid, 256, 0, s>>>(A, rows, K, t, W, nullptr, scratch, Nn, k_per_split);
    FSB_LAUNCHED();
    long long plane = rows * Nn;
    int blocks = (int)((plane + 255) / 256 > 1184 ? 1184 : (plane + 255) / 256);
    skinny_finalize_kernel<<<blocks, 256, 0, s>>>(scratch, splits, plane, Nn, bias, Z);
    FSB_LAUNCHED();
    return 0;
}

int simt_skinny_fwd(const float* A, const void* packed, float* Z, const ConvGeom& c, void* scratch, cudaStream_t s) {
    FSB_REQUIRE(c.ntaps == 1, "skinny GEMM: linear layers only");
    const float* fwd = (const float*)packed;
    const float* pb = fwd + (size_t)2 * c.CsIn * c.CsOut;
    return skinny_gemm(A, c.rows, c.CsIn, fwd, pb, Z, c.CsOut, (float*)scratch, s);
}

int simt_skinny_dgrad(const float* dZ, const void* packed, float* dA, const ConvGeom& c, void* scratch, cudaStream_t s) {
    FSB_REQUIRE(c.ntaps == 1, "skinny GEMM: linear layers only");
    const float* dgr = (const float*)packed + (size_t)c.CsIn * c.CsOut;
    return skinny_gemm(dZ, c.rows, c.CsOut, dgr, nullptr, dA, c.CsIn, (float*)scratch, s);
}

// ---------------------------------------------------------------------------------------------
// wgrad: P[split][t][ci][co] = sum over the split's rows of A[r+off_t, ci] * dZ[r, co]
static int wgrad_splits(const ConvGeom& c) {
    int tiles = ((c.CsIn + 63) / 64) * ((c.CsOut + 63) / 64) * c.ntaps;
    int want = (1184 + tiles - 1) / tiles;
    long long chunks = (c.rows + 15) / 16;
    if (want > chunks) want = (int)chunks;
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    return want;
}

size_t simt_wgrad_scratch_bytes(const ConvGeom& c) {
    return (size_t)wgrad_splits(c) * c.ntaps * c.CsIn * c.CsOut * sizeof(float);
}

__global__ void __launch_bounds__(256)
simt_wgrad_kernel(const float* __restrict__ A, const float* __restrict__ dZ, long long rows, int CsIn, int CsOut,
                  Taps taps, int splits, long long rows_per_split, float* __restrict__ P) {
    __shared__ __align__(16) float As[16][64];
    __shared__ __align__(16) float Bs[16][64];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64;
    const int t = blockIdx.z / splits, sp = blockIdx.z % splits;
    const long long off = taps.off[t];
    const long long rb = sp * rows_per_split;
    const long long re = rb + rows_per_split < rows ? rb + rows_per_split : rows;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 4, lv = tid & 15;
    for (long long r0 = rb; r0 < re; r0 += 16) {
        long long r = r0 + lr;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (r < re) {
            long long ra = r + off;
            if (ra >= 0 && ra < rows && ci0 + lv * 4 < CsIn)
                a = *reinterpret_cast<const float4*>(A + ra * CsIn + ci0 + lv * 4);
            if (co0 + lv * 4 < CsOut) b = *reinterpret_cast<const float4*>(dZ + r * CsOut + co0 + lv * 4);
        }
        *reinterpret_cast<float4*>(&As[lr][lv * 4]) = a;
        *reinterpret_cast<float4*>(&Bs[lr][lv * 4]) = b;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(aa[i], bv.x, acc[i][0]);
                acc[i][1] = fmaf(aa[i], bv.y, acc[i][1]);
                acc[i][2] = fmaf(aa[i], bv.z, acc[i][2]);
                acc[i][3] = fmaf(aa[i], bv.w, acc[i][3]);
            }
        }
        __syncthreads();
    }
    float* Pt = P + ((long long)sp * taps.n + t) * CsIn * CsOut;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int ci = ci0 + ty * 4 + i, co = co0 + tx * 4;
        if (ci < CsIn && co < CsOut)
            *reinterpret_cast<float4*>(Pt + (long long)ci * CsOut + co) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// dw[co][ci][t] = sum_split P[split][t][ci][co]   (fixed order: deterministic).  Threads walk P in its own
// order (co fastest) so every split plane is read coalesced; the `splits` loads of a thread are independent.
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ P, int splits, ConvGeom c, float* dw,
                                                                  const unsigned* dz_absmax) {
    const long long plane = (long long)c.ntaps * c.CsIn * c.CsOut;
    const float unscale = gs_inv_scale(dz_absmax);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plane;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % c.CsOut);
        const long long u = i / c.CsOut;
        const int ci = (int)(u % c.CsIn);
        const int t = (int)(u / c.CsIn);
        if (co >= c.Cout || ci >= c.Cin) continue;
        // up to 128 partial planes per element and only ~10^4 elements in the 1x1 layers: the loop is latency bound, so
        // 32 loads are kept in flight on four independent chains (fixed order: deterministic)
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int sp = 0;
        for (; sp + 32 <= splits; sp += 32) {
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] = P[(long long)(sp + u) * plane + i];
#pragma unroll
            for (int u = 0; u < 32; u += 4) { s0 += v[u]; s1 += v[u + 1]; s2 += v[u + 2]; s3 += v[u + 3]; }
        }
#pragma unroll 4
        for (; sp < splits; ++sp) s0 += P[(long long)sp * plane + i];
        dw[((long long)co * c.Cin + ci) * c.ntaps + t] = ((s0 + s1) + (s2 + s3)) * unscale;
    }
}

int wgrad_finalize(const float* P, int splits, const ConvGeom& c, float* dw, const unsigned* dz_absmax, cudaStream_t s) {
    long long total = (long long)c.ntaps * c.CsIn * c.CsOut;
    int blocks = (int)((total + 255) / 256 > 2368 ? 2368 : (total + 255) / 256);
    wgrad_finalize_kernel<<<blocks, 256, 0, s>>>(P, splits, c, dw, dz_absmax);
    FSB_LAUNCHED();
    return 0;
}

int simt_wgrad(const float* A, const float* dZ, float* dw, void* scratch, const ConvGeom& c, cudaStream_t s) {
    int splits = wgrad_splits(c);
    long long chunks = (c.rows + 15) / 16;
    long long rows_per_split = (chunks + splits - 1) / splits * 16;
    Taps t;
    t.n = c.ntaps;
    for (int i = 0; i < 9; ++i) t.off[i] = c.offs[i];
    dim3 grid((c.CsIn + 63) / 64, (c.CsOut + 63) / 64, c.ntaps * splits);
    simt_wgrad_kernel<<<grid, 256, 0, s>>>(A, dZ, c.rows, c.CsIn, c.CsOut, t, splits, rows_per_split, (float*)scratch);
    FSB_LAUNCHED();
    return wgrad_finalize((const float*)scratch, splits, c, dw, nullptr, s);
}

// ---------------------------------------------------------------------------------------------
// precision dispatch
size_t packed_weight_bytes(int precision, const ConvGeom& c) {
    return precision == 0 ? simt_packed_weight_bytes(c) : tc_packed_weight_bytes(c);
}
int pack_weights(int precision, const float* w, const float* bias, const ConvGeom& c, void* packed, cudaStream_t s) {
    return precision == 0 ? simt_pack_weights(w, bias, c, packed, s) : tc_pack_weights(w, bias, c, packed, s);
}
int conv_gemm_fwd(int precision, const void* A, const void* packed, float* Z, const ConvGeom& c, const FwdStats* st,
                  cudaStream_t s) {
    if (precision != 0) return tc_fwd(precision, A, packed, Z, c, st, s);
    FSB_TRY(simt_fwd((const float*)A, packed, Z, c, s));
    if (st) {
        FSB_TRY(pf_stats(Z, *st->g, st->partials, s));
        *st->nblk = ew_num_blocks(*st->g);
    }
    return 0;
}
int conv_gemm_fwd_act(int precision, const void* A, const void* packed, void* a_out, const ConvGeom& c, const FwdAct& act,
                      cudaStream_t s) {
    FSB_REQUIRE(precision != 0, "conv_gemm_fwd_act: tensor-core back ends only");
    return tc_fwd_act(precision, A, packed, a_out, c, act, s);
}
int conv_gemm_dgrad(int precision, const void* dZ, const void* packed, void* dA, const ConvGeom& c,
                    const unsigned* dz_absmax, const float* out_half_mul, cudaStream_t s) {
    FSB_REQUIRE(precision != 0 || !out_half_mul, "conv_gemm_dgrad: the float32 back end writes float32 only");
    return precision == 0 ? simt_dgrad((const float*)dZ, packed, (float*)dA, c, s)
                          : tc_dgrad(precision, dZ, packed, dA, c, dz_absmax, out_half_mul, s);
}
size_t wgrad_scratch_bytes(int precision, const ConvGeom& c) {
    return precision == 0 ? simt_wgrad_scratch_bytes(c) : tc_wgrad_scratch_bytes(c);
}
int conv_gemm_wgrad(int precision, const void* A, const void* dZ, float* dw, void* scratch, const ConvGeom& c,
                    const unsigned* dz_absmax, cudaStream_t s) {
    return precision == 0 ? simt_wgrad((const float*)A, (const float*)dZ, dw, scratch, c, s)
                          : tc_wgrad(precision, A, dZ, dw, scratch, c, dz_absmax, s);
}

// ---------------------------------------------------------------------------------------------
// max column L1 norm of conv weights: the Hoelder bound |dA[r, ci]| <= max|dZ| * sum_{co,t} |w[co][ci][t]| that fixes
// the GradScale of a half-precision dgrad output before the GEMM runs.  One CTA per layer, one thread per input
// channel (strided), block-wide max.
struct WeightL1Jobs {
    WeightL1Job j[32];
};

// grid (input-channel chunks of 32, layers); block (32 input channels, 8 output-channel slices)
__global__ void __launch_bounds__(256) weight_l1_kernel(const WeightL1Jobs jobs, unsigned* out_bits) {
    const WeightL1Job job = jobs.j[blockIdx.y];
    const int ci = blockIdx.x * 32 + threadIdx.x;
    float sum = 0.f;
    if (ci < job.Cin) {
        for (int co = threadIdx.y; co < job.Cout; co += 8) {
            const float* w = job.w + ((long long)co * job.Cin + ci) * job.ntaps;
            for (int t = 0; t < job.ntaps; ++t) sum += fabsf(w[t]);
        }
    }
    __shared__ float red[8][32];
    red[threadIdx.y][threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.y == 0) {
        float tot = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) tot += red[y][threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot = fmaxf(tot, __shfl_xor_sync(0xffffffffu, tot, o));
        // 1.001: the GEMM multiplies by the half-rounded weights.  Non-negative floats order like their bit patterns.
        if (threadIdx.x == 0 && blockIdx.x * 32 < job.Cin) atomicMax(out_bits + blockIdx.y, __float_as_uint(tot * 1.001f));
    }
}

int weight_l1_bounds(const WeightL1Job* jobs, int njobs, float* out, cudaStream_t s) {
    FSB_REQUIRE(njobs >= 1 && njobs <= 32, "weight_l1_bounds: 1..32 layers per launch");
    WeightL1Jobs J;
    memset(&J, 0, sizeof(J));
    int max_cin = 1;
    for (int i = 0; i < njobs; ++i) {
        J.j[i] = jobs[i];
        if (jobs[i].Cin > max_cin) max_cin = jobs[i].Cin;
    }
    FSB_CUDA(cudaMemsetAsync(out, 0, (size_t)njobs * sizeof(float), s));
    weight_l1_kernel<<<dim3((max_cin + 31) / 32, njobs), dim3(32, 8), 0, s>>>(J, reinterpret_cast<unsigned*>(out));
    FSB_LAUNCHED();
    return 0;
}

}  // namespace fsb
