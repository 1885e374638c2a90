// Library plumbing (errors, version, launch counter) + the small stand-alone kernels:
// LSEP loss forward/backward, multi-tensor Adam-amsgrad, on-device MixUp.
#include <stdarg.h>

#include "common.cuh"

namespace fsb {

static thread_local char g_err[1024] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// LSEP: one warp per sample.  L_n = log(1 + sum_{i,j : t_j < t_i} exp(s_j - s_i))
// (networks/losses.py:47-58; pairwise form kept so non-binary targets behave like the reference).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// STABLE: the reference's `lsep_loss_stable` (networks/losses.py:25-44) -- shift by m = max over ALL pairs of
// (s_j - s_i) = max s - min s, L = m + log(e^-m + sum e^(d - m)): finite where the plain form overflows.
template <bool STABLE>
__device__ __forceinline__ float lsep_shift(const float* sr, int c, int lane) {
    if (!STABLE) return 0.f;
    float hi = -INFINITY, lo = INFINITY;
    for (int j = lane; j < c; j += 32) { hi = fmaxf(hi, sr[j]); lo = fminf(lo, sr[j]); }
    return warp_max(hi) + warp_max(-lo);
}

template <bool STABLE>
__global__ void lsep_fwd_kernel(const float* __restrict__ s, const float* __restrict__ t, int n, int c, float* loss) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const float* sr = s + (long long)warp * c;
    const float* tr = t + (long long)warp * c;
    const float m = lsep_shift<STABLE>(sr, c, lane);
    float acc = 0.f;
    for (int i = 0; i < c; ++i) {
        float si = sr[i], ti = tr[i];
        for (int j = lane; j < c; j += 32)
            if (tr[j] < ti) acc += expf(sr[j] - si - m);
    }
    acc = warp_sum(acc);
    if (lane == 0) loss[warp] = STABLE ? m + logf(expf(-m) + acc) : logf(1.0f + acc);
}

// dL/ds_k = ( sum_{i: t_k < t_i} e^{s_k - s_i}  -  sum_{j: t_j < t_k} e^{s_j - s_k} ) / (1 + S)
// (STABLE: numerator and denominator both carry the factor e^-m)
template <bool STABLE>
__global__ void lsep_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t, const float* __restrict__ dloss,
                                int n, int c, float* ds) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const float* sr = s + (long long)warp * c;
    const float* tr = t + (long long)warp * c;
    const float m = lsep_shift<STABLE>(sr, c, lane);
    float total = 0.f;
    for (int k = lane; k < c; k += 32) {
        float sk = sr[k], tk = tr[k];
        for (int i = 0; i < c; ++i)
            if (tk < tr[i]) total += expf(sk - sr[i] - m);
    }
    total = warp_sum(total);
    float inv = dloss[warp] / ((STABLE ? expf(-m) : 1.0f) + total);
    for (int k = lane; k < c; k += 32) {
        float sk = sr[k], tk = tr[k];
        float g = 0.f;
        for (int i = 0; i < c; ++i) {
            float ti = tr[i];
            if (tk < ti) g += expf(sk - sr[i] - m);
            else if (ti < tk) g -= expf(sr[i] - sk - m);
        }
        ds[(long long)warp * c + k] = g * inv;
    }
}

// ---------------------------------------------------------------------------------------------
// lwlrap (ops/utils.py:17-26): sklearn's label_ranking_average_precision_score(truth > 0, scores,
// sample_weight = #positives) over the rows with >= 1 positive.  Per row: for every relevant label j,
// rank_j = #{k : score_k >= score_j} (ties share the worst rank, sklearn's rankdata(.., "max")), L_j the same count over
// relevant k; the row scores mean_j(L_j / rank_j), or exactly 1 when every label is relevant.  With weight = #positives
// the weighted mean is  sum_rows sum_j L_j / rank_j  /  sum_rows #positives: one warp per row writes its numerator and
// weight (doubles), a fixed-order pass adds them (deterministic).
// ---------------------------------------------------------------------------------------------
__global__ void lwlrap_rows_kernel(const float* __restrict__ truth, const float* __restrict__ scores, int n, int c,
                                   double* row_num, double* row_den) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const float* tr = truth + (long long)warp * c;
    const float* sr = scores + (long long)warp * c;
    double num = 0.0;
    int npos = 0;
    for (int j = lane; j < c; j += 32) {
        if (tr[j] > 0.f) {
            const float sj = sr[j];
            int rank = 0, rel = 0;
            for (int k = 0; k < c; ++k) {
                const bool ge = sr[k] >= sj;
                rank += ge;
                rel += ge && tr[k] > 0.f;
            }
            num += (double)rel / (double)rank;
            ++npos;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_xor_sync(0xffffffffu, num, o);
        npos += __shfl_xor_sync(0xffffffffu, npos, o);
    }
    if (lane == 0) {
        if (npos == c) num = (double)npos;           // every label relevant: sklearn scores the row 1
        row_num[warp] = num;
        row_den[warp] = (double)npos;
    }
}

__global__ void lwlrap_sum_kernel(const double* row_num, const double* row_den, int n, double* out, int accumulate) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = accumulate ? out[0] : 0.0, b = accumulate ? out[1] : 0.0;
        for (int i = 0; i < n; ++i) { a += row_num[i]; b += row_den[i]; }
        out[0] = a;
        out[1] = b;
        out[2] = b > 0.0 ? a / b : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Batch assembly over a device-resident PCM pool: SampleLongAudio crop (ops/transforms.py:292-309)
// -> MixUp (ops/transforms.py:44-65, ops/audio.py:32-52) -> zero-pad collate (ops/padding.py:8-32) in one pass.
// All random draws (coin, partner, alpha, offsets, crop starts) are made on the host in the reference's order and arrive
// as per-row records; the device does the PCM traffic.
//   equal lengths   : out = (a + b) / 2
//   unequal lengths : out = alpha * longer, then out[offset : offset + len(shorter)] = (1 - alpha) * shorter  (the
//                     reference's `=+` ASSIGNS; quirk kept)
//   labels          : clip(l1 + l2, 0, 1)
// ---------------------------------------------------------------------------------------------
struct AssembleRow {
    long long a_off;      // first sample of the (cropped) primary clip in the pool
    long long b_off;      // first sample of the (cropped) partner clip, unused when b_len < 0
    double alpha;         // unequal branch: scale of the longer clip
    double one_minus;     // unequal branch: scale of the shorter clip (1 - alpha, evaluated on the host)
    int a_len, b_len;     // lengths after cropping; b_len < 0: no MixUp for this row
    int a_label, b_label; // rows of the label pool
    int mix_offset;       // unequal branch: where the shorter clip lands inside the longer one
    int pad_;
};

__global__ void __launch_bounds__(256)
assemble_kernel(const float* __restrict__ pool, const float* __restrict__ label_pool, const AssembleRow* __restrict__ rows,
                int c, long long t_out, float pad_value, float* __restrict__ out, float* __restrict__ labels_out) {
    const AssembleRow r = rows[blockIdx.y];
    float* o = out + (long long)blockIdx.y * t_out;
    const float* a = pool + r.a_off;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long k0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r.b_len < 0) {
        for (long long k = k0; k < t_out; k += stride) o[k] = k < r.a_len ? a[k] : pad_value;
    } else {
        const float* b = pool + r.b_off;
        if (r.a_len == r.b_len) {
            for (long long k = k0; k < t_out; k += stride) o[k] = k < r.a_len ? (a[k] + b[k]) / 2.0f : pad_value;
        } else {
            const bool a_long = r.a_len > r.b_len;
            const float* lg = a_long ? a : b;
            const float* sh = a_long ? b : a;
            const int n_long = a_long ? r.a_len : r.b_len, n_short = a_long ? r.b_len : r.a_len;
            // float32 arithmetic with the scalars rounded to float32: what numpy evaluates for the reference's
            // `longer *= a` / `shorter * (1 - a)` (a is a Python float: float32 array op, under value-based casting and
            // under NEP 50 alike); 1 - a is formed in float64 first
            const float alpha = (float)r.alpha, one_minus = (float)r.one_minus;
            for (long long k = k0; k < t_out; k += stride) {
                float v = pad_value;
                if (k < n_long) {
                    const long long j = k - r.mix_offset;
                    v = (j >= 0 && j < n_short) ? sh[j] * one_minus : lg[k] * alpha;
                }
                o[k] = v;
            }
        }
    }
    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < c; k += blockDim.x) {
            float l = label_pool[(long long)r.a_label * c + k];
            if (r.b_len >= 0) l = fminf(fmaxf(l + label_pool[(long long)r.b_label * c + k], 0.f), 1.f);
            labels_out[(long long)blockIdx.y * c + k] = l;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Adam(amsgrad) multi-tensor
// ---------------------------------------------------------------------------------------------
static constexpr int ADAM_CHUNK = 8192;

struct AdamRec {
    float* p;
    float* g;
    float* m;
    float* v;
    float* vmax;
    long long n;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float& vm, float step_size, float sqrt_bc2,
                                            float beta1, float beta2, float eps, float wd, float gscale) {
    g *= gscale;
    if (wd != 0.f) g = fmaf(wd, p, g);
    m = m + (g - m) * (1.0f - beta1);                 // torch: exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + g * g * (1.0f - beta2);
    vm = fmaxf(vm, v);
    const float denom = sqrtf(vm) / sqrt_bc2 + eps;
    p = p - step_size * (m / denom);
}

// one CTA per (tensor record, 8192-element chunk); float4 accesses when the record's five arrays are 16-byte aligned
__global__ void __launch_bounds__(256)
adam_kernel(const AdamRec* __restrict__ table, const int* __restrict__ block_map, float step_size, float sqrt_bc2,
            float beta1, float beta2, float eps, float wd, float gscale) {
    const AdamRec r = table[block_map[2 * blockIdx.x]];
    const long long base = (long long)block_map[2 * blockIdx.x + 1] * ADAM_CHUNK;
    const long long end = base + ADAM_CHUNK < r.n ? base + ADAM_CHUNK : r.n;
    const bool aligned = ((((size_t)r.p | (size_t)r.g | (size_t)r.m | (size_t)r.v | (size_t)r.vmax) & 15) == 0);
    long long i = base + threadIdx.x;
    if (aligned) {
        const long long end4 = base + ((end - base) & ~3ll);
        for (long long j = base + 4ll * threadIdx.x; j < end4; j += 4 * 256) {
            float4 p = *reinterpret_cast<float4*>(r.p + j);
            const float4 g = *reinterpret_cast<const float4*>(r.g + j);
            float4 m = *reinterpret_cast<float4*>(r.m + j), v = *reinterpret_cast<float4*>(r.v + j),
                   vm = *reinterpret_cast<float4*>(r.vmax + j);
            adam_update(p.x, g.x, m.x, v.x, vm.x, step_size, sqrt_bc2, beta1, beta2, eps, wd, gscale);
            adam_update(p.y, g.y, m.y, v.y, vm.y, step_size, sqrt_bc2, beta1, beta2, eps, wd, gscale);
            adam_update(p.z, g.z, m.z, v.z, vm.z, step_size, sqrt_bc2, beta1, beta2, eps, wd, gscale);
            adam_update(p.w, g.w, m.w, v.w, vm.w, step_size, sqrt_bc2, beta1, beta2, eps, wd, gscale);
            *reinterpret_cast<float4*>(r.m + j) = m;
            *reinterpret_cast<float4*>(r.v + j) = v;
            *reinterpret_cast<float4*>(r.vmax + j) = vm;
            *reinterpret_cast<float4*>(r.p + j) = p;
        }
        i = end4 + threadIdx.x;
    }
    for (; i < end; i += 256) {
        float p = r.p[i], m = r.m[i], v = r.v[i], vm = r.vmax[i];
        adam_update(p, r.g[i], m, v, vm, step_size, sqrt_bc2, beta1, beta2, eps, wd, gscale);
        r.m[i] = m; r.v[i] = v; r.vmax[i] = vm;
        r.p[i] = p;
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void mixup_kernel(const float* __restrict__ pcm, const float* __restrict__ labels, const int* __restrict__ partner,
                             int n, long long t, int c, float* pcm_out, float* labels_out) {
    int i = blockIdx.y;
    int j = partner[i];
    const float* a = pcm + (long long)i * t;
    float* o = pcm_out + (long long)i * t;
    if (j < 0) {
        for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < t; k += (long long)gridDim.x * blockDim.x)
            o[k] = a[k];
    } else {
        const float* b = pcm + (long long)j * t;
        for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < t; k += (long long)gridDim.x * blockDim.x)
            o[k] = (a[k] + b[k]) / 2.0f;                  // ops/audio.py:40-41
    }
    if (blockIdx.x == 0) {
        for (int k = threadIdx.x; k < c; k += blockDim.x) {
            float l = labels[(long long)i * c + k];
            if (j >= 0) l = fminf(fmaxf(l + labels[(long long)j * c + k], 0.f), 1.f);   // np.clip(l1 + l2, 0, 1)
            labels_out[(long long)i * c + k] = l;
        }
    }
}

}  // namespace fsb

using namespace fsb;

extern "C" int fsb_version(void) { return 100; }
extern "C" const char* fsb_last_error(void) { return g_err; }

extern "C" int fsb_device_ok(void) {
    // cached per device: cudaGetDeviceProperties costs milliseconds and this sits on the per-step path
    static int cached[64];       // 0 = unknown, 1 = ok, 2 = wrong architecture
    static int cached_major[64], cached_minor[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("no CUDA device");
        return FSB_E_NODEVICE;
    }
    if (dev < 0 || dev >= 64) dev = 63;
    if (!cached[dev]) {
        int major = 0, minor = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
            set_error("no CUDA device");
            return FSB_E_NODEVICE;
        }
        cached_major[dev] = major; cached_minor[dev] = minor;
        cached[dev] = major == 10 ? 1 : 2;
    }
    if (cached[dev] != 1) {
        set_error("libfsb200 is built for sm_100a only; device is sm_%d%d", cached_major[dev], cached_minor[dev]);
        return FSB_E_NODEVICE;
    }
    return 0;
}

extern "C" long long fsb_launch_count(int reset) {
    long long v = g_launch_count;
    if (reset) g_launch_count = 0;
    return v;
}

extern "C" int fsb_lsep_forward(const float* scores, const float* targets, int n, int c, float* loss, void* stream) {
    FSB_REQUIRE(n > 0 && c > 0, "lsep: empty input");
    int blocks = (n * 32 + 127) / 128;
    lsep_fwd_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(scores, targets, n, c, loss);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_lsep_backward(const float* scores, const float* targets, const float* dloss, int n, int c,
                                 float* dscores, void* stream) {
    FSB_REQUIRE(n > 0 && c > 0, "lsep: empty input");
    int blocks = (n * 32 + 127) / 128;
    lsep_bwd_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(scores, targets, dloss, n, c, dscores);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_lsep_stable_forward(const float* scores, const float* targets, int n, int c, float* loss, void* stream) {
    FSB_REQUIRE(n > 0 && c > 0, "lsep: empty input");
    int blocks = (n * 32 + 127) / 128;
    lsep_fwd_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(scores, targets, n, c, loss);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_lsep_stable_backward(const float* scores, const float* targets, const float* dloss, int n, int c,
                                        float* dscores, void* stream) {
    FSB_REQUIRE(n > 0 && c > 0, "lsep: empty input");
    int blocks = (n * 32 + 127) / 128;
    lsep_bwd_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(scores, targets, dloss, n, c, dscores);
    FSB_LAUNCHED();
    return 0;
}

extern "C" size_t fsb_lwlrap_scratch_bytes(int n) { return (size_t)(n > 0 ? n : 0) * 2 * sizeof(double); }

extern "C" int fsb_lwlrap(const float* truth, const float* scores, int n, int c, int accumulate, void* scratch,
                          double* out, void* stream) {
    FSB_REQUIRE(n > 0 && c > 0 && truth && scores && scratch && out, "lwlrap: bad arguments");
    double* row_num = (double*)scratch;
    double* row_den = row_num + n;
    int blocks = (n * 32 + 127) / 128;
    lwlrap_rows_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(truth, scores, n, c, row_num, row_den);
    FSB_LAUNCHED();
    lwlrap_sum_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(row_num, row_den, n, out, accumulate);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_assemble_batch(const float* pool, const float* label_pool, const void* rows, int n, int c,
                                  long long t_out, float pad_value, float* out, float* labels_out, void* stream) {
    FSB_REQUIRE(pool && label_pool && rows && out && labels_out, "assemble: null argument");
    FSB_REQUIRE(n > 0 && n <= 65535 && t_out > 0 && c > 0, "assemble: bad shape (n=%d, t_out=%lld)", n, t_out);
    static_assert(sizeof(AssembleRow) == 56, "AssembleRow must match the 56-byte host record");
    dim3 grid(148, n);
    assemble_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pool, label_pool, (const AssembleRow*)rows, c, t_out, pad_value,
                                                           out, labels_out);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_adam_chunk(void) { return ADAM_CHUNK; }

extern "C" int fsb_adam_amsgrad_step(const void* table, const int* block_map, int n_blocks, int step, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                     void* stream) {
    FSB_REQUIRE(n_blocks > 0 && step >= 1, "adam: bad arguments");
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    float step_size = (float)((double)lr / bc1);
    float sqrt_bc2 = (float)sqrt(bc2);
    adam_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>((const AdamRec*)table, block_map, step_size, sqrt_bc2,
                                                           beta1, beta2, eps, weight_decay, grad_scale);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_mixup_equal(const float* pcm, const float* labels, const int* partner, int n, long long t, int c,
                               float* pcm_out, float* labels_out, void* stream) {
    FSB_REQUIRE(n > 0 && n <= 65535 && t > 0, "mixup: bad shape");
    dim3 grid(148, n);
    mixup_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pcm, labels, partner, n, t, c, pcm_out, labels_out);
    FSB_LAUNCHED();
    return 0;
}
