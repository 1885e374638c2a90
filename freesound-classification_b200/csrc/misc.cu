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

__global__ void lsep_fwd_kernel(const float* __restrict__ s, const float* __restrict__ t, int n, int c, float* loss) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const float* sr = s + (long long)warp * c;
    const float* tr = t + (long long)warp * c;
    float acc = 0.f;
    for (int i = 0; i < c; ++i) {
        float si = sr[i], ti = tr[i];
        for (int j = lane; j < c; j += 32)
            if (tr[j] < ti) acc += expf(sr[j] - si);
    }
    acc = warp_sum(acc);
    if (lane == 0) loss[warp] = logf(1.0f + acc);
}

// dL/ds_k = ( sum_{i: t_k < t_i} e^{s_k - s_i}  -  sum_{j: t_j < t_k} e^{s_j - s_k} ) / (1 + S)
__global__ void lsep_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t, const float* __restrict__ dloss,
                                int n, int c, float* ds) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const float* sr = s + (long long)warp * c;
    const float* tr = t + (long long)warp * c;
    float total = 0.f;
    for (int k = lane; k < c; k += 32) {
        float sk = sr[k], tk = tr[k];
        for (int i = 0; i < c; ++i)
            if (tk < tr[i]) total += expf(sk - sr[i]);
    }
    total = warp_sum(total);
    float inv = dloss[warp] / (1.0f + total);
    for (int k = lane; k < c; k += 32) {
        float sk = sr[k], tk = tr[k];
        float g = 0.f;
        for (int i = 0; i < c; ++i) {
            float ti = tr[i];
            if (tk < ti) g += expf(sk - sr[i]);
            else if (ti < tk) g -= expf(sr[i] - sk);
        }
        ds[(long long)warp * c + k] = g * inv;
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
    lsep_fwd_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(scores, targets, n, c, loss);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_lsep_backward(const float* scores, const float* targets, const float* dloss, int n, int c,
                                 float* dscores, void* stream) {
    FSB_REQUIRE(n > 0 && c > 0, "lsep: empty input");
    int blocks = (n * 32 + 127) / 128;
    lsep_bwd_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(scores, targets, dloss, n, c, dscores);
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
