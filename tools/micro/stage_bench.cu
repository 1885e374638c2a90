// Issue cost of the per-stage MMA asm blocks of the conv GEMM (no barriers, operands = zeros in smem): cycles per stage
// for the generic predicated block vs the straight-line specialisation.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../freesound-classification_b200/csrc -o stage_bench stage_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma_issue.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) bench(int N, int T, int ks, int taps, int fast, int iters, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint32_t row = 64u;                    // SWIZZLE_64B stages (bk = 32)
        const uint32_t a_box = 136u * row, a_tile = 2u * a_box, a_part = a_tile * T, w_plane = (uint32_t)N * row, w_tap = 2u * w_plane;
        const uint32_t stage = a_part + w_tap * taps;
        const uint32_t d_hi = ((8u * row >> 4) & 0x3FFFu) | (1u << 14) | (4u << 29), d_lo = 1u << 16;
        const uint32_t idesc = make_idesc(128, N);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t b = base + (it & 1) * stage;
            if (fsb::umma::elect_one()) {
                const uint64_t a_hi = ((uint64_t)d_hi << 32) | (d_lo | ((b & 0x3FFFFu) >> 4));
                const uint64_t b_hi = ((uint64_t)d_hi << 32) | (d_lo | (((b + a_part) & 0x3FFFFu) >> 4));
                if (fast) fsb::umma::umma_stage_x3_fast(tmem, N, a_hi, a_tile >> 4, a_box >> 4, row >> 4, b_hi, w_tap >> 4, w_plane >> 4, idesc, it ? 1u : 0u, ks, T, taps);
                else fsb::umma::umma_stage_x3(tmem, N, a_hi, a_tile >> 4, a_box >> 4, row >> 4, b_hi, w_tap >> 4, w_plane >> 4, idesc, it ? 1u : 0u, ks, T, taps);
            }
            __syncwarp();
        }
        if (fsb::umma::elect_one()) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const int iters = 400;
    const size_t smem = 220 * 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Case { int N, T, ks, taps; } cases[] = {{112, 2, 2, 3}, {160, 1, 2, 3}, {240, 1, 4, 1}, {176, 1, 4, 1}, {256, 1, 2, 1}, {112, 2, 4, 1}};
    printf("%-22s %8s %10s %10s %10s\n", "stage", "MMAs", "pipe cyc", "generic", "straight");
    for (auto c : cases) {
        double r[2];
        for (int fast = 0; fast < 2; ++fast) {
            for (int rep = 0; rep < 2; ++rep) {
                bench<<<148, 128, smem>>>(c.N, c.T, c.ks, c.taps, fast, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            r[fast] = (double)h / iters;
        }
        int mmas = c.T * c.ks * c.taps * 3;
        double pipe = mmas * (c.N / 2.0 > 44 ? (c.N >= 128 ? c.N / 2.0 : 32 + c.N / 4.0) : 44);
        char name[64];
        snprintf(name, sizeof name, "N=%d T=%d ks=%d taps=%d", c.N, c.T, c.ks, c.taps);
        printf("%-22s %8d %10.0f %10.0f %10.0f\n", name, mmas, pipe, r[0], r[1]);
    }
    return 0;
}
