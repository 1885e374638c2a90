// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N, of the
// number of accumulators the stream rotates over, and of the smem layout (SWIZZLE_128B / SWIZZLE_64B).
// Operands are whatever is in shared memory (zeros).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

// mode: number of accumulators rotated over (1, 2, 4); unroll 8 MMAs per loop iteration
template <int NACC>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int layout, long long* out, int shift_rows = 0) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
    if (threadIdx.x == 0) {
        const uint32_t row = layout == 2 ? 128u : 64u;
        const uint32_t sbo = 8u * row;
        const uint32_t idesc = make_idesc(128, N);
        // A boxes: 4 x (136 rows), B boxes after them
        const uint32_t a0 = base, b0 = base + 4u * 136u * row * 2u;
        uint64_t a[4], b[2];
        for (int t = 0; t < 4; ++t) a[t] = make_desc(a0 + t * 136u * row * 2u + (uint32_t)shift_rows * row, sbo, layout);
        b[0] = make_desc(b0, sbo, layout);
        b[1] = make_desc(b0 + 256u * row, sbo, layout);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = j % NACC;
                umma(tmem + (uint32_t)(t * N), a[t] + (uint64_t)((j / NACC) & 1) * 2, b[j & 1], idesc);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
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
    const int iters = 2000;
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int Ns[] = {16, 32, 64, 96, 112, 128, 160, 176, 240, 256};
    printf("cycles per MMA (M=128, K=16, bf16); ideal = N/2\n%6s %8s | %10s %10s %10s | %10s\n", "N", "ideal", "1 acc SW128", "2 acc", "4 acc", "4 acc SW64");
    for (int N : Ns) {
        double r[4];
        for (int m = 0; m < 4; ++m) {
            int layout = m == 3 ? 4 : 2;
            int nacc = m == 0 ? 1 : (m == 1 ? 2 : 4);
            if (nacc * N > 512) { r[m] = -1; continue; }
            for (int rep = 0; rep < 2; ++rep) {
                if (nacc == 1) bench<1><<<148, 128, smem>>>(N, iters, layout, d);
                else if (nacc == 2) bench<2><<<148, 128, smem>>>(N, iters, layout, d);
                else bench<4><<<148, 128, smem>>>(N, iters, layout, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            r[m] = (double)h / (iters * 8.0);
        }
        printf("%6d %8.1f | %10.1f %10.1f %10.1f | %10.1f\n", N, N / 2.0, r[0], r[1], r[2], r[3]);
    }
    printf("\nA operand start shifted by s rows (row-shifted conv taps), 2 accumulators\n%6s | %8s %8s %8s %8s | SW64: %8s %8s\n", "N", "s=0", "s=1", "s=2", "s=4", "s=0", "s=1");
    for (int N : Ns) {
        if (2 * N > 512) continue;
        double r[6];
        int shifts[6] = {0, 1, 2, 4, 0, 1};
        for (int m = 0; m < 6; ++m) {
            int layout = m >= 4 ? 4 : 2;
            for (int rep = 0; rep < 2; ++rep) {
                bench<2><<<148, 128, smem>>>(N, iters, layout, d, shifts[m]);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long h;
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            r[m] = (double)h / (iters * 8.0);
        }
        printf("%6d | %8.1f %8.1f %8.1f %8.1f | %14.1f %8.1f\n", N, r[0], r[1], r[2], r[3], r[4], r[5]);
    }
    return 0;
}
