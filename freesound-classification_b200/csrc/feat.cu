// K-feat: fused reflect-pad framing + periodic Hann + real FFT (radix-4 Stockham in shared memory)
// + magnitude + banded mel projection + log.  One launch replaces the reference's torch.stft /
// pow / sum / sqrt / conv1d / log chain (ops/utils.py:110-127, networks/classifiers.py:565-579).
//
// Work decomposition: one CTA of 256 threads owns FPB = 16 consecutive frames of one clip.  A real
// n_fft-point transform is computed as an M = n_fft/2 point complex transform of the even/odd packed
// samples followed by the split post-processing; 256 threads execute M/4 radix-4 butterflies per
// stage, so PAR = 1024/M frames are transformed concurrently.
//
// PCM staging: the CTA's window of raw samples is a contiguous run of the clip.  It streams through a three-slot
// shared-memory ring in chunks of PAR * hop samples (one group of concurrently transformed frames advances the window
// by exactly one chunk): an elected thread issues 1-D bulk TMA copies (`cp.async.bulk`, completion counted on one
// mbarrier per slot) two chunks ahead of the group being transformed, so HBM latency hides behind the previous
// group's FFT and every sample is read from HBM once per CTA and from shared memory by each frame that overlaps it.
// CTAs whose window needs reflected samples (the first and the last of a clip) or whose global address / hop is not
// 16-byte aligned read the samples with plain loads and per-sample reflection instead.
// Outputs are staged in shared memory and flushed with the contiguous dimension fastest.
#include <stdlib.h>

#include "common.cuh"

namespace fsb {

static constexpr int FEAT_THREADS = 256;
static constexpr int FPB = 16;           // frames per CTA
static constexpr int FEAT_RING = 3;      // PCM ring slots
static constexpr int FEAT_MAX_CHUNK = 2048;   // floats per ring slot the ring path supports (PAR * hop)

// tw[n] = (cos(2 pi n / n_fft), -sin(2 pi n / n_fft)), evaluated in double
__global__ void feat_tables_kernel(float2* tw, int n_fft) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_fft) {
        double s, c;
        sincospi(2.0 * (double)i / (double)n_fft, &s, &c);
        tw[i] = make_float2((float)c, (float)(-s));
    }
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ uint32_t feat_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void feat_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

template <int LOG2N>
__global__ void __launch_bounds__(FEAT_THREADS)
feat_kernel(const float* __restrict__ pcm, long long pcm_stride, int T, int hop, int frames, int ring_chunk,
            int mode, float eps, int n_mel, const float* __restrict__ fb_vals,
            const int* __restrict__ fb_off, const int* __restrict__ fb_start,
            const int* __restrict__ fb_len, const float2* __restrict__ tw_g, float* __restrict__ out,
            long long out_sn, long long out_sf, long long out_st) {
    constexpr int NFFT = 1 << LOG2N;
    constexpr int M = NFFT / 2;           // complex transform size
    constexpr int LOG2M = LOG2N - 1;
    constexpr int PAR = 1024 / M;         // frames transformed concurrently (power of two, >= 1)
    constexpr int BINS = M + 1;
    static_assert(M <= 1024 && M >= 64, "supported n_fft: 128 .. 2048");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);            // NFFT
    // per-stage radix-4 twiddles (w^k, w^2k, w^3k), k < Ns, stored contiguously in k: lanes read consecutive
    // entries.  Reading them from `tw` with the stride (NFFT/4)/Ns put up to 16 lanes on one bank.
    float2* tws = tw + NFFT;                                     // 3 * (1 + 4 + 16 + ...) <= M entries
    float2* buf0 = tws + M;                                      // PAR * M = 1024
    float2* buf1 = buf0 + 1024;                                  // 1024
    float* mag = reinterpret_cast<float*>(buf1 + 1024);          // PAR * (BINS + 1)
    const int f_out = (mode == 2) ? n_mel : BINS;
    float* tile = mag + PAR * (BINS + 1);                        // f_out * (FPB + 1)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(
        smem_raw + (((size_t)((unsigned char*)(tile + (size_t)f_out * (FPB + 1)) - smem_raw) + 15) & ~(size_t)15));
    float* ring = reinterpret_cast<float*>(bars + 4);            // FEAT_RING * ring_chunk floats, 16-byte aligned

    const int tid = threadIdx.x;
    const int clip = blockIdx.y;
    const int frame0 = blockIdx.x * FPB;
    const int nfr = min(FPB, frames - frame0);
    const float* x = pcm + (long long)clip * pcm_stride;

    // ---- PCM window of this CTA: samples [s0, s0 + wlen) of the clip; ring path only when no reflection is needed
    const long long s0 = (long long)frame0 * hop - M;            // n_fft/2 == M
    const int wlen = (nfr - 1) * hop + NFFT;
    const int CH = ring_chunk;                                   // PAR * hop, or 0: ring path disabled for this launch
    const bool use_ring = CH > 0 && s0 >= 0 && s0 + wlen <= T && ((reinterpret_cast<size_t>(x + s0) & 15) == 0);
    const int nchunks = use_ring ? (wlen + CH - 1) / CH : 0;
    auto issue_chunk = [&](int c) {                              // one thread: bulk copy of chunk c into slot c % 3
        const uint32_t b = feat_smem_u32(bars + (c % FEAT_RING));
        const int n = min(CH, wlen - c * CH);                    // multiple of 4 floats (hop % 4 == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)n * 4u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(feat_smem_u32(ring + (size_t)(c % FEAT_RING) * CH)), "l"(x + s0 + (long long)c * CH),
                       "r"((uint32_t)n * 4u), "r"(b)
                     : "memory");
    };
    if (tid == 0 && use_ring) {
        for (int i = 0; i < FEAT_RING; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(feat_smem_u32(bars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue_chunk(0);
        if (nchunks > 1) issue_chunk(1);
    }

    for (int i = tid; i < NFFT; i += FEAT_THREADS) tw[i] = tw_g[i];
    {
        int base = 0, Ns = 1;
#pragma unroll
        for (int st = 0; st < LOG2M / 2; ++st) {
            for (int e = tid; e < 3 * Ns; e += FEAT_THREADS) {
                const int k = e / 3, i = e - 3 * k;
                tws[base + e] = tw_g[(i + 1) * ((NFFT / 4) / Ns * k)];
            }
            base += 3 * Ns;
            Ns <<= 2;
        }
    }
    __syncthreads();

    for (int it = 0; it < nfr; it += PAR) {
        // ---- 1. windowed frames -> packed complex input (natural order): z[j] = (x[2j] w[2j], x[2j+1] w[2j+1])
        if (use_ring) {
            const int g = it / PAR;                              // group g reads chunks g and g + 1
            // slot (g + 2) % 3 held chunk g - 1, last read by group g - 1 whose reads completed before the barrier
            // that ended its step 1: refill it now, two chunks ahead
            if (tid == 0 && g + 2 < nchunks) issue_chunk(g + 2);
            feat_bar_wait(feat_smem_u32(bars + (g % FEAT_RING)), (uint32_t)((g / FEAT_RING) & 1));
            if (g + 1 < nchunks) feat_bar_wait(feat_smem_u32(bars + ((g + 1) % FEAT_RING)), (uint32_t)(((g + 1) / FEAT_RING) & 1));
            const float* c0 = ring + (size_t)(g % FEAT_RING) * CH;
            const float* c1 = ring + (size_t)((g + 1) % FEAT_RING) * CH;
            for (int e = tid; e < PAR * M; e += FEAT_THREADS) {
                const int slot = e / M, j = e - slot * M;
                float2 v = make_float2(0.f, 0.f);
                if (it + slot < nfr) {
                    const int local = slot * hop + 2 * j;        // sample offset inside chunk g (may run into g + 1)
                    const float2 xv = *reinterpret_cast<const float2*>(local < CH ? c0 + local : c1 + (local - CH));
                    v.x = xv.x * (0.5f - 0.5f * tw[2 * j].x);
                    v.y = xv.y * (0.5f - 0.5f * tw[2 * j + 1].x);
                }
                buf0[e] = v;
            }
        } else {
            for (int e = tid; e < PAR * M; e += FEAT_THREADS) {
                int slot = e / M, j = e - slot * M;
                int fr = frame0 + it + slot;
                float2 v = make_float2(0.f, 0.f);
                if (it + slot < nfr) {
                    int i0 = fr * hop - M + 2 * j;
                    int i1 = i0 + 1;
                    i0 = i0 < 0 ? -i0 : (i0 >= T ? 2 * (T - 1) - i0 : i0);
                    i1 = i1 < 0 ? -i1 : (i1 >= T ? 2 * (T - 1) - i1 : i1);
                    float w0 = 0.5f - 0.5f * tw[2 * j].x;
                    float w1 = 0.5f - 0.5f * tw[2 * j + 1].x;
                    v.x = __ldg(x + i0) * w0;
                    v.y = __ldg(x + i1) * w1;
                }
                buf0[e] = v;
            }
        }
        __syncthreads();

        // ---- 2. Stockham autosort FFT, radix 4 (+ one radix-2 stage when log2(M) is odd)
        float2* src = buf0;
        float2* dst = buf1;
        {
            const int slot = tid / (M / 4);
            const int j = tid - slot * (M / 4);
            const float2* s0p = nullptr;
            int Ns = 1, tbase = 0;
#pragma unroll
            for (int st = 0; st < LOG2M / 2; ++st) {
                s0p = src + slot * M;
                float2* d0 = dst + slot * M;
                int k = j & (Ns - 1);
                const float2* t3 = tws + tbase + 3 * k;      // exp(-2 pi i m k/(4 Ns)), m = 1, 2, 3
                tbase += 3 * Ns;
                float2 v0 = s0p[j];
                float2 v1 = cmul(s0p[j + M / 4], t3[0]);
                float2 v2 = cmul(s0p[j + M / 2], t3[1]);
                float2 v3 = cmul(s0p[j + 3 * M / 4], t3[2]);
                float2 a = make_float2(v0.x + v2.x, v0.y + v2.y);
                float2 b = make_float2(v0.x - v2.x, v0.y - v2.y);
                float2 c = make_float2(v1.x + v3.x, v1.y + v3.y);
                float2 d = make_float2(v1.x - v3.x, v1.y - v3.y);
                int idx = ((j - k) << 2) + k;
                d0[idx] = make_float2(a.x + c.x, a.y + c.y);
                d0[idx + Ns] = make_float2(b.x + d.y, b.y - d.x);
                d0[idx + 2 * Ns] = make_float2(a.x - c.x, a.y - c.y);
                d0[idx + 3 * Ns] = make_float2(b.x - d.y, b.y + d.x);
                __syncthreads();
                float2* t = src; src = dst; dst = t;
                Ns <<= 2;
            }
            if (LOG2M & 1) {
                // Ns == M/2 here: out[k] = in[k] + w^k in[k + M/2], out[k + M/2] = in[k] - ...
                for (int e = tid; e < PAR * (M / 2); e += FEAT_THREADS) {
                    int sl = e / (M / 2), jj = e - sl * (M / 2);
                    const float2* s1 = src + sl * M;
                    float2* d1 = dst + sl * M;
                    int k = jj & (Ns - 1);
                    float2 v0 = s1[jj];
                    float2 v1 = cmul(s1[jj + M / 2], tw[(NFFT / 2) / Ns * k]);
                    int idx = ((jj - k) << 1) + k;
                    d1[idx] = make_float2(v0.x + v1.x, v0.y + v1.y);
                    d1[idx + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
                }
                __syncthreads();
                float2* t = src; src = dst; dst = t;
            }
        }

        // ---- 3. split post-processing -> one-sided spectrum magnitude
        for (int e = tid; e < PAR * BINS; e += FEAT_THREADS) {
            int slot = e / BINS, k = e - slot * BINS;
            const float2* z = src + slot * M;
            float2 zk = z[k & (M - 1)];
            float2 zm = z[(M - k) & (M - 1)];
            float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);     // even-sample spectrum
            float orr = 0.5f * (zk.y + zm.y), oi = -0.5f * (zk.x - zm.x);   // odd-sample spectrum
            float2 w = (k == M) ? make_float2(-1.f, 0.f) : tw[k];
            float re = er + (orr * w.x - oi * w.y);
            float im = ei + (orr * w.y + oi * w.x);
            mag[slot * (BINS + 1) + k] = sqrtf(re * re + im * im);
        }
        __syncthreads();

        // ---- 4. mel projection / log into the staging tile
        if (mode == 2) {
            for (int e = tid; e < PAR * n_mel; e += FEAT_THREADS) {
                int slot = e / n_mel, m = e - slot * n_mel;
                if (it + slot < nfr) {
                    const float* mg = mag + slot * (BINS + 1) + fb_start[m];
                    const float* fv = fb_vals + fb_off[m];
                    int len = fb_len[m];
                    float acc = 0.f;
                    for (int i = 0; i < len; ++i) acc = fmaf(__ldg(fv + i), mg[i], acc);
                    tile[m * (FPB + 1) + it + slot] = logf(acc + eps);
                }
            }
        } else {
            for (int e = tid; e < PAR * BINS; e += FEAT_THREADS) {
                int slot = e / BINS, k = e - slot * BINS;
                if (it + slot < nfr) {
                    float v = mag[slot * (BINS + 1) + k];
                    tile[k * (FPB + 1) + it + slot] = (mode == 1) ? logf(v + eps) : v;
                }
            }
        }
        __syncthreads();
    }

    // ---- 5. flush the (f_out x nfr) tile, contiguous output dimension fastest
    float* o = out + (long long)clip * out_sn;
    if (out_st == 1) {
        for (int e = tid; e < f_out * FPB; e += FEAT_THREADS) {
            int f = e / FPB, s = e - f * FPB;
            if (s < nfr) o[(long long)f * out_sf + (frame0 + s)] = tile[f * (FPB + 1) + s];
        }
    } else {
        for (int e = tid; e < f_out * nfr; e += FEAT_THREADS) {
            int s = e / f_out, f = e - s * f_out;
            o[(long long)f * out_sf + (long long)(frame0 + s) * out_st] = tile[f * (FPB + 1) + s];
        }
    }
}

template <int LOG2N>
static int launch_feat(const float* pcm, int n, long long pcm_stride, int t, int hop, int mode,
                       float eps, int n_mel, const float* fb_vals, const int* fb_off,
                       const int* fb_start, const int* fb_len, const void* tables, float* out,
                       long long sn, long long sf, long long st, cudaStream_t stream) {
    int n_fft = 1 << LOG2N;
    int frames = 1 + t / hop;
    int f_out = mode == 2 ? n_mel : n_fft / 2 + 1;
    const int m = n_fft / 2, par = 1024 / m, bins = m + 1;
    // ring path: one chunk = PAR * hop samples; a group's frames must fit in two consecutive chunks, chunk copies must be
    // whole 16-byte units and every row of the PCM matrix must keep the 16-byte alignment of the first
    int chunk = par * hop;
    const bool ring_ok = hop % 4 == 0 && n_fft - hop <= chunk && chunk <= FEAT_MAX_CHUNK && pcm_stride % 4 == 0 &&
                         (reinterpret_cast<size_t>(pcm) & 15) == 0;
    static int ring_env = -1;            // FSB200_FEAT_RING=0: plain-load path everywhere (A/B timing)
    if (ring_env < 0) { const char* e = getenv("FSB200_FEAT_RING"); ring_env = e ? atoi(e) : 1; }
    if (!ring_ok || !ring_env) chunk = 0;
    const size_t smem = (size_t)n_fft * 8 + (size_t)m * 8 + 2 * 1024 * 8 + (size_t)par * (bins + 1) * 4 +
                        (size_t)f_out * (FPB + 1) * 4 + 16 + 32 + (size_t)FEAT_RING * chunk * 4;
    FSB_REQUIRE(smem <= 227 * 1024, "feat: shared memory %zu too large", smem);
    auto kern = feat_kernel<LOG2N>;
    static bool attr_set = false;        // per instantiation
    if (!attr_set) {
        FSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((frames + FPB - 1) / FPB, n);
    kern<<<grid, FEAT_THREADS, smem, stream>>>(pcm, pcm_stride, t, hop, frames, chunk, mode, eps, n_mel,
                                               fb_vals, fb_off, fb_start, fb_len,
                                               (const float2*)tables, out, sn, sf, st);
    FSB_LAUNCHED();
    return 0;
}

}  // namespace fsb

using namespace fsb;

extern "C" size_t fsb_feat_table_bytes(int n_fft) { return (size_t)n_fft * sizeof(float2); }

extern "C" int fsb_feat_init_tables(int n_fft, void* tables, void* stream) {
    FSB_REQUIRE(n_fft >= 128 && n_fft <= 2048 && (n_fft & (n_fft - 1)) == 0,
                "feat: n_fft must be a power of two in [128, 2048], got %d", n_fft);
    feat_tables_kernel<<<(n_fft + 255) / 256, 256, 0, (cudaStream_t)stream>>>((float2*)tables, n_fft);
    FSB_LAUNCHED();
    return 0;
}

extern "C" int fsb_feat_forward(const float* pcm, int n, long long pcm_stride, int t, int n_fft,
                                int hop, int mode, float eps, int n_mel, const float* fb_vals,
                                const int* fb_off, const int* fb_start, const int* fb_len,
                                const void* tables, float* out, long long out_sn, long long out_sf,
                                long long out_st, void* stream) {
    FSB_REQUIRE(n > 0 && t > n_fft / 2, "feat: need T > n_fft/2 (T=%d, n_fft=%d)", t, n_fft);
    FSB_REQUIRE(hop > 0 && mode >= 0 && mode <= 2, "feat: bad hop/mode");
    FSB_REQUIRE(mode != 2 || (n_mel > 0 && fb_vals && fb_off && fb_start && fb_len),
                "feat: mel mode needs a filterbank");
    FSB_REQUIRE(n <= 65535, "feat: at most 65535 clips per launch");
    cudaStream_t s = (cudaStream_t)stream;
#define FSB_FEAT_CASE(L)                                                                         \
    case (1 << L):                                                                               \
        return launch_feat<L>(pcm, n, pcm_stride, t, hop, mode, eps, n_mel, fb_vals, fb_off,     \
                              fb_start, fb_len, tables, out, out_sn, out_sf, out_st, s);
    switch (n_fft) {
        FSB_FEAT_CASE(7)
        FSB_FEAT_CASE(8)
        FSB_FEAT_CASE(9)
        FSB_FEAT_CASE(10)
        FSB_FEAT_CASE(11)
        default:
            set_error("feat: unsupported n_fft %d (power of two in [128, 2048])", n_fft);
            return FSB_E_INVALID;
    }
#undef FSB_FEAT_CASE
}
