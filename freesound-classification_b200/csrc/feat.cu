// K-feat: fused reflect-pad framing + periodic Hann + real FFT (radix-4 Stockham in shared memory)
// + magnitude + banded mel projection + log.  One launch replaces the reference's torch.stft /
// pow / sum / sqrt / conv1d / log chain (ops/utils.py:110-127, networks/classifiers.py:565-579).
//
// Two kernels: `feat2048_mel_kernel` (n_fft = 2048 + mel + log, the 2D model's features: register-resident radix-16
// passes, four frames per group, see its own header further down) and the generic `feat_kernel<LOG2N>` described here
// (every other n_fft / mode, e.g. the 1D model's `stft_256_128`).
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

// ---------------------------------------------------------------------------------------------------------------
// n_fft = 2048 + mel + log (the 2D model's `mel_2048_1024_128`): register-resident radix-16 butterflies.
// M = 1024 = 16 x 16 x 4: two radix-16 passes and one radix-4 pass instead of five radix-4 passes, FOUR frames per group
// (256 threads = 4 frames x 64 radix-16 butterflies), ONE buffer used in place (a pass reads its 16 inputs into
// registers, all threads synchronise, then it writes), the Hann window applied while the first pass reads the PCM ring,
// the magnitudes written in place over the spectrum.  Six barriers per four frames instead of nine per frame, and half
// the shared-memory round trips per frame.  Logical index i of the buffer lives at i + (i >> 4) (the first pass writes
// with stride 16).
static constexpr int F2K_PAR = 4;
static constexpr int F2K_M = 1024;
static constexpr int F2K_ROW = F2K_M + F2K_M / 16;       // padded frame row (float2)
static constexpr int F2K_RING = 2;                        // PCM ring slots (chunk g + 2 is requested once chunk g is consumed)
static constexpr int F2K_SEG = 8;                         // mel projection: filter taps per work item
static constexpr int F2K_MAX_SEGS = 512;                  // work items per frame (sum over bands of ceil(len / 8))
static constexpr int F2K_MAX_TAPS = 2304;                 // filterbank values staged in shared memory

__device__ __forceinline__ int f2k_ph(int i) { return i + (i >> 4); }

__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s0 = make_float2(a0.x + a2.x, a0.y + a2.y), s1 = make_float2(a0.x - a2.x, a0.y - a2.y);
    const float2 s2 = make_float2(a1.x + a3.x, a1.y + a3.y), s3 = make_float2(a1.x - a3.x, a1.y - a3.y);
    a0 = make_float2(s0.x + s2.x, s0.y + s2.y);
    a1 = make_float2(s1.x + s3.y, s1.y - s3.x);
    a2 = make_float2(s0.x - s2.x, s0.y - s2.y);
    a3 = make_float2(s1.x - s3.y, s1.y + s3.x);
}

// in-register 16-point DFT (forward): x[n] -> X[k], both in natural order.  n = 4 n1 + n2, k = k1 + 4 k2.
__device__ __forceinline__ void dft16(float2 (&x)[16]) {
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
    // A: four 4-point transforms over n1 (inputs x[n2], x[4 + n2], x[8 + n2], x[12 + n2]) -> y[n2][k1] in x[4 k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);
    // B: twiddles W16^(n2 k1) on y[n2][k1] = x[4 k1 + n2]
    x[5] = cmul(x[5], make_float2(C1, -S1));      // n2 = 1, k1 = 1
    x[6] = cmul(x[6], make_float2(R, -R));        // n2 = 2, k1 = 1
    x[7] = cmul(x[7], make_float2(S1, -C1));      // n2 = 3, k1 = 1
    x[9] = cmul(x[9], make_float2(R, -R));        // n2 = 1, k1 = 2
    x[10] = make_float2(x[10].y, -x[10].x);       // n2 = 2, k1 = 2: W^4 = -i
    x[11] = cmul(x[11], make_float2(-R, -R));     // n2 = 3, k1 = 2: W^6
    x[13] = cmul(x[13], make_float2(S1, -C1));    // n2 = 1, k1 = 3: W^3
    x[14] = cmul(x[14], make_float2(-R, -R));     // n2 = 2, k1 = 3: W^6
    x[15] = cmul(x[15], make_float2(-C1, S1));    // n2 = 3, k1 = 3: W^9
    // C: four 4-point transforms over n2 -> X[k1 + 4 k2] lands in x[4 k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
}
// natural-order output index of register slot s after dft16: slot 4 k1 + k2 holds X[k1 + 4 k2]
__device__ __forceinline__ int dft16_out(int s) { return (s >> 2) + 4 * (s & 3); }

__global__ void __launch_bounds__(FEAT_THREADS)
feat2048_mel_kernel(const float* __restrict__ pcm, long long pcm_stride, int T, int hop, int frames, int ring_chunk,
                    float eps, int n_mel, const float* __restrict__ fb_vals, const int* __restrict__ fb_off,
                    const int* __restrict__ fb_start, const int* __restrict__ fb_len, const float2* __restrict__ tw_g,
                    float* __restrict__ out, long long out_sn, long long out_sf, long long out_st) {
    constexpr int NFFT = 2048, M = F2K_M, PAR = F2K_PAR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);            // [0, M]: exp(-2 pi i n / NFFT)
    float2* t2 = tw + (M + 8);                                   // [15][16]: second pass, exp(-2 pi i r k / 256)
    float2* t3 = t2 + 15 * 16;                                   // [3][256]: third pass, exp(-2 pi i m k / 1024)
    float2* buf = t3 + 3 * 256;                                  // PAR padded frame rows
    float* tile = reinterpret_cast<float*>(buf + PAR * F2K_ROW); // n_mel * (FPB + 1)
    // mel projection as evenly sized work items: band m is cut into segments of <= F2K_SEG taps; seg_first[m] is the
    // band's first segment, seg_band[s] the band of segment s; part[frame][s] holds the segment sums
    float* fbv = tile + (size_t)n_mel * (FPB + 1);               // F2K_MAX_TAPS filter values
    float* part = fbv + F2K_MAX_TAPS;                            // PAR * F2K_MAX_SEGS
    int* seg_first = reinterpret_cast<int*>(part + PAR * F2K_MAX_SEGS);   // n_mel + 1 (<= 257)
    unsigned short* seg_band = reinterpret_cast<unsigned short*>(seg_first + 268);   // F2K_MAX_SEGS
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(
        smem_raw + (((size_t)((unsigned char*)(seg_band + F2K_MAX_SEGS) - smem_raw) + 15) & ~(size_t)15));
    float* ring = reinterpret_cast<float*>(bars + 4);

    const int tid = threadIdx.x;
    const int clip = blockIdx.y;
    const int frame0 = blockIdx.x * FPB;
    const int nfr = min(FPB, frames - frame0);
    const float* x = pcm + (long long)clip * pcm_stride;

    const long long s0 = (long long)frame0 * hop - M;
    const int wlen = (nfr - 1) * hop + NFFT;
    const int CH = ring_chunk;                                   // PAR * hop
    const bool use_ring = CH > 0 && s0 >= 0 && s0 + wlen <= T && ((reinterpret_cast<size_t>(x + s0) & 15) == 0);
    const int nchunks = use_ring ? (wlen + CH - 1) / CH : 0;
    auto issue_chunk = [&](int c) {
        const uint32_t b = feat_smem_u32(bars + (c % F2K_RING));
        const int n = min(CH, wlen - c * CH);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)n * 4u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(feat_smem_u32(ring + (size_t)(c % F2K_RING) * CH)), "l"(x + s0 + (long long)c * CH),
                       "r"((uint32_t)n * 4u), "r"(b)
                     : "memory");
    };
    if (tid == 0 && use_ring) {
        for (int i = 0; i < F2K_RING; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(feat_smem_u32(bars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue_chunk(0);
        if (nchunks > 1) issue_chunk(1);
    }
    for (int i = tid; i <= M; i += FEAT_THREADS) tw[i] = tw_g[i];
    for (int e = tid; e < 15 * 16; e += FEAT_THREADS) t2[e] = tw_g[(8 * (e / 16 + 1) * (e % 16)) & (NFFT - 1)];
    for (int e = tid; e < 3 * 256; e += FEAT_THREADS) t3[e] = tw_g[2 * (e / 256 + 1) * (e % 256)];
    // segment table of the mel projection (bands are 2 .. 64 taps wide: one thread per band would leave most of the CTA
    // waiting for the widest ones) and the filter values
    {
        // exclusive prefix sum of the per-band segment counts (n_mel <= 256 = one band per thread): warp scans + the
        // warp totals through seg_first[256 ..]
        const int cnt = tid < n_mel ? (fb_len[tid] + F2K_SEG - 1) / F2K_SEG : 0;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, inc, o);
            if ((tid & 31) >= o) inc += up;
        }
        if ((tid & 31) == 31) seg_first[257 + (tid >> 5)] = inc;      // scratch behind the table (260 ints + 8)
        __syncthreads();
        int base = 0;
        for (int w = 0; w < (tid >> 5); ++w) base += seg_first[257 + w];
        const int excl = base + inc - cnt;
        __syncthreads();
        if (tid < n_mel) seg_first[tid] = excl;
        if (tid == n_mel - 1) seg_first[n_mel] = excl + cnt;
    }
    __syncthreads();
    const int ntaps = fb_off[n_mel - 1] + fb_len[n_mel - 1];
    const int nseg = seg_first[n_mel];
    const bool seg_ok = ntaps <= F2K_MAX_TAPS && nseg <= F2K_MAX_SEGS;    // else: one thread per band, values from global
    if (seg_ok) {
        for (int m = tid; m < n_mel; m += FEAT_THREADS)
            for (int sg = seg_first[m]; sg < seg_first[m + 1]; ++sg) seg_band[sg] = (unsigned short)m;
        for (int i = tid; i < ntaps; i += FEAT_THREADS) fbv[i] = fb_vals[i];
    }
    __syncthreads();

    const int f = tid >> 6, j = tid & 63;                        // radix-16 passes: frame slot, butterfly
    float2* row = buf + f * F2K_ROW;
    for (int it = 0; it < nfr; it += PAR) {
        const bool live = it + f < nfr;
        float2 v[16];
        // ---- pass 1 (Ns = 1): windowed packed samples z[j + 64 r] straight from the ring, 16-point DFT, out[16 j + q]
        if (use_ring) {
            const int g = it / PAR;
            feat_bar_wait(feat_smem_u32(bars + (g % F2K_RING)), (uint32_t)((g / F2K_RING) & 1));
            if (g + 1 < nchunks) feat_bar_wait(feat_smem_u32(bars + ((g + 1) % F2K_RING)), (uint32_t)(((g + 1) / F2K_RING) & 1));
            const float* c0 = ring + (size_t)(g % F2K_RING) * CH;
            const float* c1 = ring + (size_t)((g + 1) % F2K_RING) * CH;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int n = j + 64 * r;
                v[r] = make_float2(0.f, 0.f);
                if (live) {
                    const int local = f * hop + 2 * n;
                    const float2 xv = *reinterpret_cast<const float2*>(local < CH ? c0 + local : c1 + (local - CH));
                    // window index by symmetry; r < 8 <=> 2 n + 1 < M (r is a compile-time constant here)
                    const int i0 = r < 8 ? 2 * n : NFFT - 2 * n, i1 = r < 8 ? 2 * n + 1 : NFFT - 2 * n - 1;
                    v[r].x = xv.x * (0.5f - 0.5f * tw[i0].x);
                    v[r].y = xv.y * (0.5f - 0.5f * tw[i1].x);
                }
            }
        } else {
            const int fr = frame0 + it + f;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int n = j + 64 * r;
                v[r] = make_float2(0.f, 0.f);
                if (live) {
                    int i0 = fr * hop - M + 2 * n, i1 = i0 + 1;
                    i0 = i0 < 0 ? -i0 : (i0 >= T ? 2 * (T - 1) - i0 : i0);
                    i1 = i1 < 0 ? -i1 : (i1 >= T ? 2 * (T - 1) - i1 : i1);
                    const int w0 = r < 8 ? 2 * n : NFFT - 2 * n, w1 = r < 8 ? 2 * n + 1 : NFFT - 2 * n - 1;
                    v[r].x = __ldg(x + i0) * (0.5f - 0.5f * tw[w0].x);
                    v[r].y = __ldg(x + i1) * (0.5f - 0.5f * tw[w1].x);
                }
            }
        }
        dft16(v);
#pragma unroll
        for (int s = 0; s < 16; ++s) row[f2k_ph(16 * j + dft16_out(s))] = v[s];
        __syncthreads();
        // chunk g has been consumed by every thread: its slot takes chunk g + 2 (needed by the next group's second half)
        if (use_ring && tid == 0 && it / PAR + 2 < nchunks) issue_chunk(it / PAR + 2);

        // ---- pass 2 (Ns = 16): in[j + 64 r] * w^(r k), k = j % 16, out[(j - k) 16 + k + 16 q]
        {
            const int k = j & 15;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                v[r] = row[f2k_ph(j + 64 * r)];
                if (r > 0) v[r] = cmul(v[r], t2[(r - 1) * 16 + k]);
            }
            dft16(v);
            __syncthreads();                                     // in place: every read precedes every write
#pragma unroll
            for (int s = 0; s < 16; ++s) row[f2k_ph(((j - k) << 4) + k + 16 * dft16_out(s))] = v[s];
        }
        __syncthreads();

        // ---- pass 3 (radix 4, Ns = 256): butterfly k = tid of every frame, in place on its own four elements
        {
            const float2 w1 = t3[tid], w2 = t3[256 + tid], w3 = t3[512 + tid];
#pragma unroll
            for (int ff = 0; ff < PAR; ++ff) {
                float2* rw = buf + ff * F2K_ROW;
                float2 a0 = rw[f2k_ph(tid)];
                float2 a1 = cmul(rw[f2k_ph(tid + 256)], w1);
                float2 a2 = cmul(rw[f2k_ph(tid + 512)], w2);
                float2 a3 = cmul(rw[f2k_ph(tid + 768)], w3);
                dft4(a0, a1, a2, a3);
                rw[f2k_ph(tid)] = a0;
                rw[f2k_ph(tid + 256)] = a1;
                rw[f2k_ph(tid + 512)] = a2;
                rw[f2k_ph(tid + 768)] = a3;
            }
        }
        __syncthreads();

        // ---- split post-processing -> magnitudes, written over the spectrum: bin k in .x of element k (k < M), bin M in
        //      .y of element 0.  A thread owns the pair (k, M - k).
        for (int k = j; k <= M / 2; k += 64) {                  // thread (f, j): pairs k = j, j + 64, ... of frame f
            float2* rw = row;
            const int km = (M - k) & (M - 1);
            const float2 zk = rw[f2k_ph(k)], zm = rw[f2k_ph(km)];
            auto magnitude = [&](float2 a, float2 b, float2 w) {
                const float er = 0.5f * (a.x + b.x), ei = 0.5f * (a.y - b.y);
                const float orr = 0.5f * (a.y + b.y), oi = -0.5f * (a.x - b.x);
                const float re = er + (orr * w.x - oi * w.y), im = ei + (orr * w.y + oi * w.x);
                float r2 = re * re + im * im, rt;
                asm("sqrt.approx.f32 %0, %1;" : "=f"(rt) : "f"(r2));      // <= 2 ulp; the parity gate is 1e-5 relative
                return rt;
            };
            if (k == 0) {
                const float m0 = magnitude(zk, zk, tw[0]), mM = magnitude(zk, zk, make_float2(-1.f, 0.f));
                rw[f2k_ph(0)] = make_float2(m0, mM);
            } else if (k == M / 2) {
                rw[f2k_ph(k)].x = magnitude(zk, zk, tw[k]);
            } else {
                const float mk = magnitude(zk, zm, tw[k]), mm = magnitude(zm, zk, tw[M - k]);
                rw[f2k_ph(k)].x = mk;
                rw[f2k_ph(km)].x = mm;
            }
        }
        __syncthreads();

        // ---- banded mel projection: segment sums (<= F2K_SEG taps each, evenly spread over the threads), then one
        //      thread per (frame, band) adds its band's segments in order, + log, into the staging tile
        for (int sg = tid; seg_ok && sg < nseg; sg += FEAT_THREADS) {
            const int m = seg_band[sg], o8 = (sg - seg_first[m]) * F2K_SEG;
            const int b0 = fb_start[m] + o8, len = min(F2K_SEG, fb_len[m] - o8);
            const float* fv = fbv + fb_off[m] + o8;
            float acc[PAR];
#pragma unroll
            for (int ff = 0; ff < PAR; ++ff) acc[ff] = 0.f;
#pragma unroll
            for (int i = 0; i < F2K_SEG; ++i) {
                if (i < len) {
                    const int b = b0 + i;
                    const float w = fv[i];
                    const int pb = b < M ? 2 * f2k_ph(b) : 1;          // float index of the magnitude inside a frame row
#pragma unroll
                    for (int ff = 0; ff < PAR; ++ff)
                        acc[ff] = fmaf(w, reinterpret_cast<const float*>(buf + ff * F2K_ROW)[pb], acc[ff]);
                }
            }
#pragma unroll
            for (int ff = 0; ff < PAR; ++ff) part[ff * F2K_MAX_SEGS + sg] = acc[ff];
        }
        __syncthreads();
        for (int e = tid; e < PAR * n_mel; e += FEAT_THREADS) {
            const int ff = e / n_mel, m = e - ff * n_mel;
            if (it + ff < nfr) {
                float acc = 0.f;
                if (seg_ok) {
                    for (int sg = seg_first[m]; sg < seg_first[m + 1]; ++sg) acc += part[ff * F2K_MAX_SEGS + sg];
                } else {
                    const float2* rw = buf + ff * F2K_ROW;
                    const int b0 = fb_start[m], len = fb_len[m];
                    const float* fv = fb_vals + fb_off[m];
                    for (int i = 0; i < len; ++i) {
                        const int b = b0 + i;
                        acc = fmaf(__ldg(fv + i), b < M ? rw[f2k_ph(b)].x : rw[0].y, acc);
                    }
                }
                tile[m * (FPB + 1) + it + ff] = logf(acc + eps);
            }
        }
        __syncthreads();
    }

    float* o = out + (long long)clip * out_sn;
    if (out_st == 1) {
        for (int e = tid; e < n_mel * FPB; e += FEAT_THREADS) {
            int fq = e / FPB, sidx = e - fq * FPB;
            if (sidx < nfr) o[(long long)fq * out_sf + (frame0 + sidx)] = tile[fq * (FPB + 1) + sidx];
        }
    } else {
        for (int e = tid; e < n_mel * nfr; e += FEAT_THREADS) {
            int sidx = e / n_mel, fq = e - sidx * n_mel;
            o[(long long)fq * out_sf + (long long)(frame0 + sidx) * out_st] = tile[fq * (FPB + 1) + sidx];
        }
    }
}

static int launch_feat2048_mel(const float* pcm, int n, long long pcm_stride, int t, int hop, float eps, int n_mel,
                               const float* fb_vals, const int* fb_off, const int* fb_start, const int* fb_len,
                               const void* tables, float* out, long long sn, long long sf, long long st, int chunk,
                               cudaStream_t stream) {
    const int frames = 1 + t / hop;
    const size_t smem = (size_t)(F2K_M + 8) * 8 + 15 * 16 * 8 + 3 * 256 * 8 + (size_t)F2K_PAR * F2K_ROW * 8 +
                        (size_t)n_mel * (FPB + 1) * 4 + (size_t)F2K_MAX_TAPS * 4 + (size_t)F2K_PAR * F2K_MAX_SEGS * 4 +
                        268 * 4 + F2K_MAX_SEGS * 2 + 16 + 32 + (size_t)F2K_RING * chunk * 4;
    FSB_REQUIRE(smem <= 227 * 1024, "feat: shared memory %zu too large", smem);
    static bool attr_set = false;
    if (!attr_set) {
        FSB_CUDA(cudaFuncSetAttribute(feat2048_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    dim3 grid((frames + FPB - 1) / FPB, n);
    feat2048_mel_kernel<<<grid, FEAT_THREADS, smem, stream>>>(pcm, pcm_stride, t, hop, frames, chunk, eps, n_mel, fb_vals,
                                                              fb_off, fb_start, fb_len, (const float2*)tables, out, sn, sf, st);
    FSB_LAUNCHED();
    return 0;
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
    static int r16_env = -1;             // FSB200_FEAT_R16=0: the radix-4 kernel for n_fft = 2048 too (A/B timing)
    if (r16_env < 0) { const char* e = getenv("FSB200_FEAT_R16"); r16_env = e ? atoi(e) : 1; }
    if (LOG2N == 11 && mode == 2 && r16_env && n_mel <= 256) {
        // register-resident radix-16 kernel, four frames per group: ring chunk = 4 hops (a group's frames span
        // 3 hops + n_fft samples, which must fit in two consecutive chunks)
        int c4 = F2K_PAR * hop;
        const bool ok4 = hop % 4 == 0 && n_fft <= (F2K_PAR + 1) * hop && c4 <= 8192 && pcm_stride % 4 == 0 &&
                         (reinterpret_cast<size_t>(pcm) & 15) == 0;
        if (!ok4 || !ring_env) c4 = 0;
        return launch_feat2048_mel(pcm, n, pcm_stride, t, hop, eps, n_mel, fb_vals, fb_off, fb_start, fb_len, tables, out,
                                   sn, sf, st, c4, stream);
    }
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
