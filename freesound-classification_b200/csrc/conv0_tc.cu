// Block-0 entry convolution of the 2D network on the tensor cores (networks/classifiers.py:524-532 for k = 0):
// BatchNorm2d(2) -> Conv2d(2 -> C0, 3x3, pad 1) -> MaxPool2d(2), fused.  K = 2 * 9 = 18 is far too small for a TMA-fed
// row-shifted GEMM (the generic conv kernel would spend a 16-wide k-step per tap and write the un-pooled 1.4 GB
// tensor), so this kernel builds the im2col operand itself:
//
//   tile = 128 POOLED pixels.  For each of the four pool-window positions the builder warps write a K-major
//   [128 x 32] half tile (18 patch values of the BN-applied 2-channel input, zero padded to two k-steps; hi and lo
//   planes, x = hi + lo) straight into shared memory in the SWIZZLE_64B layout the UMMA descriptors expect -- channel 1
//   (the frequency encoding, networks/classifiers.py:553-561) is synthesised, never loaded.  One elected thread issues
//   the tile's 4 positions x 2 k-steps x 3 products as M128 x N(C0 padded) x K16 tcgen05 MMAs into FOUR TMEM
//   accumulators (one per window position); the epilogue takes the max over the four accumulators, records the
//   arg-max position for the backward pass, adds the bias and writes the pooled row.
//   The operand tiles are double buffered: building tile i + 1 overlaps the MMAs of tile i.
//
// The weight-gradient / BatchNorm-input gradient pass stays on the CUDA cores (conv0.cu: with the stored arg-max it is
// 18 FMAs per pooled pixel and channel).
#include "conv0.cuh"
#include "tc_ptx.cuh"
#include "umma_issue.cuh"

namespace fsb {
namespace {

constexpr int C0T_THREADS = 160;                  // warps 0-3: operand builders + epilogue, warp 4: MMA issuer
constexpr int C0T_PX = 128;                       // pooled pixels per tile = UMMA M = TMEM lanes
constexpr uint32_t C0T_PLANE = C0T_PX * 64;       // one [128 x 32] half plane, 64-byte rows
constexpr uint32_t C0T_POS = 2 * C0T_PLANE;       // hi + lo
constexpr uint32_t C0T_ABUF = 4 * C0T_POS;        // four window positions

struct Conv0TcParams {
    const float* feat;
    int N, H, W;
    const float* scale;
    const float* shift;
    const float* w;
    const float* b;
    float* zp;
    unsigned char* amax;
    Geo gp;
    int BN;                 // padded output channels (gp.Cs, multiple of 16, <= 128)
    int planes;             // 2: three products (hi + lo operands), 1: single pass
    long long npix;         // pooled pixels
    int ntiles;
};

__device__ __forceinline__ float c0t_freq_enc(int h, int H) {
    const float step = 2.0f / (float)(H - 1);
    return h < H / 2 ? -1.0f + step * (float)h : 1.0f - step * (float)(H - 1 - h);
}

// 16-byte chunk `c` (0..3) of 64-byte row `r` of a SWIZZLE_64B tile (Swizzle<2,4,3>: address bits [4,6) ^= bits [7,9))
__device__ __forceinline__ uint32_t sw64(uint32_t r, uint32_t c) { return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4); }

// writes K values v[0..17] (zero padded to 32) as row r of the hi (and lo) plane at smem address `base`
__device__ __forceinline__ void write_k_row(uint32_t base, uint32_t r, const float* v, int planes) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        __half h0, l0, h1, l1;
        split_h16(v[2 * i], h0, l0);
        split_h16(v[2 * i + 1], h1, l1);
        hi[i] = pack_h2(h0, h1);
        lo[i] = pack_h2(l0, l1);
    }
#pragma unroll
    for (int i = 9; i < 16; ++i) hi[i] = lo[i] = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        st_shared_v4_u32(base + sw64(r, c), hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
        if (planes == 2)
            st_shared_v4_u32(base + C0T_PLANE + sw64(r, c), lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
    }
}

__global__ void __launch_bounds__(C0T_THREADS, 1) conv0_tc_fwd_kernel(const Conv0TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_buf0 = smem_base;                                   // two operand buffers
    const uint32_t b_tile = a_buf0 + 2u * C0T_ABUF;                      // weights: hi plane, lo plane ([BN x 32] each)
    const uint32_t w_plane = C0T_PLANE;                                  // plane stride of the weight tile
    const uint32_t bars = b_tile + 2u * 128u * 64u;                      // a_full[2], acc_full, tmem slot
    const uint32_t b_afull = bars, b_acc = bars + 16u, tmem_slot = bars + 24u;
    float* bias_s = reinterpret_cast<float*>(smem_raw + (bars + 64u - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float sc0 = p.scale[0], sh0 = p.shift[0], sc1 = p.scale[1], sh1 = p.shift[1];

    // ---- one-time set-up: weight tile (K-major, row = output channel, k = ci * 9 + dy * 3 + dx), bias, barriers, TMEM
    for (int n = threadIdx.x; n < p.BN; n += C0T_THREADS) {
        float v[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) v[k] = n < p.gp.C ? p.w[n * 18 + k] : 0.f;
        write_k_row(b_tile, (uint32_t)n, v, 2);          // lo plane C0T_PLANE bytes after the hi plane
        bias_s[n] = n < p.gp.C ? p.b[n] : 0.f;
    }
    if (threadIdx.x == 0) {
        mbar_init(b_afull, 128);
        mbar_init(b_afull + 8u, 128);
        mbar_init(b_acc, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    fence_async_smem();                       // the weight tile was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int H2 = p.gp.H, W2 = p.gp.W;
    if (warp < 4) {
        const uint32_t r = threadIdx.x;                        // tile row = TMEM lane
        uint32_t it = 0;
        long long prev_q = -1;                                 // pooled pixel of the tile whose accumulators are in flight
        auto epilogue = [&](long long q, uint32_t parity) {
            mbar_wait(b_acc, parity);
            tc_fence_after();
            const bool live = q >= 0 && q < p.npix;            // rows past the last pixel: TMEM loads are warp-wide, stores are not
            float* zrow = nullptr;
            unsigned char* arow = nullptr;
            if (live) {
                const int px = (int)(q % W2);
                const long long t = q / W2;
                const int py = (int)(t % H2), n = (int)(t / H2);
                const long long row = geo_row(p.gp, n, py, px);
                zrow = p.zp + row * p.gp.Cs;
                arow = p.amax ? p.amax + row * p.gp.Cs : nullptr;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                float v0[16], v1[16], v2[16], v3[16];
                tmem_ld16_async(taddr + (uint32_t)c0, v0);
                tmem_ld16_async(taddr + (uint32_t)(p.BN + c0), v1);
                tmem_ld16_async(taddr + (uint32_t)(2 * p.BN + c0), v2);
                tmem_ld16_async(taddr + (uint32_t)(3 * p.BN + c0), v3);
                tmem_ld_wait();
                float o[16];
                uint32_t pos4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    // first maximum in window scan order (0,0) (0,1) (1,0) (1,1), like nn.MaxPool2d
                    float best = v0[i];
                    uint32_t bp = 0u;
                    if (v1[i] > best) { best = v1[i]; bp = 1u; }
                    if (v2[i] > best) { best = v2[i]; bp = 2u; }
                    if (v3[i] > best) { best = v3[i]; bp = 3u; }
                    o[i] = best + bias_s[c0 + i];
                    pos4[i >> 2] |= bp << (8 * (i & 3));
                }
                if (live) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(zrow + c0 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
                    if (arow) *reinterpret_cast<uint4*>(arow + c0) = make_uint4(pos4[0], pos4[1], pos4[2], pos4[3]);
                }
            }
            tc_fence_before();
        };
        // raw input patch of a tile, loaded one tile ahead: the loads complete behind the previous tile's epilogue
        auto load_patch = [&](int tile, float (&f)[4][4], int& px, int& py, bool& live) {
            const long long q = (long long)tile * C0T_PX + r;
            live = tile < p.ntiles && q < p.npix;
            px = 0; py = 0;
            if (live) {
                px = (int)(q % W2);
                const long long t = q / W2;
                py = (int)(t % H2);
                const int n = (int)(t / H2);
                const float* img = p.feat + (long long)n * p.H * p.W;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int y = 2 * py - 1 + i;
                    const bool yok = y >= 0 && y < p.H;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int x = 2 * px - 1 + j;
                        f[i][j] = (yok && x >= 0 && x < p.W) ? __ldg(img + (long long)y * p.W + x) : 0.f;
                    }
                }
            }
        };
        float fr[4][4];
        int cpx, cpy;
        bool clive;
        load_patch(blockIdx.x, fr, cpx, cpy, clive);
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t buf = a_buf0 + (it & 1u) * C0T_ABUF;
            const long long q = (long long)tile * C0T_PX + r;
            // ---- BN-applied 4 x 4 x 2 input patch of this pooled pixel (zero outside the image = conv padding)
            float u0[4][4], u1[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int y = 2 * cpy - 1 + i;
                const bool yok = clive && y >= 0 && y < p.H;
                const float e1 = yok ? fmaf(c0t_freq_enc(y, p.H), sc1, sh1) : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = 2 * cpx - 1 + j;
                    const bool ok = yok && x >= 0 && x < p.W;
                    u0[i][j] = ok ? fmaf(fr[i][j], sc0, sh0) : 0.f;
                    u1[i][j] = ok ? e1 : 0.f;
                }
            }
            load_patch(tile + (int)gridDim.x, fr, cpx, cpy, clive);      // the next tile's loads fly during the epilogue below
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                const int sy = pos >> 1, sx = pos & 1;
                float v[18];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        v[dy * 3 + dx] = u0[sy + dy][sx + dx];
                        v[9 + dy * 3 + dx] = u1[sy + dy][sx + dx];
                    }
                write_k_row(buf + (uint32_t)pos * C0T_POS, r, v, p.planes);
            }
            fence_async_smem();                                // generic-proxy writes -> visible to the tensor core
            // the accumulators of the previous tile must be drained before this tile's MMAs overwrite them
            if (it > 0) epilogue(prev_q, (it - 1u) & 1u);
            mbar_arrive(b_afull + 8u * (it & 1u));
            prev_q = q;
        }
        if (it > 0) epilogue(prev_q, (it - 1u) & 1u);
    } else {
        // ---- MMA issuer: the whole warp walks the tiles, one elected lane issues
        const uint32_t idesc = make_idesc(C0T_PX, p.BN, 0, 0);
        const uint32_t d_lo = 1u << 16;
        const uint32_t d_hi = ((512u >> 4) & 0x3FFFu) | (1u << 14) | (4u << 29);       // SBO = 8 rows x 64 B, SWIZZLE_64B
        const uint64_t b_desc = ((uint64_t)d_hi << 32) | (d_lo | ((b_tile & 0x3FFFFu) >> 4));
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const uint32_t buf = a_buf0 + (it & 1u) * C0T_ABUF;
            mbar_wait(b_afull + 8u * (it & 1u), (it >> 1) & 1u);
            tc_fence_after();
            if (umma::elect_one()) {
                const uint64_t a_desc = ((uint64_t)d_hi << 32) | (d_lo | ((buf & 0x3FFFFu) >> 4));
                bool ok = true;
#pragma unroll
                for (int half = 0; half < 2; ++half) {        // window positions {0, 1} then {2, 3}: two "row tiles" per block
                    const uint64_t a2 = a_desc + (uint64_t)(half * 2) * (C0T_POS >> 4);
                    const uint32_t d2 = tmem_base + (uint32_t)(half * 2 * p.BN);
                    ok = ok && (p.planes == 2
                        ? umma::umma_stage_x3(d2, (uint32_t)p.BN, a2, C0T_POS >> 4, C0T_PLANE >> 4, 0u, b_desc, 0u, w_plane >> 4,
                                              idesc, 0u, 2, 2, 1)
                        : umma::umma_stage_x1(d2, (uint32_t)p.BN, a2, C0T_POS >> 4, C0T_PLANE >> 4, 0u, b_desc, 0u, w_plane >> 4,
                                              idesc, 0u, 2, 2, 1));
                }
                if (!ok) __trap();
                umma_commit(b_acc);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    (void)lane;
}

// ---------------------------------------------------------------------------------------------------------------
// backward: D[c][k] = sum over pooled pixels q of g[q][c] * [amax[q][c] == pos] * F_pos[q][k], summed over the four
// window positions, where F_pos[q][0..17] is the BN-applied input patch of position pos, F_pos[q][18..26] flags the taps
// that fall inside the image and F_pos[q][27..44] is the normalised input patch (xhat).  D[c][0..17] is the weight
// gradient; d(beta_in)_ci = sum_c sum_t w[c][ci][t] D[c][18 + t] and d(gamma_in)_ci = sum_c sum_t w[c][ci][t] D[c][27 + 9 ci + t].  Per 64-pixel tile: four masked-gradient tiles [64 pixels x 128 channels] (M side, MN-major,
// two 64-channel SWIZZLE_128B boxes) and four patch tiles [64 pixels x 64 features] (N side), 4 positions x 4 k-steps
// of M128 x N64 x K16 MMAs into ONE accumulator that lives in TMEM for the whole life of the CTA.
constexpr int C0B_PX = 64;                         // pooled pixels per tile (reduction dimension of the MMAs)
constexpr uint32_t C0B_BOX = C0B_PX * 128u;        // one 64-channel box: 64 rows x 128 B
constexpr int C0B_THREADS = 160;

struct Conv0TcBwdParams {
    const float* feat;
    int N, H, W;
    const float* scale;
    const float* shift;
    const float* mean;
    const float* invstd;
    const float* w;
    const float* dzp;
    int dzp_half;           // 1: dzp is ONE half plane that already carries the GradScale of `absmax` (compact backward)
    const unsigned char* amax;
    const unsigned* absmax;
    Geo gp;
    int planes;
    long long npix;
    int ntiles;
    float* partials;        // [gridDim.x][22][Cs]
};

// 16-byte chunk c (0..7) of 128-byte row r of a SWIZZLE_128B box
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

__global__ void __launch_bounds__(C0B_THREADS, 2) conv0_tc_bwd_kernel(const Conv0TcBwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // [G: position][plane][box 0 | box 1]   [F: position][plane][box]
    const uint32_t g_plane = 2u * C0B_BOX, g_pos = g_plane * (uint32_t)p.planes;
    const uint32_t f_plane = C0B_BOX, f_pos = f_plane * (uint32_t)p.planes;
    const uint32_t g_base = smem_base, f_base = g_base + 4u * g_pos;
    const uint32_t bars = f_base + 4u * f_pos;
    const uint32_t b_full = bars, b_empty = bars + 8u, tmem_slot = bars + 16u;
    const int warp = threadIdx.x >> 5;
    const float sc0 = p.scale[0], sh0 = p.shift[0], sc1 = p.scale[1], sh1 = p.shift[1];
    const float mean0 = p.mean[0], istd0 = p.invstd[0], mean1 = p.mean[1], istd1 = p.invstd[1];

    // zero the operand region once: channel chunks >= Cs and feature chunks 6..7 are never written afterwards
    for (uint32_t o = threadIdx.x * 16u; o < 4u * (g_pos + f_pos); o += C0B_THREADS * 16u)
        st_shared_v4_u32(g_base + o, 0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        mbar_init(b_full, 128);
        mbar_init(b_empty, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 64);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int H2 = p.gp.H, W2 = p.gp.W, Cs = p.gp.Cs;
    uint32_t it = 0;
    if (warp < 4) {
        const uint32_t r = threadIdx.x & 63u, half = threadIdx.x >> 6;       // pixel row of the tile, channel half / position pair
        const float gscale = gs_scale(p.absmax);
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const long long q = (long long)tile * C0B_PX + r;
            const bool live = q < p.npix;
            int px = 0, py = 0, n = 0;
            if (live) {
                px = (int)(q % W2);
                const long long t = q / W2;
                py = (int)(t % H2);
                n = (int)(t / H2);
            }
            const long long grow = live ? geo_row(p.gp, n, py, px) * Cs : 0;
            // ---- every global load of the tile is issued BEFORE the wait for the operand buffers (one round trip per
            //      tile, overlapping the previous tile's MMAs; the loop below used to pay one per 8-channel chunk)
            float fr[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int y = 2 * py - 1 + i;
                const bool yok = live && y >= 0 && y < p.H;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = 2 * px - 1 + j;
                    fr[i][j] = (yok && x >= 0 && x < p.W) ? __ldg(p.feat + ((long long)n * p.H + y) * p.W + x) : 0.f;
                }
            }
            uint4 graw[8];
            uint2 araw[8];
            if (p.dzp_half) {
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const int ch0 = (int)(half * 64u + c * 8u);
                    graw[c] = make_uint4(0u, 0u, 0u, 0u);
                    araw[c] = make_uint2(0u, 0u);
                    if (live && ch0 < Cs) {
                        graw[c] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.dzp) + grow + ch0);
                        araw[c] = *reinterpret_cast<const uint2*>(p.amax + grow + ch0);
                    }
                }
            }
            if (it > 0) mbar_wait(b_empty, (it - 1u) & 1u);                   // the previous tile's MMAs have read the operands
            // ---- masked gradient rows: channels [half * 64, half * 64 + 64) of the four positions
            if (p.dzp_half) {
                // the half plane is the hi operand as it stands (same GradScale slot, single pass: no lo plane); the
                // arg-max bytes become 16-bit lane masks (byte compare, byte permute)
#pragma unroll
                for (uint32_t c = 0; c < 8; ++c) {
                    const int ch0 = (int)(half * 64u + c * 8u);
                    if (ch0 < Cs) {
#pragma unroll
                        for (uint32_t pos = 0; pos < 4; ++pos) {
                            const uint32_t m_lo = __vcmpeq4(araw[c].x, pos * 0x01010101u), m_hi = __vcmpeq4(araw[c].y, pos * 0x01010101u);
                            const uint32_t dst = g_base + pos * g_pos + half * C0B_BOX + sw128(r, c);
                            st_shared_v4_u32(dst, graw[c].x & __byte_perm(m_lo, 0u, 0x1100u), graw[c].y & __byte_perm(m_lo, 0u, 0x3322u),
                                             graw[c].z & __byte_perm(m_hi, 0u, 0x1100u), graw[c].w & __byte_perm(m_hi, 0u, 0x3322u));
                        }
                    }
                }
            } else {
#pragma unroll
            for (uint32_t c = 0; c < 8; ++c) {
                const int ch0 = (int)(half * 64u + c * 8u);
                if (ch0 < Cs) {
                    uint32_t a8[2] = {0u, 0u};
                    __half gh[8], gl[8];
                    {
                        float g8[8];
                        if (live) {
                            const float4 lo4 = *reinterpret_cast<const float4*>(p.dzp + grow + ch0);
                            const float4 hi4 = *reinterpret_cast<const float4*>(p.dzp + grow + ch0 + 4);
                            g8[0] = lo4.x; g8[1] = lo4.y; g8[2] = lo4.z; g8[3] = lo4.w;
                            g8[4] = hi4.x; g8[5] = hi4.y; g8[6] = hi4.z; g8[7] = hi4.w;
                            const uint2 am = *reinterpret_cast<const uint2*>(p.amax + grow + ch0);
                            a8[0] = am.x; a8[1] = am.y;
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) g8[i] = 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) split_h16(g8[i] * gscale, gh[i], gl[i]);
                    }
#pragma unroll
                    for (uint32_t pos = 0; pos < 4; ++pos) {
                        uint32_t wh[4], wl[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const bool m0 = ((a8[(2 * i) >> 2] >> (8 * ((2 * i) & 3))) & 0xFFu) == pos;
                            const bool m1 = ((a8[(2 * i + 1) >> 2] >> (8 * ((2 * i + 1) & 3))) & 0xFFu) == pos;
                            const __half z = __float2half_rn(0.f);
                            wh[i] = pack_h2(m0 ? gh[2 * i] : z, m1 ? gh[2 * i + 1] : z);
                            wl[i] = pack_h2(m0 ? gl[2 * i] : z, m1 ? gl[2 * i + 1] : z);
                        }
                        const uint32_t dst = g_base + pos * g_pos + half * C0B_BOX + sw128(r, c);
                        st_shared_v4_u32(dst, wh[0], wh[1], wh[2], wh[3]);
                        if (p.planes == 2) st_shared_v4_u32(dst + g_plane, wl[0], wl[1], wl[2], wl[3]);
                    }
                }
            }
            }
            // ---- patch feature rows of positions 2 * half and 2 * half + 1 (window rows half .. half + 2, window columns
            //      pp .. pp + 2); every index into the register arrays is a compile-time constant
            float fw[3][4], ew[3];
            uint32_t inw[3];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int y = 2 * py - 1 + (int)half + dy;
                const bool yok = live && y >= 0 && y < p.H;
                ew[dy] = yok ? c0t_freq_enc(y, p.H) : 0.f;
                inw[dy] = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = 2 * px - 1 + j;
                    fw[dy][j] = half ? fr[dy + 1][j] : fr[dy][j];
                    if (yok && x >= 0 && x < p.W) inw[dy] |= 1u << j;
                }
            }
#pragma unroll
            for (uint32_t pp = 0; pp < 2; ++pp) {
                const uint32_t pos = 2u * half + pp;
                float v[48];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const bool in = (inw[dy] >> (pp + dx)) & 1u;
                        const float f = fw[dy][pp + dx], e = ew[dy];
                        v[dy * 3 + dx] = in ? fmaf(f, sc0, sh0) : 0.f;
                        v[9 + dy * 3 + dx] = in ? fmaf(e, sc1, sh1) : 0.f;
                        v[18 + dy * 3 + dx] = in ? 1.f : 0.f;
                        v[27 + dy * 3 + dx] = in ? (f - mean0) * istd0 : 0.f;
                        v[36 + dy * 3 + dx] = in ? (e - mean1) * istd1 : 0.f;
                    }
#pragma unroll
                for (int i = 45; i < 48; ++i) v[i] = 0.f;
#pragma unroll
                for (uint32_t c = 0; c < 6; ++c) {
                    uint32_t wh[4], wl[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __half h0, l0, h1, l1;
                        split_h16(v[8 * c + 2 * i], h0, l0);
                        split_h16(v[8 * c + 2 * i + 1], h1, l1);
                        wh[i] = pack_h2(h0, h1);
                        wl[i] = pack_h2(l0, l1);
                    }
                    const uint32_t dst = f_base + pos * f_pos + sw128(r, c);
                    st_shared_v4_u32(dst, wh[0], wh[1], wh[2], wh[3]);
                    if (p.planes == 2) st_shared_v4_u32(dst + f_plane, wl[0], wl[1], wl[2], wl[3]);
                }
            }
            fence_async_smem();
            mbar_arrive(b_full);
        }
        // ---- epilogue: this thread's channel row of the accumulator -> the 22-value record of conv0.cu
        if (it > 0) mbar_wait(b_empty, (it - 1u) & 1u);
        tc_fence_after();
        const int c = threadIdx.x;                                             // TMEM lane = output channel
        float d[48];
        tmem_ld32_async(tmem_base + ((uint32_t)(warp * 32) << 16), d);
        tmem_ld16_async(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u, d + 32);
        tmem_ld_wait();
        const float unscale = gs_inv_scale(p.absmax);
        float* o = p.partials + (long long)blockIdx.x * 22 * Cs;
        if (c < Cs) {
            float rec[22];
#pragma unroll
            for (int i = 0; i < 22; ++i) rec[i] = 0.f;
            if (c < p.gp.C && it > 0) {
                float dg0 = 0.f, dg1 = 0.f, db0 = 0.f, db1 = 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float w0 = p.w[c * 18 + t], w1 = p.w[c * 18 + 9 + t];
                    rec[t] = d[t] * unscale;
                    rec[9 + t] = d[9 + t] * unscale;
                    db0 = fmaf(w0, d[18 + t], db0);
                    db1 = fmaf(w1, d[18 + t], db1);
                    dg0 = fmaf(w0, d[27 + t], dg0);
                    dg1 = fmaf(w1, d[36 + t], dg1);
                }
                rec[18] = dg0 * unscale;      // d(gamma_in) terms: sum_t w_t * sum_q g xhat_t
                rec[19] = dg1 * unscale;
                rec[20] = db0 * unscale;      // d(beta_in) terms: sum_t w_t * sum_q g valid_t
                rec[21] = db1 * unscale;
            }
#pragma unroll
            for (int i = 0; i < 22; ++i) o[i * Cs + c] = rec[i];
        }
        tc_fence_before();
    } else {
        const uint32_t idesc = make_idesc(128, 64, 1, 1);
        // MN-major SWIZZLE_128B descriptors: LBO = stride between 64-channel boxes, SBO = 8 rows x 128 B
        const uint32_t hi_word = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t lbo = (C0B_BOX >> 4) << 16;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            mbar_wait(b_full, it & 1u);
            tc_fence_after();
            if (umma::elect_one()) {
#pragma unroll 1
                for (uint32_t pos = 0; pos < 4; ++pos) {
                    const uint32_t gh = g_base + pos * g_pos, fh = f_base + pos * f_pos;
                    const uint64_t m_hi = ((uint64_t)hi_word << 32) | (lbo | ((gh & 0x3FFFFu) >> 4));
                    const uint64_t m_lo = ((uint64_t)hi_word << 32) | (lbo | (((gh + g_plane) & 0x3FFFFu) >> 4));
                    const uint64_t n_hi = ((uint64_t)hi_word << 32) | (lbo | ((fh & 0x3FFFFu) >> 4));
                    const uint64_t n_lo = ((uint64_t)hi_word << 32) | (lbo | (((fh + f_plane) & 0x3FFFFu) >> 4));
                    const uint32_t acc = (it > 0 || pos > 0) ? 1u : 0u;
                    if (p.planes == 2) umma_wgrad_x3(tmem_base, 64u, m_hi, m_lo, n_hi, n_lo, 0u, 0u, idesc, acc, 1);
                    else umma_wgrad_x1(tmem_base, 64u, m_hi, m_lo, n_hi, n_lo, 0u, 0u, idesc, acc, 1);
                }
                umma_commit(b_empty);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

}  // namespace

int conv0_tc_backward(int precision, const float* feat, int N, int H, int W, const float* scale, const float* shift,
                      const float* mean, const float* invstd, const float* w, const void* dzp, int dzp_half,
                      const unsigned char* amax, const unsigned* dz_absmax, const Geo& gp, float* dw, float* db,
                      float* dgamma_in, float* dbeta_in, void* scratch, cudaStream_t s) {
    FSB_REQUIRE(conv0_tc_supported(gp) && (precision == 1 || precision == 2) && amax, "conv0_tc_backward: unsupported");
    FSB_REQUIRE(!dzp_half || (precision == 2 && dz_absmax), "conv0_tc_backward: a half dzp needs the single-pass mode and its GradScale");
    Conv0TcBwdParams p;
    p.feat = feat; p.N = N; p.H = H; p.W = W;
    p.scale = scale; p.shift = shift; p.mean = mean; p.invstd = invstd; p.w = w;
    p.dzp = (const float*)dzp; p.dzp_half = dzp_half; p.amax = amax; p.absmax = dz_absmax; p.gp = gp;
    p.planes = precision == 1 ? 2 : 1;
    p.npix = (long long)N * gp.H * gp.W;
    p.ntiles = (int)((p.npix + C0B_PX - 1) / C0B_PX);
    p.partials = (float*)scratch;
    const size_t smem = 1024 + (size_t)4 * p.planes * (2 * C0B_BOX + C0B_BOX) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        FSB_CUDA(cudaFuncSetAttribute(conv0_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    int sms = 0, dev = 0;
    FSB_CUDA(cudaGetDevice(&dev));
    FSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int per_sm = p.planes == 1 ? 2 : 1;            // 97 KB (single pass) / 193 KB (three products) of operands per CTA
    int grid = sms * per_sm;
    if (grid > p.ntiles) grid = p.ntiles;
    if (grid > conv0_bwd_blocks()) grid = conv0_bwd_blocks();
    conv0_tc_bwd_kernel<<<grid, C0B_THREADS, smem, s>>>(p);
    FSB_LAUNCHED();
    return conv0_bwd_finalize((const float*)scratch, grid, gp, dw, db, dgamma_in, dbeta_in, s);
}

bool conv0_tc_supported(const Geo& gp) { return gp.Cs <= 128 && gp.Cs % 16 == 0 && gp.C * 18 < (1 << 30); }

int conv0_tc_forward(int precision, const float* feat, int N, int H, int W, const float* scale, const float* shift,
                     const float* w, const float* b, float* zp, unsigned char* amax, const Geo& gp, cudaStream_t s) {
    FSB_REQUIRE(gp.H == H / 2 && gp.W == W / 2 && gp.N == N, "conv0_tc: geometry mismatch");
    FSB_REQUIRE(conv0_tc_supported(gp) && (precision == 1 || precision == 2), "conv0_tc: unsupported shape / precision");
    Conv0TcParams p;
    p.feat = feat; p.N = N; p.H = H; p.W = W;
    p.scale = scale; p.shift = shift; p.w = w; p.b = b;
    p.zp = zp; p.amax = amax; p.gp = gp;
    p.BN = gp.Cs;
    p.planes = precision == 1 ? 2 : 1;
    p.npix = (long long)N * gp.H * gp.W;
    p.ntiles = (int)((p.npix + C0T_PX - 1) / C0T_PX);
    const size_t smem = 1024 + 2 * (size_t)C0T_ABUF + 2 * 128 * 64 + 64 + 128 * sizeof(float) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        FSB_CUDA(cudaFuncSetAttribute(conv0_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int sms = 0, dev = 0;
    FSB_CUDA(cudaGetDevice(&dev));
    FSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    conv0_tc_fwd_kernel<<<grid, C0T_THREADS, smem, s>>>(p);
    FSB_LAUNCHED();
    return 0;
}

}  // namespace fsb
