// tcgen05 / TMEM / TMA back end of the row-shifted GEMM (precision 1 = three-product split half, 2 = single pass).
//
// Every convolution of the network is a sum of row-shifted GEMMs over padded-flat NHWC tensors
// (gemm.cuh).  Activations arrive as two IEEE-half planes (hi, lo with x ~= hi + lo, common.cuh); weights are
// packed once per step into K-major half hi/lo tiles.  Products are formed on the 5th-generation tensor
// cores as   hi*hi + hi*lo + lo*hi   with float32 accumulation in TMEM (each product keeps ~2^-22
// relative error), or hi*hi only in precision 2 (2^-12 per operand; the backward GEMMs of the mixed mode).
//
//   conv_tc_kernel  (forward and dgrad)   D[r, n] = sum_t sum_k A[r + off_t, k] * W_t[n, k]
//       persistent, warp specialised: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (the whole warp walks
//       the schedule, one elected lane issues), warps 2..5 = epilogue.  ONE smem ring whose stage is a (dy, k-chunk):
//       the (128 + 8)-row activation boxes of the group's 1-2 row tiles (hi and lo planes; the three dx taps of a dy
//       share a box, the tap shift is a one-row start-address offset of the UMMA descriptor) plus the weight tiles
//       of those taps; 64-channel SWIZZLE_128B or 32-channel SWIZZLE_64B rows, whichever gives >= 2-3 stages.  All
//       MMAs of a stage are issued from one asm block (umma_issue.cuh).  Two accumulator sets in TMEM let the
//       epilogue of group i overlap the main loop of group i + 1.  Epilogue: TMEM -> registers in 128-column chunks
//       -> (+bias) swizzled smem boxes -> TMA stores (32-row x 32-column box per warp, double buffered); BatchNorm
//       column statistics are summed from the staged boxes, per warp, in a fixed order.
//   wgrad_tc_kernel                       dW_t[co, ci] = sum_r dZ[r, co] * A[r + off_t, ci]
//       both operands MN-major (channels contiguous, the reduction runs over pixel rows in 64-row stages); the M side
//       (128-channel tiles) is dZ or A, whichever wastes less padding; one CTA owns an (M tile, N tile <= 160
//       channels) pair for the three dx taps of one dy (3 accumulators) over a slice of the rows; partial sums go to
//       scratch and are reduced in a fixed order (deterministic).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <array>
#include <map>

#include "gemm.cuh"
#include "tc_ptx.cuh"
#include "umma_issue.cuh"

namespace fsb {

int wgrad_finalize(const float* P, int splits, const ConvGeom& c, float* dw, const unsigned* dz_absmax, cudaStream_t s);

namespace {

constexpr int BM = 128;            // UMMA M (cta_group::1)
constexpr int BK = 64;             // channels per smem row: 64 halves = 128 bytes = one swizzle span
constexpr int A_HALO = 8;          // extra rows of the A box (row shifts 0..2 used)
constexpr int TC_THREADS = 192;    // 6 warps
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int WG_R = 64;           // wgrad: pixel rows per pipeline stage

// ---------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ---------------------------------------------------------------------------------------------
// Pipeline stage = one (dy, k-chunk): the activation boxes of the group's T row tiles (hi + lo planes) and the weight
// tiles of the dx taps that share them, under ONE full / empty barrier pair; the MMA warp issues the whole stage from
// one asm block.  Two accumulator sets in TMEM let the epilogue of group i overlap the main loop of group i + 1.
struct ConvTcParams {
    long long rows;        // rows of the output (and of each activation plane)
    int m_tiles, n_tiles;
    int T;                 // row tiles per group: 1 or 2
    int nbuf;              // accumulator sets (2 whenever 2 * T * BN <= 512)
    int BN;                // output channels per tile (multiple of 16, <= 256)
    int K;                 // padded input channels CsIn (multiple of 16)
    int bk;                // channels per smem stage: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows)
    int ngroups, tpg;      // A boxes per k-chunk sweep, taps sharing one box
    int goff[9];           // row offset of the box origin relative to the tile's first row
    int wrow[9][3];        // first row of tap (group, shift) in the packed weight matrix (hi plane)
    int w_lo_row;          // row offset of the lo plane in the packed weight matrix
    int a_box_rows;        // 128 (tpg == 1) or 136
    int nstages;           // ring depth
    int planes;            // 2 = hi + lo planes (three products), 1 = hi only (single pass)
    int nstg;              // epilogue staging buffers per warp: 2 (double buffered) or 1 (smem is tight)
    int base_off_mode;     // 1: descriptor base_offset = row shift, 0: always 0
    float* Z;
    int ldz;               // CsOut
    const float* bias;     // CsOut entries or nullptr
    // optional fused BatchNorm statistics of Z over interior pixels: partials[blockIdx.x][2][ldz] (doubles)
    double* stats;
    const unsigned char* mask;   // interior mask of the output geometry (nullptr = every row is interior)
    const unsigned* out_scale;   // GradScale of the A operand (dgrad): the output is multiplied by its inverse; nullptr = 1
    // compact backward: the output is ONE half plane holding 2^k * D, k = gs_exponent2(out_scale, out_mul) (out_mul =
    // device float, the max column L1 norm of the weights: |D| <= max|dZ| * out_mul); tmZ32 / tmZ16 are half maps then
    int out_half;
    const float* out_mul;
    // eval forward with the following BatchNorm + PReLU folded into the epilogue (EPI 2): the output is the NEXT GEMM's
    // operand, a = prelu((acc + bias) * act_scale + act_shift), written as hi / lo half planes (tmZ32 / tmZ16 are 3-D half
    // maps then); act_slope may be nullptr (no activation); vectors have ldz entries except act_slope (act_c)
    const float* act_scale;
    const float* act_shift;
    const float* act_slope;
    int act_c;
};

constexpr int EPI_BOX_BYTES = 4096;             // one staged box: 32 rows x 128 bytes; 4 epilogue warps x nstg boxes
constexpr int MAX_T = 2;

// EPI 0: float32 output (+ bias, + BatchNorm statistics); 1: the compact-backward dgrad instantiation (one scaled half
// plane); 2: eval forward with BatchNorm + PReLU folded in (hi / lo operand planes of the next GEMM).  A template
// parameter so that the float32 epilogue of the training forward keeps its register allocation (a run-time branch cost
// the 1x1 layers 9 %)
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmZ32, const __grid_constant__ CUtensorMap tmZ16, const ConvTcParams p) {
    constexpr bool OUT_HALF = EPI == 1, OUT_ACT = EPI == 2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [stage ring: T x (hi, lo) activation boxes | tpg x (hi, lo) weight tiles][epilogue staging][barriers][bias][statistics]
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t row_bytes = (uint32_t)p.bk * 2u;                 // 128 or 64
    const uint32_t a_box = (uint32_t)p.a_box_rows * row_bytes;      // one plane of one tile
    const uint32_t a_tile = a_box * (uint32_t)p.planes;             // hi (+ lo)
    const uint32_t a_part = a_tile * (uint32_t)p.T;
    const uint32_t w_plane = (uint32_t)p.BN * row_bytes;
    const uint32_t w_tap = w_plane * (uint32_t)p.planes;
    const uint32_t stage_bytes = a_part + w_tap * (uint32_t)p.tpg;  // multiple of 1 KB
    const uint32_t ring = smem_base;
    const uint32_t epi_stage = (ring + stage_bytes * p.nstages + 1023u) & ~1023u;     // SWIZZLE_128B store boxes
    const uint32_t bars = epi_stage + 4u * p.nstg * EPI_BOX_BYTES;                    // 8-byte mbarriers
    // barrier layout: full[nstages] empty[nstages] T_full[2 MAX_T] T_empty[2 MAX_T], then tmem ptr
    const uint32_t b_full = bars, b_empty = b_full + 8u * p.nstages;
    const uint32_t bT_full = b_empty + 8u * p.nstages, bT_empty = bT_full + 16u * MAX_T;
    const uint32_t tmem_slot = bT_empty + 16u * MAX_T;
    unsigned char* tail = smem_raw + (bars - smem_u32(smem_raw)) + 256u;
    float* bias_s = reinterpret_cast<float*>(tail);                                              // [BN]
    double* stat_acc = reinterpret_cast<double*>(tail + (((size_t)p.BN * 4 + 15) & ~(size_t)15));  // [4 warps][2][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // schedule: a CTA keeps one column tile nt for its whole life and walks groups of T row tiles
    const int nt = blockIdx.x % p.n_tiles;
    const int grp0 = blockIdx.x / p.n_tiles, grp_step = gridDim.x / p.n_tiles;
    const int n_groups = (p.m_tiles + p.T - 1) / p.T;
    const int kchunks = (p.K + p.bk - 1) / p.bk;

    for (int i = threadIdx.x; i < p.BN; i += TC_THREADS) {
        const int cg = nt * p.BN + i;
        bias_s[i] = (p.bias && cg < p.ldz) ? p.bias[cg] : 0.f;
    }
    if (p.stats)
        for (int i = threadIdx.x; i < 8 * p.BN; i += TC_THREADS) stat_acc[i] = 0.0;
    // EPI 2: per-column (scale, shift incl. the conv bias, slope) in the statistics region (no statistics in eval)
    float* act_s = reinterpret_cast<float*>(stat_acc);                                          // [3][BN]
    if (OUT_ACT)
        for (int i = threadIdx.x; i < p.BN; i += TC_THREADS) {
            const int cg = nt * p.BN + i;
            const bool in = cg < p.ldz;
            const float sc = in ? p.act_scale[cg] : 0.f, sh = in ? p.act_shift[cg] : 0.f;
            const float b = (p.bias && in) ? p.bias[cg] : 0.f;
            act_s[i] = sc;
            act_s[p.BN + i] = fmaf(sc, b, sh);
            act_s[2 * p.BN + i] = (p.act_slope && cg < p.act_c) ? p.act_slope[cg] : 1.f;
        }

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nstages; ++i) { mbar_init(b_full + 8u * i, 1); mbar_init(b_empty + 8u * i, 1); }
        for (int i = 0; i < 2 * MAX_T; ++i) { mbar_init(bT_full + 8u * i, 1); mbar_init(bT_empty + 8u * i, 128); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmZ32);
        tma_prefetch_desc(&tmZ16);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int grp = grp0; grp < n_groups; grp += grp_step) {
                const int mt_first = grp * p.T;
                const int ntile = min(p.T, p.m_tiles - mt_first);
                for (int g = 0; g < p.ngroups; ++g) {
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(b_empty + 8u * st, ph ^ 1u);
                        const uint32_t full = b_full + 8u * st;
                        mbar_expect_tx(full, a_box * (uint32_t)(p.planes * ntile) + w_plane * (uint32_t)(p.planes * p.tpg));
                        const uint32_t base = ring + stage_bytes * st;
                        for (int t = 0; t < ntile; ++t) {
                            const uint32_t dst = base + a_tile * (uint32_t)t;
                            const int r0 = (mt_first + t) * BM + p.goff[g];
                            tma_load_3d(dst, &tmA, kc * p.bk, r0, 0, full);
                            if (p.planes == 2) tma_load_3d(dst + a_box, &tmA, kc * p.bk, r0, 1, full);
                        }
                        for (int s = 0; s < p.tpg; ++s) {
                            const uint32_t wd = base + a_part + w_tap * (uint32_t)s;
                            const int wr = p.wrow[g][s] + nt * p.BN;
                            tma_load_2d(wd, &tmW, kc * p.bk, wr, full);
                            if (p.planes == 2) tma_load_2d(wd + w_plane, &tmW, kc * p.bk, p.w_lo_row + wr, full);
                        }
                        if (++st == (uint32_t)p.nstages) { st = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the schedule and waits on the barriers (warp-uniform control flow); one elected lane
        // issues each stage's MMAs from a single asm block and commits.
        {
            const uint32_t idesc = make_idesc(BM, p.BN, 0, 0);
            const uint32_t layout = p.bk == 64 ? 2u : 4u;            // SWIZZLE_128B / SWIZZLE_64B
            const uint32_t sbo = 8u * row_bytes;                      // 8-row core-matrix groups
            // descriptor words: lo = start address >> 4 | LBO (16 B) << 16 ; hi = SBO >> 4 | version 1 << 14 | layout << 29
            const uint32_t d_lo = 1u << 16;
            const uint32_t d_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
            const int last_stage = p.ngroups * kchunks - 1;
            uint32_t st = 0, ph = 0;
            uint32_t it = 0;
            for (int grp = grp0; grp < n_groups; grp += grp_step, ++it) {
                const int ntile = min(p.T, p.m_tiles - grp * p.T);
                const uint32_t set = p.nbuf == 2 ? (it & 1u) : 0u, use = p.nbuf == 2 ? (it >> 1) : it;
                const uint32_t d_tmem0 = tmem_base + set * (uint32_t)(p.T * p.BN);
                // the epilogue must have drained this accumulator set (earlier group)
                for (int t = 0; t < ntile; ++t) mbar_wait(bT_empty + 8u * (set * MAX_T + (uint32_t)t), (use & 1u) ^ 1u);
                tc_fence_after();
                int kc = 0;
                for (int sidx = 0; sidx <= last_stage; ++sidx) {
                    mbar_wait(b_full + 8u * st, ph);
                    tc_fence_after();
                    const uint32_t base = ring + stage_bytes * st;
                    int ksteps = (p.K - kc * p.bk + 15) / 16;
                    if (ksteps > p.bk / 16) ksteps = p.bk / 16;
                    if (umma::elect_one()) {
                        const uint64_t a_hi = ((uint64_t)d_hi << 32) | (d_lo | ((base & 0x3FFFFu) >> 4));
                        const uint64_t b_hi = ((uint64_t)d_hi << 32) | (d_lo | (((base + a_part) & 0x3FFFFu) >> 4));
                        const bool ok = p.planes == 2
                            ? umma::umma_stage_x3(d_tmem0, (uint32_t)p.BN, a_hi, a_tile >> 4, a_box >> 4, row_bytes >> 4, b_hi,
                                                  w_tap >> 4, w_plane >> 4, idesc, sidx == 0 ? 0u : 1u, ksteps, ntile, p.tpg)
                            : umma::umma_stage_x1(d_tmem0, (uint32_t)p.BN, a_hi, a_tile >> 4, a_box >> 4, row_bytes >> 4, b_hi,
                                                  w_tap >> 4, w_plane >> 4, idesc, sidx == 0 ? 0u : 1u, ksteps, ntile, p.tpg);
                        if (!ok) __trap();               // stage shape without an issue block: host-side planning bug
                        umma_commit(b_empty + 8u * st);
                        if (sidx == last_stage)
                            for (int t = 0; t < ntile; ++t) umma_commit(bT_full + 8u * (set * MAX_T + (uint32_t)t));
                    }
                    __syncwarp();
                    if (++st == (uint32_t)p.nstages) { st = 0; ph ^= 1u; }
                    if (++kc == kchunks) kc = 0;
                }
            }
        }
    } else {
        // ====== epilogue: TMEM -> registers (128-column chunks) -> (+bias) swizzled smem boxes -> TMA stores;
        //        column statistics are summed from the staged boxes ======
        const int q = warp & 3;                    // TMEM lane quadrant this warp may read
        const int rl = q * 32 + lane;              // this thread's row within the tile
        const uint32_t stg0 = epi_stage + (uint32_t)(q * p.nstg) * EPI_BOX_BYTES;
        const unsigned char* stg0_g = smem_raw + (stg0 - smem_u32(smem_raw));
        double* my_acc = stat_acc + (size_t)q * 2 * p.BN;
        const float oscale = gs_inv_scale(p.out_scale) * (OUT_HALF ? gs_pow2(gs_exponent2(p.out_scale, p.out_mul)) : 1.f);
        // per-lane column statistics of this CTA's column tile: compensated float32 sums in registers (a DADD per panel
        // through shared memory was the top stall of the 1x1 layers); [128-column chunk][32-column panel][sum, sum of squares]
        float accS[2][4][2], accC[2][4][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) { (&accS[0][0][0])[i] = 0.f; (&accC[0][0][0])[i] = 0.f; }
        uint32_t sb = 0;                           // staging buffer toggle
        uint32_t it = 0;
        // interior flags of this thread's row in each tile of the NEXT group (global loads issued a group ahead)
        unsigned char inext[MAX_T];
        auto load_flags = [&](int grp) {
#pragma unroll
            for (int t = 0; t < MAX_T; ++t) {
                const long long row = ((long long)grp * p.T + t) * BM + rl;
                inext[t] = ((p.stats || OUT_ACT) && grp < n_groups && t < p.T && row < p.rows) ? (p.mask ? p.mask[row] : 1) : 0;
            }
        };
        load_flags(grp0);
        for (int grp = grp0; grp < n_groups; grp += grp_step, ++it) {
            const int mt_first = grp * p.T;
            const int ntile = min(p.T, p.m_tiles - mt_first);
            const uint32_t set = p.nbuf == 2 ? (it & 1u) : 0u, use = p.nbuf == 2 ? (it >> 1) : it;
            uint32_t ibits[MAX_T];
#pragma unroll
            for (int t = 0; t < MAX_T; ++t) ibits[t] = __ballot_sync(0xffffffffu, inext[t] != 0);
            load_flags(grp + grp_step);
#pragma unroll
            for (int t = 0; t < MAX_T; ++t) {
                if (t < ntile) {
                    const long long m0 = (long long)(mt_first + t) * BM;
                    const uint32_t slot = set * MAX_T + (uint32_t)t;
                    mbar_wait(bT_full + 8u * slot, use & 1u);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + (set * (uint32_t)p.T + (uint32_t)t) * (uint32_t)p.BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
                    for (int ci = 0; ci < 2; ++ci) {
                        const int c0 = ci * 128;
                        if (c0 >= p.BN) break;
                        // ---- this thread's row, columns [c0, c0 + width): one burst of TMEM loads, one wait
                        const int width = min(128, p.BN - c0);               // multiple of 16
                        float v[128];
                        {
                            const uint32_t ta = taddr + (uint32_t)c0;       // register positions are compile-time
                            if (width >= 64) {
                                tmem_ld64_async(ta, v);
                                const int r = width - 64;
                                if (r >= 64) tmem_ld64_async(ta + 64u, v + 64);
                                else if (r >= 32) { tmem_ld32_async(ta + 64u, v + 64); if (r >= 48) tmem_ld16_async(ta + 96u, v + 96); }
                                else if (r >= 16) tmem_ld16_async(ta + 64u, v + 64);
                            } else if (width >= 32) {
                                tmem_ld32_async(ta, v);
                                if (width >= 48) tmem_ld16_async(ta + 32u, v + 32);
                            } else {
                                tmem_ld16_async(ta, v);
                            }
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 128; ++i) pin(v[i]);
                        }
                        if (c0 + 128 >= p.BN) {            // accumulator drained: hand it back to the MMA warp at once
                            tc_fence_before();
                            mbar_arrive(bT_empty + 8u * slot);
                        }
#pragma unroll
                        for (int pn = 0; pn < 4; ++pn) {
                            const int cl0 = c0 + pn * 32;                           // first column within the CTA's column tile
                            if (pn * 32 < width) {
                                const bool wide = width - pn * 32 >= 32;           // 32 or 16 columns
                                // the staging buffer about to be overwritten was read by the TMA store issued nstg panels ago
                                if (lane == 0) {
                                    if (p.nstg == 2) bulk_wait_read<1>();
                                    else bulk_wait_read<0>();
                                }
                                __syncwarp();
                                const uint32_t stg = stg0 + sb * EPI_BOX_BYTES;
                                const float* stg_g = reinterpret_cast<const float*>(stg0_g + sb * EPI_BOX_BYTES);
                                if (OUT_ACT) {
                                    // a = prelu(acc * scale + shift') split into hi / lo halves: two staged boxes (hi at
                                    // the slot base, lo 2 KB above), rows as in the half-output case below
                                    // border rows of the padded-flat output are the zero padding of the next conv: they
                                    // must stay zero (the stand-alone BN-apply pass simply skips them)
                                    const uint32_t rbh = wide ? 64u : 32u;
                                    const bool has_sl = p.act_slope != nullptr;
                                    const bool interior = (ibits[t] >> lane) & 1u;
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        if (wide || j < 2) {
                                            const uint32_t ch = wide ? (uint32_t)((j ^ ((lane >> 1) & 3)) << 4) : (uint32_t)(j << 4);
                                            const float* vv = v + pn * 32 + 8 * j;
                                            const float* sc = act_s + cl0 + 8 * j;
                                            uint32_t wh[4], wl[4];
#pragma unroll
                                            for (int e = 0; e < 4; ++e) {
                                                float y0 = fmaf(vv[2 * e], sc[2 * e], sc[p.BN + 2 * e]);
                                                float y1 = fmaf(vv[2 * e + 1], sc[2 * e + 1], sc[p.BN + 2 * e + 1]);
                                                if (has_sl) {
                                                    y0 = y0 > 0.f ? y0 : sc[2 * p.BN + 2 * e] * y0;
                                                    y1 = y1 > 0.f ? y1 : sc[2 * p.BN + 2 * e + 1] * y1;
                                                }
                                                __half h0, l0, h1, l1;
                                                split_h16(interior ? y0 : 0.f, h0, l0);
                                                split_h16(interior ? y1 : 0.f, h1, l1);
                                                wh[e] = pack_h2(h0, h1);
                                                wl[e] = pack_h2(l0, l1);
                                            }
                                            st_shared_v4_u32(stg + (uint32_t)lane * rbh + ch, wh[0], wh[1], wh[2], wh[3]);
                                            st_shared_v4_u32(stg + 2048u + (uint32_t)lane * rbh + ch, wl[0], wl[1], wl[2], wl[3]);
                                        }
                                    }
                                } else if (OUT_HALF) {
                                    // half output: 64-byte staged rows in the SWIZZLE_64B pattern (16-byte chunk index xor
                                    // address bits 7..8), or plain 32-byte rows for a 16-column panel
                                    const uint32_t rbh = wide ? 64u : 32u;
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        if (wide || j < 2) {
                                            const uint32_t ch = wide ? (uint32_t)((j ^ ((lane >> 1) & 3)) << 4) : (uint32_t)(j << 4);
                                            const float* vv = v + pn * 32 + 8 * j;
                                            const float* bb = bias_s + cl0 + 8 * j;
                                            uint32_t w4[4];
#pragma unroll
                                            for (int e = 0; e < 4; ++e) {
                                                const __half2 h2 = __floats2half2_rn(fmaf(vv[2 * e], oscale, bb[2 * e]),
                                                                                     fmaf(vv[2 * e + 1], oscale, bb[2 * e + 1]));
                                                w4[e] = *reinterpret_cast<const uint32_t*>(&h2);
                                            }
                                            st_shared_v4_u32(stg + (uint32_t)lane * rbh + ch, w4[0], w4[1], w4[2], w4[3]);
                                        }
                                    }
                                } else {
                                const uint32_t rb = wide ? 128u : 64u;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    if (wide || j < 4) {
                                        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cl0 + 4 * j);
                                        const uint32_t ch = wide ? (uint32_t)((j ^ (lane & 7)) << 4) : (uint32_t)(j << 4);
                                        const float* vv = v + pn * 32 + 4 * j;
                                        st_shared_v4(stg + (uint32_t)lane * rb + ch, fmaf(vv[0], oscale, b4.x), fmaf(vv[1], oscale, b4.y),
                                                     fmaf(vv[2], oscale, b4.z), fmaf(vv[3], oscale, b4.w));
                                    }
                                }
                                }
                                fence_async_smem();
                                __syncwarp();
                                if (lane == 0 && m0 + q * 32 < p.rows) {
                                    if (OUT_ACT) {
                                        tma_store_3d(wide ? &tmZ32 : &tmZ16, stg, nt * p.BN + cl0, (int)m0 + q * 32, 0);
                                        tma_store_3d(wide ? &tmZ32 : &tmZ16, stg + 2048u, nt * p.BN + cl0, (int)m0 + q * 32, 1);
                                    } else {
                                        tma_store_2d(wide ? &tmZ32 : &tmZ16, stg, nt * p.BN + cl0, (int)m0 + q * 32);
                                    }
                                    bulk_commit();
                                }
                                if (EPI == 0 && p.stats && (wide || lane < 16)) {
                                    // lane l sums column cl0 + l over this warp's interior rows, straight from the staged box
                                    const uint32_t bits = ibits[t];
                                    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};     // four independent chains
                                    if (wide) {
#pragma unroll
                                        for (int r = 0; r < 32; ++r) {
                                            const float x = stg_g[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
                                            if ((bits >> r) & 1u) { s1[r & 3] += x; s2[r & 3] = fmaf(x, x, s2[r & 3]); }
                                        }
                                    } else {
#pragma unroll
                                        for (int r = 0; r < 32; ++r) {
                                            const float x = stg_g[r * 16 + lane];
                                            if ((bits >> r) & 1u) { s1[r & 3] += x; s2[r & 3] = fmaf(x, x, s2[r & 3]); }
                                        }
                                    }
                                    const float add[2] = {(s1[0] + s1[1]) + (s1[2] + s1[3]), (s2[0] + s2[1]) + (s2[2] + s2[3])};
#pragma unroll
                                    for (int h = 0; h < 2; ++h) {        // Neumaier-compensated accumulation across tiles
                                        const float a = accS[ci][pn][h], x = add[h];
                                        const float tsum = a + x;
                                        accC[ci][pn][h] += fabsf(a) >= fabsf(x) ? (a - tsum) + x : (x - tsum) + a;
                                        accS[ci][pn][h] = tsum;
                                    }
                                }
                                sb = (sb + 1u) & (uint32_t)(p.nstg - 1);
                            }
                        }
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_all();            // the staged boxes must be written before the CTA retires
        if (EPI == 0 && p.stats) {
#pragma unroll
            for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                for (int pn = 0; pn < 4; ++pn) {
                    const int cl = ci * 128 + pn * 32 + lane;
                    if (cl < p.BN) {
                        my_acc[cl] = (double)accS[ci][pn][0] + (double)accC[ci][pn][0];
                        my_acc[p.BN + cl] = (double)accS[ci][pn][1] + (double)accC[ci][pn][1];
                    }
                }
            // merge the four warps' partials in warp order (deterministic) and publish the CTA record: zero
            // outside this CTA's column tile
            asm volatile("bar.sync 1, 128;" ::: "memory");
            double* o = p.stats + (long long)blockIdx.x * 2 * p.ldz;
            const int e = threadIdx.x - 64;
            for (int i = e; i < 2 * p.ldz; i += 128) {
                const int h = i >= p.ldz ? 1 : 0;
                const int cl = i - h * p.ldz - nt * p.BN;
                double tot = 0.0;
                if (cl >= 0 && cl < p.BN) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) tot += stat_acc[(size_t)(w * 2 + h) * p.BN + cl];
                }
                o[i] = tot;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// wgrad kernel
// ---------------------------------------------------------------------------------------------
struct WgradTcParams {
    long long rows;
    long long rows_per_split;   // multiple of WG_R
    int splits;
    int swap;                   // 0: M = output channels (dZ), N = input channels (A);  1: M = input, N = output
    int m_tiles, n_tiles, bn;   // M tiles of 128 channels, N tiles of bn <= 160 channels (3 x bn TMEM columns)
    int ngroups, tpg;
    int goff[9];                // row offset of the A box origin per group
    int tap_of[9][3];           // torch tap index of (group, shift)
    int ntaps;
    int CsIn, CsOut;
    int nstages;
    int planes;
    float* P;                   // [split][tap][CsIn][CsOut]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmA, const WgradTcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t dz_box = WG_R * 128u;                  // one 64-channel box of dZ
    constexpr uint32_t a_box = (WG_R + A_HALO) * 128u;        // one 64-channel box of A (with halo rows)
    // channels of the two operands in this CTA's tile, in 64-channel boxes
    const int cz = p.swap ? p.bn : 128, ca = p.swap ? 128 : p.bn;
    const uint32_t nbz = (uint32_t)(cz + 63) / 64u, nba = (uint32_t)(ca + 63) / 64u;
    const uint32_t dz_plane = nbz * dz_box, a_plane = nba * a_box;
    const uint32_t stage_bytes = (dz_plane + a_plane) * (uint32_t)p.planes;   // [dZ hi (, lo)][A hi (, lo)]
    const uint32_t ring = smem_base;
    const uint32_t bars = ring + stage_bytes * p.nstages;
    const uint32_t b_full = bars, b_empty = b_full + 8u * p.nstages, b_done = b_empty + 8u * p.nstages;
    const uint32_t tmem_slot = b_done + 8u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item
    int item = blockIdx.x;
    const int sp = item % p.splits; item /= p.splits;
    const int ntile = item % p.n_tiles; item /= p.n_tiles;
    const int mtile = item % p.m_tiles; item /= p.m_tiles;
    const int g = item;
    const int m0 = mtile * 128, n0 = ntile * p.bn;
    const int co0 = p.swap ? n0 : m0, ci0 = p.swap ? m0 : n0;
    const long long rb = (long long)sp * p.rows_per_split;
    long long re = rb + p.rows_per_split;
    if (re > p.rows) re = p.rows;
    const int nchunks = re > rb ? (int)((re - rb + WG_R - 1) / WG_R) : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nstages; ++i) { mbar_init(b_full + 8u * i, 1); mbar_init(b_empty + 8u * i, 1); }
        mbar_init(b_done, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmDZ);
        tma_prefetch_desc(&tmA);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int ch = 0; ch < nchunks; ++ch) {
                const int r = (int)(rb + (long long)ch * WG_R);
                mbar_wait(b_empty + 8u * st, ph ^ 1u);
                mbar_expect_tx(b_full + 8u * st, (dz_plane + a_plane) * p.planes);
                const uint32_t base = ring + stage_bytes * st;
                for (int pl = 0; pl < p.planes; ++pl) {
                    const uint32_t dzd = base + pl * dz_plane;
                    const uint32_t ad = base + (uint32_t)p.planes * dz_plane + pl * a_plane;
                    for (uint32_t b = 0; b < nbz; ++b) tma_load_3d(dzd + b * dz_box, &tmDZ, co0 + 64 * (int)b, r, pl, b_full + 8u * st);
                    for (uint32_t b = 0; b < nba; ++b) tma_load_3d(ad + b * a_box, &tmA, ci0 + 64 * (int)b, r + p.goff[g], pl, b_full + 8u * st);
                }
                if (++st == (uint32_t)p.nstages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // whole warp walks the stages (uniform control flow), one elected lane issues each stage from one asm block
        const uint32_t idesc = make_idesc(128, p.bn, 1, 1);
        // MN-major SWIZZLE_128B descriptors: LBO = stride between 64-channel boxes, SBO = 8 rows x 128 B
        const uint32_t hi_word = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t lbo_z = (dz_box >> 4) << 16, lbo_a = (a_box >> 4) << 16;
        uint32_t st = 0, ph = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
            mbar_wait(b_full + 8u * st, ph);
            tc_fence_after();
            const uint32_t base = ring + stage_bytes * st;
            const uint32_t z_hi = base, z_lo = base + dz_plane;
            const uint32_t a_hi = base + (uint32_t)p.planes * dz_plane, a_lo = a_hi + a_plane;
            if (umma::elect_one()) {
                const uint64_t dzh = ((uint64_t)hi_word << 32) | (lbo_z | ((z_hi & 0x3FFFFu) >> 4));
                const uint64_t dzl = ((uint64_t)hi_word << 32) | (lbo_z | ((z_lo & 0x3FFFFu) >> 4));
                const uint64_t dah = ((uint64_t)hi_word << 32) | (lbo_a | ((a_hi & 0x3FFFFu) >> 4));
                const uint64_t dal = ((uint64_t)hi_word << 32) | (lbo_a | ((a_lo & 0x3FFFFu) >> 4));
                const uint32_t acc = ch > 0 ? 1u : 0u;
                if (p.planes == 2) {
                    if (p.swap) umma_wgrad_x3(tmem_base, (uint32_t)p.bn, dah, dal, dzh, dzl, 8u, 0u, idesc, acc, p.tpg);
                    else umma_wgrad_x3(tmem_base, (uint32_t)p.bn, dzh, dzl, dah, dal, 0u, 8u, idesc, acc, p.tpg);
                } else {
                    if (p.swap) umma_wgrad_x1(tmem_base, (uint32_t)p.bn, dah, dal, dzh, dzl, 8u, 0u, idesc, acc, p.tpg);
                    else umma_wgrad_x1(tmem_base, (uint32_t)p.bn, dzh, dzl, dah, dal, 0u, 8u, idesc, acc, p.tpg);
                }
                umma_commit(b_empty + 8u * st);
                if (ch == nchunks - 1) umma_commit(b_done);
            }
            __syncwarp();
            if (++st == (uint32_t)p.nstages) { st = 0; ph ^= 1u; }
        }
    } else {
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;                     // this thread's M channel
        const int m_limit = p.swap ? p.CsIn : p.CsOut, n_limit = p.swap ? p.CsOut : p.CsIn;
        if (nchunks > 0) {
            mbar_wait(b_done, 0);
            tc_fence_after();
        }
        for (int s = 0; s < p.tpg; ++s) {
            const int t = p.tap_of[g][s];
            float* Pt = p.P + ((long long)sp * p.ntaps + t) * p.CsIn * p.CsOut;
            const uint32_t taddr = tmem_base + (uint32_t)(s * p.bn) + ((uint32_t)(q * 32) << 16);
            for (int c = 0; c < p.bn; c += 16) {
                float v[16];
                if (nchunks > 0) {
                    tmem_ld16(taddr + (uint32_t)c, v);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (m < m_limit) {
                    if (p.swap) {            // m = input channel: 16 consecutive output channels are contiguous
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const int co = n0 + c + i;
                            if (co < n_limit)
                                *reinterpret_cast<float4*>(Pt + (long long)m * p.CsOut + co) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        }
                    } else {                 // m = output channel: consecutive lanes are contiguous
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int ci = n0 + c + i;
                            if (ci < n_limit) Pt[(long long)ci * p.CsOut + m] = v[i];
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// weight packing: torch (Cout, Cin, taps) float32 -> half hi/lo K-major tiles for forward and dgrad
// ---------------------------------------------------------------------------------------------
struct TcPackLayout {
    int bn_f, nt_f, npad_f, kpad_f;      // forward : N = Cout, K = Cin
    int bn_d, nt_d, npad_d, kpad_d;      // dgrad   : N = Cin,  K = Cout
    size_t fwd_elems, dgr_elems;         // elements per plane
    size_t off_fwd, off_dgr, off_bias, total;
};

void split_n(int cs, int& bn, int& nt) {
    nt = (cs + 255) / 256;
    bn = round_up((cs + nt - 1) / nt, 16);
}

TcPackLayout pack_layout(const ConvGeom& c) {
    TcPackLayout L;
    split_n(c.CsOut, L.bn_f, L.nt_f);
    split_n(c.CsIn, L.bn_d, L.nt_d);
    L.npad_f = L.bn_f * L.nt_f; L.kpad_f = round_up(c.CsIn, BK);
    L.npad_d = L.bn_d * L.nt_d; L.kpad_d = round_up(c.CsOut, BK);
    L.fwd_elems = (size_t)c.ntaps * L.npad_f * L.kpad_f;
    L.dgr_elems = (size_t)c.ntaps * L.npad_d * L.kpad_d;
    L.off_fwd = 0;
    L.off_dgr = align_up(L.off_fwd + 2 * L.fwd_elems * 2, 1024);
    L.off_bias = align_up(L.off_dgr + 2 * L.dgr_elems * 2, 1024);
    L.total = L.off_bias + (size_t)c.CsOut * 4;
    return L;
}

__global__ void tc_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, ConvGeom c, TcPackLayout L,
                               __half* fwd, __half* dgr, float* pb) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long i = i0; i < (long long)L.fwd_elems; i += stride) {
        int k = (int)(i % L.kpad_f);
        long long u = i / L.kpad_f;
        int n = (int)(u % L.npad_f);
        int t = (int)(u / L.npad_f);
        float v = (n < c.Cout && k < c.Cin) ? w[((long long)n * c.Cin + k) * c.ntaps + t] : 0.f;
        __half hi, lo;
        split_h16(v, hi, lo);
        fwd[i] = hi;
        fwd[L.fwd_elems + i] = lo;
    }
    for (long long i = i0; i < (long long)L.dgr_elems; i += stride) {
        int k = (int)(i % L.kpad_d);           // output channel
        long long u = i / L.kpad_d;
        int n = (int)(u % L.npad_d);           // input channel
        int t = (int)(u / L.npad_d);
        float v = (k < c.Cout && n < c.Cin) ? w[((long long)k * c.Cin + n) * c.ntaps + t] : 0.f;
        __half hi, lo;
        split_h16(v, hi, lo);
        dgr[i] = hi;
        dgr[L.dgr_elems + i] = lo;
    }
    for (long long i = i0; i < c.CsOut; i += stride) pb[i] = (bias && i < c.Cout) ? bias[i] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
    }
    return fn;
}

// activation planes: (C, rows, 2 planes) half, box (64, box_rows, 1)
int make_act_map(CUtensorMap* m, const void* base, long long rows, int Cs, int box_rows, int bk = BK) {
    auto fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return FSB_E_NODEVICE;
    }
    cuuint64_t dims[3] = {(cuuint64_t)Cs, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)Cs * 2, (cuuint64_t)rows * Cs * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(act rows=%lld Cs=%d box=%d) failed: %d", rows, Cs, box_rows, (int)r);
        return FSB_E_INVALID;
    }
    return 0;
}

// packed weights: (Kpad, total_rows) half, box (64, box_rows)
int make_w_map(CUtensorMap* m, const void* base, long long total_rows, int kpad, int box_rows, int bk = BK) {
    auto fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return FSB_E_NODEVICE;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights rows=%lld kpad=%d box=%d) failed: %d", total_rows, kpad, box_rows, (int)r);
        return FSB_E_INVALID;
    }
    return 0;
}

// float32 output matrix (ldz, rows), box (box_cols, 32): the epilogue's TMA-store target
int make_out_map(CUtensorMap* m, const float* base, long long rows, int ldz, int box_cols, bool swizzle) {
    auto fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return FSB_E_NODEVICE;
    }
    cuuint64_t dims[2] = {(cuuint64_t)ldz, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ldz * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(output rows=%lld ldz=%d box=%d) failed: %d", rows, ldz, box_cols, (int)r);
        return FSB_E_INVALID;
    }
    return 0;
}

// half output matrix (ldz, rows), box (box_cols, 32): the epilogue's TMA-store target of the compact backward
int make_out_map_h(CUtensorMap* m, const void* base, long long rows, int ldz, int box_cols, bool swizzle) {
    auto fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return FSB_E_NODEVICE;
    }
    cuuint64_t dims[2] = {(cuuint64_t)ldz, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ldz * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(half output rows=%lld ldz=%d box=%d) failed: %d", rows, ldz, box_cols, (int)r);
        return FSB_E_INVALID;
    }
    return 0;
}

// hi / lo half operand planes (Cs, rows, 2), box (box_cols, 32, 1): TMA-store target of the eval epilogue with the folded
// BatchNorm + PReLU
int make_out_map_act(CUtensorMap* m, const void* base, long long rows, int Cs, int box_cols, bool swizzle) {
    auto fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return FSB_E_NODEVICE;
    }
    cuuint64_t dims[3] = {(cuuint64_t)Cs, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)Cs * 2, (cuuint64_t)rows * Cs * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(act output rows=%lld Cs=%d box=%d) failed: %d", rows, Cs, box_cols, (int)r);
        return FSB_E_INVALID;
    }
    return 0;
}

// FSB200_* tuning switches, read once per process
struct TcEnv {
    int mode;      // FSB200_TC_MODE: bit 0 = one TMA box per tap (no dx sharing), bit 1 = descriptor base_offset = row
                   // shift (measured on B200: the swizzle is a function of the absolute smem address, so a
                   // 128-byte-shifted start address needs base_offset 0; base_offset = shift gives wrong results)
    int t1;        // FSB200_TC_T=1: one row tile per group
    int bk32;      // FSB200_TC_BK=32: 32-channel stages
    int wg_swap;   // FSB200_WG_SWAP: 1 / 2 force the wgrad operand orientation
};
const TcEnv& tc_env() {
    static TcEnv e = {-1, 0, 0, 0};
    if (e.mode < 0) {
        auto rd = [](const char* n) { const char* v = getenv(n); return v ? atoi(v) : 0; };
        e.t1 = rd("FSB200_TC_T") == 1;
        e.bk32 = rd("FSB200_TC_BK") == 32;
        e.wg_swap = rd("FSB200_WG_SWAP");
        e.mode = rd("FSB200_TC_MODE");
    }
    return e;
}
int tc_mode() { return tc_env().mode; }

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Groups taps that differ only in their column shift (consecutive row offsets base-1, base, base+1).
// sign = +1: forward offsets; -1: dgrad (negated).  Fills goff / tap index per (group, shift).
void group_taps(const ConvGeom& c, int sign, bool share, int& ngroups, int& tpg, int goff[9], int tap_of[9][3]) {
    const bool triple = share && (c.ntaps == 3 || c.ntaps == 9);
    if (!triple) {
        ngroups = c.ntaps;
        tpg = 1;
        for (int t = 0; t < c.ntaps; ++t) {
            goff[t] = sign * c.offs[t];
            tap_of[t][0] = t;
            tap_of[t][1] = tap_of[t][2] = 0;
        }
        return;
    }
    ngroups = c.ntaps / 3;
    tpg = 3;
    for (int g = 0; g < ngroups; ++g) {
        // taps 3g, 3g+1, 3g+2 have offsets o-1, o, o+1 (dx = -1, 0, +1)
        const int centre = sign * c.offs[3 * g + 1];
        goff[g] = centre - 1;
        for (int s = 0; s < 3; ++s) tap_of[g][s] = sign > 0 ? 3 * g + s : 3 * g + (2 - s);
    }
}

// One planned launch (kernel parameters + the four tensor maps); planning costs a few microseconds of host time per
// launch (four cuTensorMapEncodeTiled calls), so plans are cached per (operand pointers, shape): the workspace of a
// bound network is stable across steps and every step replays the same ~60 entries.
struct ConvLaunch {
    ConvTcParams p;
    CUtensorMap tmA, tmW, tmZ32, tmZ16;
    int grid;
    size_t smem;
};
typedef std::array<uint64_t, 16> LaunchKey;

template <typename V>
struct LaunchCache {
    std::map<LaunchKey, V> map;
    V* find(const LaunchKey& k) {
        auto it = map.find(k);
        return it == map.end() ? nullptr : &it->second;
    }
    V* insert(const LaunchKey& k, const V& v) {
        if (map.size() > 8192) map.clear();          // variable-length batches: bounded growth
        return &map.emplace(k, v).first->second;
    }
};
LaunchCache<ConvLaunch> g_conv_cache;

int plan_conv_tc(ConvLaunch& L, int precision, const void* A, const void* wpacked, int w_kpad, int w_npad, int bn, int nt,
                 const float* bias, float* Z, long long rows, int K, int ldz, const ConvGeom& c, int sign,
                 const FwdStats* st, const unsigned* out_scale, const float* out_half_mul, const FwdAct* act) {
    ConvTcParams& p = L.p;
    memset(&p, 0, sizeof(p));
    p.rows = rows;
    p.m_tiles = (int)((rows + BM - 1) / BM);
    p.n_tiles = nt;
    p.BN = bn;
    p.K = K;
    p.w_lo_row = c.ntaps * w_npad;
    p.planes = precision == 1 ? 2 : 1;
    p.base_off_mode = (tc_mode() & 2) ? 1 : 0;
    p.Z = Z;
    p.ldz = ldz;
    p.bias = bias;
    p.out_scale = out_scale;
    p.out_half = out_half_mul ? 1 : 0;
    p.out_mul = out_half_mul;
    FSB_REQUIRE(!(st && out_half_mul), "conv_tc: fused statistics need the float32 output");
    FSB_REQUIRE(!(act && (st || out_half_mul)), "conv_tc: the folded BatchNorm + PReLU epilogue is an eval-forward option");
    if (act) {
        p.act_scale = act->scale; p.act_shift = act->shift; p.act_slope = act->slope; p.act_c = act->C;
        p.mask = act->mask;
    }
    if (st) {
        const Geo& g = *st->g;
        FSB_REQUIRE(g.rows == rows && g.Cs == ldz, "conv_tc: statistics geometry does not match the output");
        p.stats = st->partials;
        p.mask = g.mask;
    }
    FSB_REQUIRE(p.n_tiles <= num_sms() && bn <= 256 && bn % 16 == 0, "conv_tc: bad column tiling (%d x %d)", p.n_tiles, bn);
    // Two row tiles per group when two accumulator sets of two tiles fit in the 512 TMEM columns (BN <= 128), else
    // one; never more tiles per group than needed to give every CTA a group.
    const int ctas_per_col = num_sms() / p.n_tiles;
    int T = 4 * bn <= 512 ? 2 : 1;
    if (tc_env().t1) T = 1;
    while (T > 1 && (p.m_tiles + T - 1) / T < ctas_per_col) --T;
    const int nbuf = 2 * T * bn <= 512 ? 2 : 1;
    const size_t fixed0 = 1024 /*align*/ + 1024 /*staging align*/ + 256 /*barriers*/ +
                          (((size_t)bn * 4 + 15) & ~(size_t)15) /*bias*/ + (size_t)64 * bn /*statistics*/;
    // Stage shape.  Preferred: the three dx taps of a dy share one (128 + 8)-row activation box (a third of the
    // activation fills); wide tiles whose three weight tiles do not fit beside it fall back to one box per tap.
    // Stage width: 64 channels (SWIZZLE_128B) when three stages fit, else 32 channels (SWIZZLE_64B, at least two).
    size_t stage = 0, fixed = 0;
    bool ok = false;
    for (int share = (tc_mode() & 1) ? 0 : 1; share >= 0 && !ok; --share) {
        int tap_of[9][3];
        group_taps(c, sign, share != 0, p.ngroups, p.tpg, p.goff, tap_of);
        for (int g = 0; g < p.ngroups; ++g)
            for (int t = 0; t < p.tpg; ++t) p.wrow[g][t] = tap_of[g][t] * w_npad;
        p.a_box_rows = p.tpg == 1 ? BM : BM + A_HALO;
        // double-buffered store staging first (with a single box per warp every panel waits for its TMA store to
        // drain: the 1x1 layers ran at half speed), then the wider stage
        for (int nstg = 2; nstg >= 1 && !ok; --nstg) {
            fixed = fixed0 + (size_t)4 * nstg * EPI_BOX_BYTES;
            for (int bk = (tc_env().bk32 ? 32 : 64); bk >= 32 && !ok; bk -= 32) {
                stage = ((size_t)p.a_box_rows * T + (size_t)bn * p.tpg) * bk * 2 * p.planes;
                const int need = bk == 64 ? 3 : 2;
                if (fixed + need * stage <= SMEM_LIMIT) { p.bk = bk; p.nstg = nstg; ok = true; }
            }
        }
    }
    FSB_REQUIRE(ok, "conv_tc: tile does not fit in shared memory (BN=%d)", bn);
    p.T = T;
    p.nbuf = nbuf;
    p.nstages = (int)((SMEM_LIMIT - fixed) / stage);
    if (p.nstages > 8) p.nstages = 8;
    L.smem = fixed + stage * p.nstages;
    FSB_TRY(make_act_map(&L.tmA, A, rows, K, p.a_box_rows, p.bk));
    FSB_TRY(make_w_map(&L.tmW, wpacked, (long long)2 * c.ntaps * w_npad, w_kpad, bn, p.bk));
    if (act) {
        FSB_TRY(make_out_map_act(&L.tmZ32, Z, rows, ldz, 32, true));
        FSB_TRY(make_out_map_act(&L.tmZ16, Z, rows, ldz, 16, false));
    } else if (p.out_half) {
        FSB_TRY(make_out_map_h(&L.tmZ32, Z, rows, ldz, 32, true));
        FSB_TRY(make_out_map_h(&L.tmZ16, Z, rows, ldz, 16, false));
    } else {
        FSB_TRY(make_out_map(&L.tmZ32, Z, rows, ldz, 32, true));
        FSB_TRY(make_out_map(&L.tmZ16, Z, rows, ldz, 16, false));
    }
    // every CTA owns one column tile: the grid is a multiple of n_tiles
    const int n_groups = (p.m_tiles + T - 1) / T;
    L.grid = (n_groups < ctas_per_col ? n_groups : ctas_per_col) * p.n_tiles;
    return 0;
}

int launch_conv_tc(int precision, const void* A, const void* wpacked, int w_kpad, int w_npad, int bn, int nt,
                   const float* bias, float* Z, long long rows, int K, int ldz, const ConvGeom& c, int sign,
                   const FwdStats* st, const unsigned* out_scale, const float* out_half_mul, const FwdAct* act,
                   cudaStream_t s) {
    const LaunchKey key = {(uint64_t)(uintptr_t)A, (uint64_t)(uintptr_t)wpacked, (uint64_t)(uintptr_t)Z, (uint64_t)rows,
                           (uint64_t)(uintptr_t)bias, (uint64_t)(uintptr_t)(st ? st->partials : nullptr),
                           (uint64_t)(uintptr_t)(st ? st->g->mask : nullptr), (uint64_t)(uintptr_t)out_scale,
                           ((uint64_t)(uint32_t)K << 32) | (uint32_t)ldz, ((uint64_t)(uint32_t)bn << 32) | (uint32_t)nt,
                           ((uint64_t)(uint32_t)w_kpad << 32) | (uint32_t)w_npad,
                           ((uint64_t)(uint32_t)precision << 32) | (uint32_t)(sign + 1),
                           ((uint64_t)(uint32_t)c.ntaps << 32) | (uint32_t)c.offs[0], (uint64_t)(uint32_t)c.offs[1],
                           (uint64_t)(uintptr_t)out_half_mul,
                           act ? ((uint64_t)(uintptr_t)act->scale ^ ((uint64_t)(uintptr_t)act->slope << 1) ^ ((uint64_t)act->C << 48) ^ 1u) : 0};
    ConvLaunch* L = g_conv_cache.find(key);
    if (!L) {
        ConvLaunch fresh;
        FSB_TRY(plan_conv_tc(fresh, precision, A, wpacked, w_kpad, w_npad, bn, nt, bias, Z, rows, K, ldz, c, sign, st, out_scale, out_half_mul, act));
        L = g_conv_cache.insert(key, fresh);
    }
    static bool attr_set = false;
    if (!attr_set) {
        FSB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        FSB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        FSB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_set = true;
    }
    if (L->p.act_scale) conv_tc_kernel<2><<<L->grid, TC_THREADS, L->smem, s>>>(L->tmA, L->tmW, L->tmZ32, L->tmZ16, L->p);
    else if (L->p.out_half) conv_tc_kernel<1><<<L->grid, TC_THREADS, L->smem, s>>>(L->tmA, L->tmW, L->tmZ32, L->tmZ16, L->p);
    else conv_tc_kernel<0><<<L->grid, TC_THREADS, L->smem, s>>>(L->tmA, L->tmW, L->tmZ32, L->tmZ16, L->p);
    FSB_LAUNCHED();
    if (st) *st->nblk = L->grid;
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// back-end entry points
// ---------------------------------------------------------------------------------------------
size_t tc_packed_weight_bytes(const ConvGeom& c) { return pack_layout(c).total + 1024; }

static char* tc_pack_base(const void* packed) { return (char*)align_up((size_t)packed, 1024); }

int tc_pack_weights(const float* w, const float* bias, const ConvGeom& c, void* packed, cudaStream_t s) {
    TcPackLayout L = pack_layout(c);
    char* base = tc_pack_base(packed);
    size_t total = L.fwd_elems > L.dgr_elems ? L.fwd_elems : L.dgr_elems;
    int blocks = (int)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
    tc_pack_kernel<<<blocks, 256, 0, s>>>(w, bias, c, L, (__half*)(base + L.off_fwd), (__half*)(base + L.off_dgr),
                                          (float*)(base + L.off_bias));
    FSB_LAUNCHED();
    return 0;
}

int tc_max_ctas() { return num_sms(); }

int tc_fwd(int precision, const void* A, const void* packed, float* Z, const ConvGeom& c, const FwdStats* st,
           cudaStream_t s) {
    TcPackLayout L = pack_layout(c);
    const char* base = tc_pack_base(packed);
    return launch_conv_tc(precision, A, base + L.off_fwd, L.kpad_f, L.npad_f, L.bn_f, L.nt_f,
                          (const float*)(base + L.off_bias), Z, c.rows, c.CsIn, c.CsOut, c, +1, st, nullptr, nullptr, nullptr, s);
}

int tc_fwd_act(int precision, const void* A, const void* packed, void* a_out, const ConvGeom& c, const FwdAct& act,
               cudaStream_t s) {
    TcPackLayout L = pack_layout(c);
    const char* base = tc_pack_base(packed);
    FSB_REQUIRE(act.scale && act.shift, "tc_fwd_act: BatchNorm scale / shift required");
    return launch_conv_tc(precision, A, base + L.off_fwd, L.kpad_f, L.npad_f, L.bn_f, L.nt_f,
                          (const float*)(base + L.off_bias), (float*)a_out, c.rows, c.CsIn, c.CsOut, c, +1, nullptr, nullptr,
                          nullptr, &act, s);
}

int tc_dgrad(int precision, const void* dZ, const void* packed, void* dA, const ConvGeom& c, const unsigned* dz_absmax,
             const float* out_half_mul, cudaStream_t s) {
    TcPackLayout L = pack_layout(c);
    const char* base = tc_pack_base(packed);
    FSB_REQUIRE(!out_half_mul || dz_absmax, "tc_dgrad: the half output needs the GradScale of dZ");
    return launch_conv_tc(precision, dZ, base + L.off_dgr, L.kpad_d, L.npad_d, L.bn_d, L.nt_d, nullptr, (float*)dA, c.rows,
                          c.CsOut, c.CsIn, c, -1, nullptr, dz_absmax, out_half_mul, nullptr, s);
}

struct WgradShape {
    int ngroups, tpg, swap, m_tiles, n_tiles, bn, splits;
    long long rows_per_split;
};

// N tiling of `cs` channels: equal tiles of at most 160 channels (3 accumulators x bn <= 512 TMEM columns)
static void wgrad_n_tiles(int cs, int& nt, int& bn) {
    nt = (cs + 159) / 160;
    bn = round_up((cs + nt - 1) / nt, 16);
}

static WgradShape wgrad_shape(const ConvGeom& c) {
    WgradShape w;
    int goff[9], tap_of[9][3];
    group_taps(c, +1, !(tc_mode() & 1), w.ngroups, w.tpg, goff, tap_of);
    // orientation: the M side is padded to 128-channel tiles, the N side to 16; an MMA costs max(N/2, 44) cycles
    long long cost[2];
    int mt[2], nt[2], bn[2];
    for (int sw = 0; sw < 2; ++sw) {
        const int cm = sw ? c.CsIn : c.CsOut, cn = sw ? c.CsOut : c.CsIn;
        mt[sw] = (cm + 127) / 128;
        wgrad_n_tiles(cn, nt[sw], bn[sw]);
        cost[sw] = (long long)mt[sw] * nt[sw] * (bn[sw] > 88 ? bn[sw] : 88);
    }
    w.swap = cost[1] < cost[0] ? 1 : 0;
    if (tc_env().wg_swap == 1) w.swap = 0;
    if (tc_env().wg_swap == 2) w.swap = 1;
    w.m_tiles = mt[w.swap]; w.n_tiles = nt[w.swap]; w.bn = bn[w.swap];
    const int items = w.ngroups * w.m_tiles * w.n_tiles;
    long long chunks = (c.rows + WG_R - 1) / WG_R;
    int want = (2 * num_sms()) / items;                 // at most two full waves (one CTA per SM at a time)
    if (want > chunks) want = (int)chunks;
    if (want > 128) want = 128;
    if (want < 1) want = 1;
    w.splits = want;
    w.rows_per_split = (chunks + w.splits - 1) / w.splits * WG_R;
    return w;
}

size_t tc_wgrad_scratch_bytes(const ConvGeom& c) {
    return (size_t)wgrad_shape(c).splits * c.ntaps * c.CsIn * c.CsOut * sizeof(float);
}

namespace {
struct WgradLaunch {
    WgradTcParams p;
    CUtensorMap tmDZ, tmA;
    int grid;
    size_t smem;
};
LaunchCache<WgradLaunch> g_wgrad_cache;

int plan_wgrad_tc(WgradLaunch& L, int precision, const void* A, const void* dZ, void* scratch, const ConvGeom& c) {
    WgradTcParams& p = L.p;
    memset(&p, 0, sizeof(p));
    const WgradShape w = wgrad_shape(c);
    group_taps(c, +1, !(tc_mode() & 1), p.ngroups, p.tpg, p.goff, p.tap_of);
    p.swap = w.swap; p.m_tiles = w.m_tiles; p.n_tiles = w.n_tiles; p.bn = w.bn;
    p.splits = w.splits; p.rows_per_split = w.rows_per_split;
    p.rows = c.rows;
    p.ntaps = c.ntaps;
    p.CsIn = c.CsIn;
    p.CsOut = c.CsOut;
    p.planes = precision == 1 ? 2 : 1;
    p.P = (float*)scratch;
    const int cz = p.swap ? p.bn : 128, ca = p.swap ? 128 : p.bn;
    const size_t stage = (size_t)(((cz + 63) / 64) * WG_R * 128 + ((ca + 63) / 64) * (WG_R + A_HALO) * 128) * p.planes;
    p.nstages = (int)((SMEM_LIMIT - 1024 - 256) / stage);
    if (p.nstages > 6) p.nstages = 6;
    FSB_REQUIRE(p.nstages >= 2 && p.tpg * p.bn <= 512, "wgrad_tc: tile does not fit (bn=%d)", p.bn);
    L.smem = 1024 + 256 + stage * p.nstages;
    FSB_TRY(make_act_map(&L.tmDZ, dZ, c.rows, c.CsOut, WG_R));
    FSB_TRY(make_act_map(&L.tmA, A, c.rows, c.CsIn, WG_R + A_HALO));
    L.grid = p.ngroups * p.m_tiles * p.n_tiles * p.splits;
    return 0;
}
}  // namespace

int tc_wgrad(int precision, const void* A, const void* dZ, float* dw, void* scratch, const ConvGeom& c,
             const unsigned* dz_absmax, cudaStream_t s) {
    const LaunchKey key = {(uint64_t)(uintptr_t)A, (uint64_t)(uintptr_t)dZ, (uint64_t)(uintptr_t)scratch, (uint64_t)c.rows,
                           ((uint64_t)(uint32_t)c.CsIn << 32) | (uint32_t)c.CsOut,
                           ((uint64_t)(uint32_t)c.ntaps << 32) | (uint32_t)c.offs[0], (uint64_t)(uint32_t)c.offs[1],
                           (uint64_t)(uint32_t)precision, 0, 0, 0, 0, 0, 0, 0, 0};
    WgradLaunch* L = g_wgrad_cache.find(key);
    if (!L) {
        WgradLaunch fresh;
        FSB_TRY(plan_wgrad_tc(fresh, precision, A, dZ, scratch, c));
        L = g_wgrad_cache.insert(key, fresh);
    }
    static bool attr_set = false;
    if (!attr_set) {
        FSB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_set = true;
    }
    wgrad_tc_kernel<<<L->grid, TC_THREADS, L->smem, s>>>(L->tmDZ, L->tmA, L->p);
    FSB_LAUNCHED();
    return wgrad_finalize((const float*)scratch, L->p.splits, c, dw, dz_absmax, s);
}

}  // namespace fsb
