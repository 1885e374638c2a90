// Whole-network executor ("plan"): feature kernel -> conv blocks -> global-max heads -> FC head,
// forward and backward, for TwoDimensionalCNNClassificationModel (networks/classifiers.py:483-607)
// and HierarchicalCNNClassificationModel (:107-217).  The host side is a flat C++ schedule of kernel
// launches on the caller's stream; every tensor lives in the caller-provided workspace.
//
// Parameter order ("canonical order", == named_parameters() of the reference module tree), per block k:
//   0 bn_in.w 1 bn_in.b 2 conv.w 3 conv.b 4 bn_a.w 5 bn_a.b 6 prelu_a.w
//   7 res.conv1.w 8 res.conv1.b 9 res.bn1.w 10 res.bn1.b 11 res.conv2.w 12 res.conv2.b 13 res.bn2.w 14 res.bn2.b
//   15 res.conv3.w 16 res.conv3.b 17 res.bn3.w 18 res.bn3.b 19 res.prelu1.w 20 res.prelu2.w 21 res.prelu3.w
// then the head: bn0.w bn0.b lin1.w lin1.b bn2.w bn2.b prelu.w lin5.w lin5.b
// BN buffer order: per block bn_in, bn_a, bn1, bn2, bn3 ; head bn0, bn2.
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv0.cuh"
#include "eltwise.cuh"
#include "gemm.cuh"
#include "rnn.cuh"

using namespace fsb;

namespace {

enum { P_BNIN_W = 0, P_BNIN_B, P_CONV_W, P_CONV_B, P_BNA_W, P_BNA_B, P_PRELUA, P_C1_W, P_C1_B, P_BN1_W, P_BN1_B,
       P_C2_W, P_C2_B, P_BN2_W, P_BN2_B, P_C3_W, P_C3_B, P_BN3_W, P_BN3_B, P_PRELU1, P_PRELU2, P_PRELU3,
       P_PER_BLOCK };
enum { H_BN0_W = 0, H_BN0_B, H_L1_W, H_L1_B, H_BN2_W, H_BN2_B, H_PRELU, H_L5_W, H_L5_B, H_COUNT };
enum { B_IN = 0, B_A, B_1, B_2, B_3, B_PER_BLOCK };
enum { R_PER_HEAD = 10 };      // rnn head parameters: ln.w ln.b, then (w_ih w_hh b_ih b_hh) x {forward, reverse}
enum { RNN_SIZE = 128 };

enum Cat { CAT_FEAT = 0, CAT_GEMM_FWD, CAT_GEMM_DGRAD, CAT_GEMM_WGRAD, CAT_CONV0, CAT_ELT_FWD, CAT_ELT_BWD,
           CAT_HEAD, CAT_PACK, CAT_COUNT };
const char* kCatNames[CAT_COUNT] = {"feat", "gemm_fwd", "gemm_dgrad", "gemm_wgrad", "conv0", "eltwise_fwd",
                                    "eltwise_bwd", "head", "pack"};

struct BnBuf {
    float *scale, *shift, *mean, *invstd, *c1, *c2;
    int C, Cs;
    BnCoef coef(const float* slope) const { return BnCoef{scale, shift, slope, mean, invstd}; }
};

struct Bump {
    char* base;
    size_t off;
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
    void* take_bytes(size_t bytes) { return take<char>(bytes); }
};

struct BlockPlan {
    int Cin, C;
    Geo g_in, g_full, g;
    ConvGeom entry, c1, c2, c3;
    BnBuf bn_in, bn_a, bn1, bn2, bn3;
    void *pk_entry, *pk1, *pk2, *pk3;
    float *x_in;         // float32 PF block input (1D block 0: features; k>0: previous out)
    void *u;             // BN_in output, GEMM format (unused for 2D block 0)
    float *zf, *zp, *z1, *z2, *z3, *out;
    void *r0, *a1, *a2;
    int* argrow;
    void* gmax_scratch;
    unsigned char *mask, *mask_in;
    int head_off;        // column offset in the concatenated head input, -1 if no head
    int rnn_index;       // aggregation_type == "rnn": index of this block's head among the rnn heads, else -1
    RnnHead rnn;
    void* pk_rnn[2];     // packed W_ih / b_ih per direction
    unsigned char* pool_amax;   // training, blocks with an entry GEMM: arg-max position of every pool window
    unsigned char* sign3;       // training: PReLU-branch bytes of the block output (compact bn3 backward), rows * Cs / 8
    // backward (compact mode: da2 / da1 / dr0a / dr0b / dzp / du are scaled half planes, common.cuh GradRef)
    float* d_out;
    void *da2, *da1, *dr0a, *dr0b, *dzp, *du;
    void *dz3, *dz2, *dz1, *dzf;
};

}  // namespace

struct fsb_net {
    fsb_net_config cfg;
    std::vector<float> fb_vals;
    std::vector<int> fb_off, fb_start, fb_len;
    int D, Ds, CsCls;
    int n_rnn = 0;              // rnn heads (0 for aggregation_type == "max")
    int head_param_base = 0;    // index of the first FC-head parameter (after the blocks and the rnn heads)
    std::vector<long long> param_numel, param_offset;
    long long total_params;

    // binding state
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int N = 0, T = 0, training = 0, frames = 0;
    bool tables_ready = false, fwd_done = false;
    bool overlap = true;        // side-stream overlap of weight packing / weight-gradient GEMMs (fsb_net_set_overlap)
    bool conv0_tc = true;       // block-0 entry conv on the tensor cores (FSB200_CONV0_TC=0: CUDA-core direct conv)
    // Compact backward (mixed mode; FSB200_COMPACT_BWD=0 turns it off): dgrad outputs, the residual-branch gradient and
    // dzp are single scaled half planes, BatchNorm-backward reads the hi plane of the stored activation instead of the
    // float32 pre-activation, max-pool backward routes by stored arg-max bytes.
    bool compact = false;
    int fuse_eval = 14;         // eval forward folds (FSB200_FUSE_EVAL): bit 0 / 1: BN1 / BN2 + PReLU in the epilogue of conv1 /
                                // conv2; bit 2: the next block's BN_in in the block-output kernel; bit 3: BN_a + PReLU_a in
                                // the pooling kernel.  Default 14: everything but conv1 (measured best)
    float* wl1 = nullptr;       // [num_blocks][4] max column L1 norm of the entry / conv1 / conv2 / conv3 weights
    // CUDA graphs: the launch sequence of a forward (or backward) call with a given set of pointers / shapes is captured
    // on its second occurrence and replayed afterwards (~130 launches become one cudaGraphLaunch).
    bool graphs = true;         // FSB200_GRAPHS=0 / fsb_net_set_graphs(net, 0): always launch eagerly
    struct GraphEntry { unsigned long long key; cudaGraphExec_t exec; long long launches; unsigned long long stamp; };
    std::vector<GraphEntry> graph_cache;
    std::vector<unsigned long long> seen_keys;      // keys met once (captured on the next occurrence)
    unsigned long long graph_clock = 0;
    unsigned long long* d_seed = nullptr;           // device copy of the dropout seed (kernels read it through a pointer)
    // Stream capture is not allowed on the legacy default stream (torch's default current stream), so with graphs enabled
    // the plan executes on its own non-blocking stream, forked from / joined to the caller's stream with events: the
    // caller still sees plain stream semantics.
    cudaStream_t exec_stream = nullptr;
    cudaEvent_t exec_in = nullptr, exec_out = nullptr;
    unsigned long long dropout_seed = 0;

    // carved buffers
    std::vector<BlockPlan> blocks;
    void* feat_tables;
    float* d_fb_vals;
    int *d_fb_off, *d_fb_start, *d_fb_len;
    float* feat;          // 2D: (N, F, frames) plain
    double* partials;     // shared reduction scratch
    void* wgrad_scratch;
    void* rnn_scratch = nullptr;    // split-reduction scratch of the rnn heads' weight gradients (main stream; the conv
                                    // weight gradients own wgrad_scratch on the side stream)
    float *feats, *h0, *z1h, *h1, *zl, *dzl, *dh1, *dz1h, *dh0, *dfeats;
    BnBuf hbn0, hbn2;
    ConvGeom lin1, lin5;
    void *pk_l1, *pk_l5, *head_scratch;
    Geo g_head, g_cls;

    // backward side stream: weight-gradient GEMMs (tensor pipe) overlap the BatchNorm-backward passes (HBM) of the
    // critical path.  Fork / join with events, so the caller still sees plain stream semantics.
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> fork_events;
    size_t fork_used = 0;
    cudaEvent_t join_event = nullptr, pack_event = nullptr;
    void* conv0_scratch = nullptr;
    unsigned char* conv0_amax = nullptr;   // 2D block 0: arg-max position of every pool window (training)
    unsigned* gscale = nullptr;     // [num_blocks][B_PER_BLOCK] GradScale slots (common.cuh), zeroed at the start of backward
    unsigned* gscale_out = nullptr; // [num_blocks + 1]: GradScale of the half d_out planes (compact backward); last: max |dfeats|

    // precision of the forward / backward GEMMs (cfg.precision 3 = mixed: three-product forward, single-pass backward)
    int prec_f = 0, prec_b = 0;

    // profiling
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    struct Rec { int cat; int e0, e1; double flops; };
    std::vector<Rec> recs;
    size_t ev_used = 0;
};

namespace {

long long conv_numel(int cout, int cin, int taps) { return (long long)cout * cin * taps; }

void build_param_table(fsb_net* net) {
    const fsb_net_config& c = net->cfg;
    int taps = c.two_d ? 9 : 3;
    net->param_numel.clear();
    for (int k = 0; k < c.num_blocks; ++k) {
        int cin = k == 0 ? (c.two_d ? 2 : c.n_features) : c.depth[k - 1];
        int d = c.depth[k];
        long long v[P_PER_BLOCK] = {cin, cin, conv_numel(d, cin, taps), d, d, d, d,
                                    conv_numel(d, d, 1), d, d, d, conv_numel(d, d, taps), d, d, d,
                                    conv_numel(d, d, 1), d, d, d, d, d, d};
        for (int i = 0; i < P_PER_BLOCK; ++i) net->param_numel.push_back(v[i]);
    }
    // rnn heads sit between the conv blocks and the FC head (module registration order of the reference: conv_modules,
    // rnns, output_transform); per head LayerNorm weight / bias, then the GRU tensors in named_parameters() order
    for (int k = c.start_deep_supervision_on; k < c.num_blocks && c.aggregation == 1; ++k) {
        long long d = c.depth[k];
        long long v[R_PER_HEAD] = {d, d, 3 * RNN_SIZE * d, 3 * RNN_SIZE * RNN_SIZE, 3 * RNN_SIZE, 3 * RNN_SIZE,
                                   3 * RNN_SIZE * d, 3 * RNN_SIZE * RNN_SIZE, 3 * RNN_SIZE, 3 * RNN_SIZE};
        for (int i = 0; i < R_PER_HEAD; ++i) net->param_numel.push_back(v[i]);
    }
    net->head_param_base = (int)net->param_numel.size();
    long long D = net->D;
    long long h[H_COUNT] = {D, D, D * D, D, D, D, D, (long long)c.n_classes * D, c.n_classes};
    for (int i = 0; i < H_COUNT; ++i) net->param_numel.push_back(h[i]);
    net->param_offset.resize(net->param_numel.size());
    long long off = 0;
    for (size_t i = 0; i < net->param_numel.size(); ++i) {
        net->param_offset[i] = off;
        off += net->param_numel[i];
    }
    net->total_params = off;
}

void carve_bn(Bump& b, BnBuf& bn, int C) {
    bn.C = C;
    bn.Cs = round_up(C, 16);
    bn.scale = b.take<float>(bn.Cs);
    bn.shift = b.take<float>(bn.Cs);
    bn.mean = b.take<float>(bn.Cs);
    bn.invstd = b.take<float>(bn.Cs);
    bn.c1 = b.take<float>(bn.Cs);
    bn.c2 = b.take<float>(bn.Cs);
}

size_t plane_bytes(const Geo& g) { return (size_t)g.rows * g.Cs * 4; }   // f32 plane == hi + lo half planes

// Carves every buffer for (N, T); with base == nullptr this is the size query.
size_t carve(fsb_net* net, char* base, int N, int T, int training) {
    const fsb_net_config& c = net->cfg;
    const int prec = net->prec_f;            // packed-weight / scratch sizes do not depend on the number of products
    Bump b{base, 0};
    int frames = 1 + T / c.hop;
    net->frames = frames;
    net->feat_tables = b.take_bytes(fsb_feat_table_bytes(c.n_fft));
    net->d_fb_vals = b.take<float>(net->fb_vals.size() + 1);
    net->d_fb_off = b.take<int>(net->fb_off.size() + 1);
    net->d_fb_start = b.take<int>(net->fb_start.size() + 1);
    net->d_fb_len = b.take<int>(net->fb_len.size() + 1);
    net->feat = c.two_d ? b.take<float>((size_t)N * c.n_features * frames) : nullptr;

    net->blocks.assign(c.num_blocks, BlockPlan());
    size_t max_partials = (size_t)plain_stats_blocks() * 2 * 16;
    size_t max_wgrad = 0, max_rnn = 0;
    int Hin = c.two_d ? c.n_features : 1, Win = frames;
    int head_off = 0;
    for (int k = 0; k < c.num_blocks; ++k) {
        BlockPlan& B = net->blocks[k];
        B.Cin = k == 0 ? (c.two_d ? 2 : c.n_features) : c.depth[k - 1];
        B.C = c.depth[k];
        int padH = c.two_d ? 1 : 0;
        int H = c.two_d ? Hin / 2 : 1, W = Win / 2;
        B.g_in = make_geo(N, Hin, Win, B.Cin, padH, 1);
        B.g_full = make_geo(N, Hin, Win, B.C, padH, 1);
        B.g = make_geo(N, H, W, B.C, padH, 1);
        int kh = c.two_d ? 3 : 1;
        B.entry = make_conv_geom(B.g_in, B.Cin, B.C, kh, 3);
        B.c1 = make_conv_geom(B.g, B.C, B.C, 1, 1);
        B.c2 = make_conv_geom(B.g, B.C, B.C, kh, 3);
        B.c3 = make_conv_geom(B.g, B.C, B.C, 1, 1);
        carve_bn(b, B.bn_in, B.Cin);
        carve_bn(b, B.bn_a, B.C);
        carve_bn(b, B.bn1, B.C);
        carve_bn(b, B.bn2, B.C);
        carve_bn(b, B.bn3, B.C);
        bool direct0 = c.two_d && k == 0;
        // interior masks of the two geometries (element-wise kernels skip border rows through them)
        B.mask = b.take<unsigned char>((size_t)B.g.rows);
        B.mask_in = direct0 ? nullptr : b.take<unsigned char>((size_t)B.g_in.rows);
        B.g.mask = B.mask;
        B.g_in.mask = B.mask_in;
        B.g_full.mask = B.mask_in;
        B.pk_entry = direct0 ? nullptr : b.take_bytes(packed_weight_bytes(prec, B.entry));
        B.pk1 = b.take_bytes(packed_weight_bytes(prec, B.c1));
        B.pk2 = b.take_bytes(packed_weight_bytes(prec, B.c2));
        B.pk3 = b.take_bytes(packed_weight_bytes(prec, B.c3));
        if (direct0) {
            B.x_in = nullptr; B.u = nullptr; B.zf = nullptr;
        } else {
            B.x_in = k == 0 ? b.take<float>((size_t)B.g_in.rows * B.g_in.Cs) : net->blocks[k - 1].out;
            B.u = b.take_bytes(plane_bytes(B.g_in));
            B.zf = b.take<float>((size_t)B.g_full.rows * B.g_full.Cs);
        }
        size_t pe = (size_t)B.g.rows * B.g.Cs;
        B.pool_amax = (!direct0 && training) ? b.take<unsigned char>(pe) : nullptr;
        B.sign3 = training ? b.take<unsigned char>(pe / 8) : nullptr;
        B.zp = b.take<float>(pe);
        B.r0 = b.take_bytes(pe * 4);
        B.z1 = b.take<float>(pe);
        B.a1 = b.take_bytes(pe * 4);
        B.z2 = b.take<float>(pe);
        B.a2 = b.take_bytes(pe * 4);
        B.z3 = b.take<float>(pe);
        B.out = b.take<float>(pe);
        B.rnn_index = -1;
        B.argrow = nullptr;
        B.gmax_scratch = nullptr;
        if (k >= c.start_deep_supervision_on && c.aggregation == 1) {
            B.head_off = head_off;
            head_off += 2 * RNN_SIZE;
            B.rnn_index = k - c.start_deep_supervision_on;
            float* base_f = b.take<float>(rnn_head_floats(N, B.g.W, B.g.Cs, training));
            rnn_head_carve(B.rnn, base_f, N, B.g.H, B.g.W, B.C, B.g.Cs, training);
            for (int d = 0; d < 2; ++d) B.pk_rnn[d] = b.take_bytes(rnn_packed_bytes(B.C));
            if (training) {
                max_rnn = std::max(max_rnn, simt_wgrad_scratch_bytes(B.rnn.g_ih));
                max_rnn = std::max(max_rnn, simt_wgrad_scratch_bytes(B.rnn.g_hh));
            }
        } else if (k >= c.start_deep_supervision_on) {
            B.head_off = head_off;
            head_off += B.C;
            B.argrow = b.take<int>((size_t)N * B.C);
            B.gmax_scratch = b.take_bytes(gmax_scratch_bytes(B.g));
        } else {
            B.head_off = -1;
        }
        if (training) {
            B.d_out = b.take<float>(pe);
            B.da2 = b.take<float>(pe);
            B.da1 = b.take<float>(pe);
            B.dr0a = b.take<float>(pe);
            B.dr0b = b.take<float>(pe);
            B.dzp = b.take<float>(pe);      // (sized for float32; the compact backward uses half of each)
            B.dz3 = b.take_bytes(pe * 4);
            B.dz2 = b.take_bytes(pe * 4);
            B.dz1 = b.take_bytes(pe * 4);
            if (!direct0) {
                B.dzf = b.take_bytes(plane_bytes(B.g_full));
                B.du = b.take<float>((size_t)B.g_in.rows * B.g_in.Cs);
                max_wgrad = std::max(max_wgrad, wgrad_scratch_bytes(prec, B.entry));
            } else {
                net->conv0_scratch = b.take_bytes(conv0_bwd_scratch_bytes(B.g));
                net->conv0_amax = b.take<unsigned char>((size_t)B.g.rows * B.g.Cs);
            }
            max_wgrad = std::max(max_wgrad, wgrad_scratch_bytes(prec, B.c1));
            max_wgrad = std::max(max_wgrad, wgrad_scratch_bytes(prec, B.c2));
        }
        max_partials = std::max(max_partials, (size_t)ew_num_blocks(B.g_in) * 5 * B.g_in.Cs);
        max_partials = std::max(max_partials, (size_t)ew_num_blocks(B.g) * 5 * B.g.Cs);
        Hin = H; Win = W;
    }
    // head
    net->g_head = make_geo(1, 1, N, net->D, 0, 0);
    net->g_cls = make_geo(1, 1, N, c.n_classes, 0, 0);
    net->lin1 = make_conv_geom(net->g_head, net->D, net->D, 1, 1);
    net->lin5 = make_conv_geom(net->g_head, net->D, c.n_classes, 1, 1);
    carve_bn(b, net->hbn0, net->D);
    carve_bn(b, net->hbn2, net->D);
    net->pk_l1 = b.take_bytes(simt_packed_weight_bytes(net->lin1));
    net->pk_l5 = b.take_bytes(simt_packed_weight_bytes(net->lin5));
    net->head_scratch = b.take_bytes(std::max(simt_skinny_scratch_bytes(net->lin1), simt_skinny_scratch_bytes(net->lin5)));
    size_t he = (size_t)N * net->Ds;
    net->feats = b.take<float>(he);
    net->h0 = b.take<float>(he);
    net->z1h = b.take<float>(he);
    net->h1 = b.take<float>(he);
    net->zl = b.take<float>((size_t)N * net->CsCls);
    if (training) {
        net->dzl = b.take<float>((size_t)N * net->CsCls);
        net->dh1 = b.take<float>(he);
        net->dz1h = b.take<float>(he);
        net->dh0 = b.take<float>(he);
        net->dfeats = b.take<float>(he);
        max_wgrad = std::max(max_wgrad, simt_wgrad_scratch_bytes(net->lin1));
        max_wgrad = std::max(max_wgrad, simt_wgrad_scratch_bytes(net->lin5));
    }
    max_partials = std::max(max_partials, (size_t)ew_num_blocks(net->g_head) * 5 * net->g_head.Cs);
    for (int k = 0; k < c.num_blocks; ++k)
        max_partials = std::max(max_partials, (size_t)256 * 2 * net->blocks[k].g.Cs);   // one record per GEMM CTA
    net->partials = b.take<double>(max_partials);
    net->wgrad_scratch = b.take_bytes(max_wgrad + 256);
    net->rnn_scratch = max_rnn ? b.take_bytes(max_rnn + 256) : nullptr;
    net->gscale = b.take<unsigned>((size_t)c.num_blocks * B_PER_BLOCK);
    net->wl1 = b.take<float>((size_t)c.num_blocks * 4);
    net->gscale_out = b.take<unsigned>((size_t)c.num_blocks + 1);
    net->d_seed = b.take<unsigned long long>(1);
    return align_up(b.off, 256);
}

// ---- profiling helpers ---------------------------------------------------------------------------
struct Scope {
    fsb_net* net;
    cudaStream_t s;
    int idx;
    Scope(fsb_net* n, cudaStream_t st, int cat, double flops = 0.0) : net(n), s(st), idx(-1) {
        if (!net->profiling) return;
        if (net->ev_used + 2 > net->ev_pool.size()) {
            size_t old = net->ev_pool.size();
            net->ev_pool.resize(old + 256);
            for (size_t i = old; i < net->ev_pool.size(); ++i) cudaEventCreate(&net->ev_pool[i]);
        }
        fsb_net::Rec r{cat, (int)net->ev_used, (int)net->ev_used + 1, flops};
        net->ev_used += 2;
        cudaEventRecord(net->ev_pool[r.e0], s);
        idx = (int)net->recs.size();
        net->recs.push_back(r);
    }
    ~Scope() {
        if (idx >= 0) cudaEventRecord(net->ev_pool[net->recs[idx].e1], s);
    }
};

double conv_flops(const ConvGeom& c, const Geo& g) { return 2.0 * c.Cin * c.Cout * c.ntaps * (double)g.pixels; }

#define RUN(cat, flops, expr)               \
    do {                                    \
        Scope _sc(net, s, cat, flops);      \
        FSB_TRY(expr);                      \
    } while (0)

#define RUN_S(stream, cat, flops, expr)     \
    do {                                    \
        Scope _sc(net, stream, cat, flops); \
        FSB_TRY(expr);                      \
    } while (0)

const Residual kNoRes = {nullptr, nullptr, nullptr, nullptr};
const Dropout kNoDrop = {0.f, 0ull, nullptr};

// finalize a BatchNorm from `nblk` partial records already sitting in net->partials (training) or from the
// running statistics (eval)
int bn_finalize_from(fsb_net* net, cudaStream_t s, int nblk, const Geo& g, BnBuf& bn, const float* gamma,
                     const float* beta, float* rm, float* rv, long long* cnt, int training, int cat) {
    RUN(cat, 0, bn_finalize(net->partials, nblk, g.pixels, gamma, beta, rm, rv, cnt, training, bn.C, bn.Cs, bn.scale,
                            bn.shift, bn.mean, bn.invstd, s));
    return 0;
}

// stand-alone statistics pass + finalize
int bn_forward_stats(fsb_net* net, cudaStream_t s, const float* x, const Geo& g, BnBuf& bn, const float* gamma,
                     const float* beta, float* rm, float* rv, long long* cnt, int training, int cat) {
    if (training) RUN(cat, 0, pf_stats(x, g, net->partials, s));
    return bn_finalize_from(net, s, ew_num_blocks(g), g, bn, gamma, beta, rm, rv, cnt, training, cat);
}

}  // namespace

// =================================================================================================
extern "C" int fsb_net_create(const fsb_net_config* cfg, const float* fb_vals, const int* fb_off,
                              const int* fb_start, const int* fb_len, int fb_nnz, fsb_net** out) {
    FSB_REQUIRE(cfg && out, "net_create: null argument");
    FSB_REQUIRE(cfg->num_blocks >= 1 && cfg->num_blocks <= FSB_MAX_BLOCKS, "net_create: num_blocks out of range");
    FSB_REQUIRE(cfg->precision >= 0 && cfg->precision <= 3, "net_create: precision must be 0 (fp32), 1 (fp16x3), 2 (fp16) or 3 (mixed)");
    FSB_REQUIRE(cfg->feat_mode == 1 || cfg->feat_mode == 2, "net_create: feat_mode must be 1 (stft) or 2 (mel)");
    FSB_REQUIRE(cfg->start_deep_supervision_on < cfg->num_blocks, "net_create: no deep-supervision head");
    FSB_REQUIRE(cfg->dropout_p >= 0.f && cfg->dropout_p < 1.f, "net_create: dropout must be in [0, 1)");
    fsb_net* net = new fsb_net();
    net->cfg = *cfg;
    net->prec_f = cfg->precision == 3 ? 1 : cfg->precision;
    net->prec_b = cfg->precision == 3 ? 2 : cfg->precision;
    if (cfg->feat_mode == 2) {
        FSB_REQUIRE(fb_vals && fb_off && fb_start && fb_len && fb_nnz > 0, "net_create: mel mode needs a filterbank");
        int n_mel = cfg->n_features;
        net->fb_vals.assign(fb_vals, fb_vals + fb_nnz);
        net->fb_off.assign(fb_off, fb_off + n_mel);
        net->fb_start.assign(fb_start, fb_start + n_mel);
        net->fb_len.assign(fb_len, fb_len + n_mel);
    }
    FSB_REQUIRE(cfg->aggregation == 0 || (cfg->aggregation == 1 && cfg->two_d),
                "net_create: aggregation must be 0 (max) or 1 (rnn, 2D model only)");
    net->D = 0;
    for (int k = cfg->start_deep_supervision_on; k < cfg->num_blocks; ++k)
        net->D += cfg->aggregation == 1 ? 2 * RNN_SIZE : cfg->depth[k];
    net->n_rnn = cfg->aggregation == 1 ? cfg->num_blocks - cfg->start_deep_supervision_on : 0;
    net->Ds = round_up(net->D, 16);
    net->CsCls = round_up(cfg->n_classes, 16);
    build_param_table(net);
    net->overlap = getenv("FSB200_NO_OVERLAP") == nullptr;      // read once; fsb_net_set_overlap changes it later
    {
        const char* e = getenv("FSB200_CONV0_TC");
        net->conv0_tc = !(e && atoi(e) == 0);
        e = getenv("FSB200_GRAPHS");
        net->graphs = !(e && atoi(e) == 0);
        e = getenv("FSB200_FUSE_EVAL");
        net->fuse_eval = e ? atoi(e) & 15 : 14;
        e = getenv("FSB200_COMPACT_BWD");
        net->compact = net->prec_b == 2 && !(e && atoi(e) == 0);
    }
    *out = net;
    return 0;
}

static void drop_graphs(fsb_net* net) {
    for (auto& g : net->graph_cache) cudaGraphExecDestroy(g.exec);
    net->graph_cache.clear();
    net->seen_keys.clear();
}

extern "C" void fsb_net_destroy(fsb_net* net) {
    if (!net) return;
    drop_graphs(net);
    for (cudaEvent_t e : net->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : net->fork_events) cudaEventDestroy(e);
    if (net->exec_in) cudaEventDestroy(net->exec_in);
    if (net->exec_out) cudaEventDestroy(net->exec_out);
    if (net->exec_stream) cudaStreamDestroy(net->exec_stream);
    if (net->join_event) cudaEventDestroy(net->join_event);
    if (net->pack_event) cudaEventDestroy(net->pack_event);
    if (net->side) cudaStreamDestroy(net->side);
    delete net;
}

extern "C" int fsb_net_num_params(const fsb_net* net) { return (int)net->param_numel.size(); }
extern "C" int fsb_net_num_bn(const fsb_net* net) { return net->cfg.num_blocks * B_PER_BLOCK + 2; }
extern "C" long long fsb_net_param_numel(const fsb_net* net, int index) {
    return index >= 0 && index < (int)net->param_numel.size() ? net->param_numel[index] : -1;
}

static int check_shape(const fsb_net* net, int n, int t) {
    const fsb_net_config& c = net->cfg;
    FSB_REQUIRE(n >= 1 && t > c.n_fft / 2, "net: need N >= 1 and T > n_fft/2 (N=%d, T=%d)", n, t);
    int frames = 1 + t / c.hop;
    FSB_REQUIRE((long long)n * (c.two_d ? c.n_features : 1) * frames < (1ll << 31),
                "net: batch too large for 32-bit pixel indices (N=%d, frames=%d)", n, frames);
    int w = frames, h = c.two_d ? c.n_features : 1;
    for (int k = 0; k < c.num_blocks; ++k) {
        w /= 2;
        if (c.two_d) h /= 2;
    }
    FSB_REQUIRE(w >= 1 && h >= 1, "net: clip too short for %d pooling stages (T=%d -> %d frames)", c.num_blocks, t,
                frames);
    return 0;
}

extern "C" size_t fsb_net_workspace_bytes(const fsb_net* net, int n, int t, int training) {
    if (check_shape(net, n, t) != 0) return 0;
    fsb_net tmp = *net;     // carve mutates the plan; run the size query on a copy
    tmp.ev_pool.clear();
    tmp.fork_events.clear();
    return carve(&tmp, nullptr, n, t, training);
}

static int bind(fsb_net* net, void* ws, size_t ws_bytes, int n, int t, int training, cudaStream_t s) {
    if (net->ws == ws && net->N == n && net->T == t && net->training == training && net->ws_bytes == ws_bytes) return 0;
    FSB_TRY(check_shape(net, n, t));
    size_t need = carve(net, nullptr, n, t, training);
    if (ws_bytes < need) {
        set_error("net: workspace too small (%zu < %zu)", ws_bytes, need);
        return FSB_E_WORKSPACE;
    }
    carve(net, (char*)ws, n, t, training);
    drop_graphs(net);          // captured launch sequences point into the previous carving
    net->ws = ws; net->ws_bytes = ws_bytes; net->N = n; net->T = t; net->training = training;
    net->fwd_done = false;
    // zero once: GEMM inputs rely on zero borders / zero channel tails, which the element-wise
    // kernels (interior-only writers) then preserve for the lifetime of the binding
    FSB_CUDA(cudaMemsetAsync(ws, 0, need, s));
    const fsb_net_config& c = net->cfg;
    for (BlockPlan& B : net->blocks) {
        FSB_TRY(pf_build_mask(B.g, B.mask, s));
        if (B.mask_in) FSB_TRY(pf_build_mask(B.g_in, B.mask_in, s));
    }
    FSB_TRY(fsb_feat_init_tables(c.n_fft, net->feat_tables, s));
    if (c.feat_mode == 2) {
        FSB_CUDA(cudaMemcpyAsync(net->d_fb_vals, net->fb_vals.data(), net->fb_vals.size() * 4, cudaMemcpyHostToDevice, s));
        FSB_CUDA(cudaMemcpyAsync(net->d_fb_off, net->fb_off.data(), net->fb_off.size() * 4, cudaMemcpyHostToDevice, s));
        FSB_CUDA(cudaMemcpyAsync(net->d_fb_start, net->fb_start.data(), net->fb_start.size() * 4, cudaMemcpyHostToDevice, s));
        FSB_CUDA(cudaMemcpyAsync(net->d_fb_len, net->fb_len.data(), net->fb_len.size() * 4, cudaMemcpyHostToDevice, s));
    }
    return 0;
}

static int ensure_side_stream(fsb_net* net) {
    if (!net->side) {
        int lo = 0, hi = 0;
        FSB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        FSB_CUDA(cudaStreamCreateWithPriority(&net->side, cudaStreamNonBlocking, lo));     // off the critical path
        FSB_CUDA(cudaEventCreateWithFlags(&net->join_event, cudaEventDisableTiming));
        FSB_CUDA(cudaEventCreateWithFlags(&net->pack_event, cudaEventDisableTiming));
    }
    return 0;
}

// ---- CUDA-graph replay of a launch sequence ------------------------------------------------------------------------
namespace {

__global__ void set_seed_kernel(unsigned long long* dst, unsigned long long v) { *dst = v; }

struct KeyHash {
    unsigned long long h = 1469598103934665603ull;
    void add(const void* p, size_t bytes) {
        const unsigned char* b = (const unsigned char*)p;
        for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    }
    template <typename T> void add(const T& v) { add(&v, sizeof(T)); }
};

// Runs `body` (which only enqueues work on `s` and on the plan's side stream) either eagerly or as a captured graph:
// a key seen for the first time runs eagerly (one-off shapes never pay for a capture, and every lazy initialisation --
// streams, function attributes, tensor maps -- happens outside capture); its second occurrence is captured and
// instantiated; later occurrences replay the graph.
template <typename Body>
int run_graphed(fsb_net* net, unsigned long long key, cudaStream_t caller, Body body) {
    if (!net->graphs || net->profiling) return body(caller);
    if (!net->exec_stream) {
        FSB_CUDA(cudaStreamCreateWithFlags(&net->exec_stream, cudaStreamNonBlocking));
        FSB_CUDA(cudaEventCreateWithFlags(&net->exec_in, cudaEventDisableTiming));
        FSB_CUDA(cudaEventCreateWithFlags(&net->exec_out, cudaEventDisableTiming));
    }
    cudaStream_t s = net->exec_stream;
    FSB_CUDA(cudaEventRecord(net->exec_in, caller));
    FSB_CUDA(cudaStreamWaitEvent(s, net->exec_in, 0));
    auto finish = [&](int rc) -> int {
        if (rc != 0) return rc;
        FSB_CUDA(cudaEventRecord(net->exec_out, s));
        FSB_CUDA(cudaStreamWaitEvent(caller, net->exec_out, 0));
        return 0;
    };
    ++net->graph_clock;
    for (auto& g : net->graph_cache)
        if (g.key == key) {
            g.stamp = net->graph_clock;
            FSB_CUDA(cudaGraphLaunch(g.exec, s));
            g_launch_count += g.launches;
            return finish(0);
        }
    bool seen = false;
    for (unsigned long long k : net->seen_keys) seen = seen || k == key;
    if (!seen) {
        if (net->seen_keys.size() > 256) net->seen_keys.clear();
        net->seen_keys.push_back(key);
        return finish(body(s));
    }
    const long long before = g_launch_count;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        net->graphs = false;              // capture is not possible in this context: stay eager from here on
        return finish(body(s));
    }
    const int rc = body(s);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc != 0 || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc != 0) return rc;
        net->graphs = false;
        return finish(body(s));
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess || !exec) {
        cudaGetLastError();
        net->graphs = false;
        return finish(body(s));
    }
    if (net->graph_cache.size() >= 16) {          // evict the least recently used sequence
        size_t victim = 0;
        for (size_t i = 1; i < net->graph_cache.size(); ++i)
            if (net->graph_cache[i].stamp < net->graph_cache[victim].stamp) victim = i;
        cudaGraphExecDestroy(net->graph_cache[victim].exec);
        net->graph_cache.erase(net->graph_cache.begin() + victim);
    }
    net->graph_cache.push_back({key, exec, g_launch_count - before, net->graph_clock});
    FSB_CUDA(cudaGraphLaunch(exec, s));
    return finish(0);
}

}  // namespace

static int forward_impl(fsb_net* net, const float* signal, const float* features, int n, int t, long long signal_stride,
                        const float* const* params, float* const* bn_mean, float* const* bn_var,
                        long long* const* bn_count, int training, unsigned long long dropout_seed, float* logits,
                        cudaStream_t s);

static int forward_entry(fsb_net* net, const float* signal, const float* features, int n, int t, long long signal_stride,
                         const float* const* params, float* const* bn_mean, float* const* bn_var,
                         long long* const* bn_count, int training, unsigned long long dropout_seed, void* workspace,
                         size_t workspace_bytes, float* logits, void* stream);

extern "C" int fsb_net_forward(fsb_net* net, const float* signal, int n, int t, long long signal_stride,
                               const float* const* params, float* const* bn_mean, float* const* bn_var,
                               long long* const* bn_count, int training, unsigned long long dropout_seed,
                               void* workspace, size_t workspace_bytes, float* logits, void* stream) {
    FSB_REQUIRE(signal, "net_forward: null signal");
    return forward_entry(net, signal, nullptr, n, t, signal_stride, params, bn_mean, bn_var, bn_count, training, dropout_seed,
                         workspace, workspace_bytes, logits, stream);
}

extern "C" int fsb_net_forward_features(fsb_net* net, const float* features, int n, int t, const float* const* params,
                                        float* const* bn_mean, float* const* bn_var, long long* const* bn_count,
                                        int training, unsigned long long dropout_seed, void* workspace,
                                        size_t workspace_bytes, float* logits, void* stream) {
    FSB_REQUIRE(features, "net_forward_features: null features");
    return forward_entry(net, nullptr, features, n, t, 0, params, bn_mean, bn_var, bn_count, training, dropout_seed, workspace,
                         workspace_bytes, logits, stream);
}

static int forward_entry(fsb_net* net, const float* signal, const float* features, int n, int t, long long signal_stride,
                         const float* const* params, float* const* bn_mean, float* const* bn_var,
                         long long* const* bn_count, int training, unsigned long long dropout_seed, void* workspace,
                         size_t workspace_bytes, float* logits, void* stream) {
    FSB_REQUIRE(net && params && bn_mean && bn_var && workspace && logits, "net_forward: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    FSB_TRY(fsb_device_ok());
    FSB_TRY(bind(net, workspace, workspace_bytes, n, t, training ? 1 : 0, s));
    net->fwd_done = false;
    // the dropout seed changes every call: it travels through device memory, outside the replayed sequence
    set_seed_kernel<<<1, 1, 0, s>>>(net->d_seed, dropout_seed);
    FSB_LAUNCHED();
    KeyHash k;
    const int np = fsb_net_num_params(net), nb = fsb_net_num_bn(net);
    k.add(signal); k.add(features); k.add(n); k.add(t); k.add(signal_stride); k.add(training); k.add(logits);
    k.add(workspace); k.add(net->overlap); k.add(1);
    k.add(params, sizeof(float*) * np);
    k.add(bn_mean, sizeof(float*) * nb);
    k.add(bn_var, sizeof(float*) * nb);
    if (bn_count) k.add(bn_count, sizeof(long long*) * nb);
    FSB_TRY(run_graphed(net, k.h, s, [&](cudaStream_t es) {
        return forward_impl(net, signal, features, n, t, signal_stride, params, bn_mean, bn_var, bn_count, training,
                            dropout_seed, logits, es);
    }));
    net->fwd_done = true;
    return 0;
}

// features != nullptr: log features (N, n_features, frames) computed earlier (fsb_feat_forward) replace the feature kernel
static int forward_impl(fsb_net* net, const float* signal, const float* features, int n, int t, long long signal_stride,
                        const float* const* params, float* const* bn_mean, float* const* bn_var,
                        long long* const* bn_count, int training, unsigned long long dropout_seed, float* logits,
                        cudaStream_t s) {
    const fsb_net_config& c = net->cfg;
    const int prec = net->prec_f, fmt = act_fmt(prec);
    net->recs.clear();
    net->ev_used = 0;
    net->dropout_seed = dropout_seed;
    const int frames = net->frames;
    int carried_nblk = 0;      // partial records left in net->partials by the previous block's last pass
    bool u_ready = false;      // eval: the previous block's output kernel already wrote this block's BN_in output u

    // Weight packing (float32 -> bf16 hi/lo K-major tiles) depends on the parameters only: it runs on the side stream
    // in the shadow of the feature kernel and the block-0 entry conv; the first GEMM waits for it.
    FSB_TRY(ensure_side_stream(net));
    {
        const bool overlap = net->overlap;
        cudaStream_t ps = overlap ? net->side : s;
        if (overlap) {
            FSB_CUDA(cudaEventRecord(net->pack_event, s));        // orders the packs after the caller's earlier work (optimizer step)
            FSB_CUDA(cudaStreamWaitEvent(net->side, net->pack_event, 0));
        }
        for (int k = 0; k < c.num_blocks; ++k) {
            BlockPlan& B = net->blocks[k];
            const float* const* P = params + (size_t)k * P_PER_BLOCK;
            if (B.pk_entry) RUN_S(ps, CAT_PACK, 0, pack_weights(prec, P[P_CONV_W], P[P_CONV_B], B.entry, B.pk_entry, ps));
            RUN_S(ps, CAT_PACK, 0, pack_weights(prec, P[P_C1_W], P[P_C1_B], B.c1, B.pk1, ps));
            RUN_S(ps, CAT_PACK, 0, pack_weights(prec, P[P_C2_W], P[P_C2_B], B.c2, B.pk2, ps));
            RUN_S(ps, CAT_PACK, 0, pack_weights(prec, P[P_C3_W], P[P_C3_B], B.c3, B.pk3, ps));
        }
        {
            const float* const* P = params + net->head_param_base;
            RUN_S(ps, CAT_HEAD, 0, simt_pack_weights(P[H_L1_W], P[H_L1_B], net->lin1, net->pk_l1, ps));
            RUN_S(ps, CAT_HEAD, 0, simt_pack_weights(P[H_L5_W], P[H_L5_B], net->lin5, net->pk_l5, ps));
        }
        if (training && net->compact) {
            // Hoelder bounds of the dgrad outputs (GradScale of the half planes the compact backward writes)
            WeightL1Job jobs[32];
            int nj = 0;
            for (int k = 0; k < c.num_blocks && nj + 4 <= 32; ++k) {
                const BlockPlan& B = net->blocks[k];
                const float* const* P = params + (size_t)k * P_PER_BLOCK;
                jobs[nj++] = WeightL1Job{P[P_CONV_W], B.entry.Cin, B.entry.Cout, B.entry.ntaps};
                jobs[nj++] = WeightL1Job{P[P_C1_W], B.c1.Cin, B.c1.Cout, B.c1.ntaps};
                jobs[nj++] = WeightL1Job{P[P_C2_W], B.c2.Cin, B.c2.Cout, B.c2.ntaps};
                jobs[nj++] = WeightL1Job{P[P_C3_W], B.c3.Cin, B.c3.Cout, B.c3.ntaps};
            }
            RUN_S(ps, CAT_PACK, 0, weight_l1_bounds(jobs, nj, net->wl1, ps));
        }
        if (overlap) FSB_CUDA(cudaEventRecord(net->pack_event, net->side));
    }
    bool packs_joined = !net->overlap;
    auto join_packs = [&]() -> int {
        if (!packs_joined) {
            FSB_CUDA(cudaStreamWaitEvent(s, net->pack_event, 0));
            packs_joined = true;
        }
        return 0;
    };

    for (int k = 0; k < c.num_blocks; ++k) {
        BlockPlan& B = net->blocks[k];
        bool zp_stats_ready = false, r0_ready = false;
        const bool fuse_pool = !training && prec != 0 && (net->fuse_eval & 8);
        const float* const* P = params + (size_t)k * P_PER_BLOCK;
        float* const* RM = bn_mean + (size_t)k * B_PER_BLOCK;
        float* const* RV = bn_var + (size_t)k * B_PER_BLOCK;
        long long* const* CT = bn_count ? bn_count + (size_t)k * B_PER_BLOCK : nullptr;
        auto cnt = [&](int i) { return CT ? CT[i] : nullptr; };

        if (c.two_d && k == 0) {
            if (features)
                FSB_CUDA(cudaMemcpyAsync(net->feat, features, (size_t)n * c.n_features * frames * sizeof(float),
                                         cudaMemcpyDeviceToDevice, s));
            else
                RUN(CAT_FEAT, 0, fsb_feat_forward(signal, n, signal_stride, t, c.n_fft, c.hop, c.feat_mode, 1e-4f,
                                                  c.n_features, net->d_fb_vals, net->d_fb_off, net->d_fb_start,
                                                  net->d_fb_len, net->feat_tables, net->feat,
                                                  (long long)c.n_features * frames, frames, 1, s));
            long long cnt0 = (long long)n * c.n_features * frames;
            if (training) {
                RUN(CAT_ELT_FWD, 0, plain_stats(net->feat, cnt0, net->partials, s));
                RUN(CAT_ELT_FWD, 0, freq_encoding_stats(c.n_features, (long long)n * frames, net->partials, s));
            }
            RUN(CAT_ELT_FWD, 0, bn_finalize(net->partials, plain_stats_blocks(), cnt0, P[P_BNIN_W], P[P_BNIN_B],
                                            RM[B_IN], RV[B_IN], cnt(B_IN), training, 2, 16, B.bn_in.scale,
                                            B.bn_in.shift, B.bn_in.mean, B.bn_in.invstd, s));
            if (prec != 0 && net->conv0_tc && conv0_tc_supported(B.g))
                RUN(CAT_CONV0, 2.0 * 2 * B.C * 9 * (double)n * c.n_features * frames,
                    conv0_tc_forward(prec, net->feat, n, c.n_features, frames, B.bn_in.scale, B.bn_in.shift, P[P_CONV_W],
                                     P[P_CONV_B], B.zp, training ? net->conv0_amax : nullptr, B.g, s));
            else
                RUN(CAT_CONV0, 2.0 * 2 * B.C * 9 * (double)n * c.n_features * frames,
                    conv0_forward(net->feat, n, c.n_features, frames, B.bn_in.scale, B.bn_in.shift, P[P_CONV_W],
                                  P[P_CONV_B], B.zp, training ? net->conv0_amax : nullptr, B.g, s));
        } else {
            if (k == 0) {
                // 1D: features land directly in the padded-flat block input (channels = STFT bins)
                float* dst = B.x_in + geo_row(B.g_in, 0, 0, 0) * B.g_in.Cs;
                if (features)       // (N, F, frames) = NCHW with H = 1: transpose into the padded-flat block input
                    RUN(CAT_FEAT, 0, nchw_to_pf(features, B.g_in, B.x_in, FMT_F32, s));
                else
                    RUN(CAT_FEAT, 0, fsb_feat_forward(signal, n, signal_stride, t, c.n_fft, c.hop, c.feat_mode, 1e-4f,
                                                      c.n_features, net->d_fb_vals, net->d_fb_off, net->d_fb_start,
                                                      net->d_fb_len, net->feat_tables, dst,
                                                      (long long)B.g_in.Hp * B.g_in.Wp * B.g_in.Cs, 1, B.g_in.Cs, s));
                FSB_TRY(bn_forward_stats(net, s, B.x_in, B.g_in, B.bn_in, P[P_BNIN_W], P[P_BNIN_B], RM[B_IN], RV[B_IN],
                                         cnt(B_IN), training, CAT_ELT_FWD));
            } else if (!u_ready) {
                // statistics of the block input were gathered by the previous block's last element-wise pass
                FSB_TRY(bn_finalize_from(net, s, carried_nblk, B.g_in, B.bn_in, P[P_BNIN_W], P[P_BNIN_B], RM[B_IN],
                                         RV[B_IN], cnt(B_IN), training, CAT_ELT_FWD));
            }
            if (!u_ready)
                RUN(CAT_ELT_FWD, 0, bn_act_forward(B.x_in, B.g_in, B.bn_in.coef(nullptr), kNoRes, kNoDrop, B.u, fmt,
                                                   nullptr, nullptr, s));
            u_ready = false;
            FSB_TRY(join_packs());
            RUN(CAT_GEMM_FWD, conv_flops(B.entry, B.g_in),
                conv_gemm_fwd(prec, B.u, B.pk_entry, B.zf, B.entry, nullptr, s));
            if (fuse_pool) {
                // eval: BN_a + PReLU_a (fixed affine map) ride in the pooling kernel: zp and r0 in one pass
                FSB_TRY(bn_finalize_from(net, s, 0, B.g, B.bn_a, P[P_BNA_W], P[P_BNA_B], RM[B_A], RV[B_A], cnt(B_A), 0,
                                         CAT_ELT_FWD));
                const BnCoef ca = B.bn_a.coef(P[P_PRELUA]);
                RUN(CAT_ELT_FWD, 0, maxpool_forward(B.zf, B.g_full, B.zp, B.g, c.two_d ? 2 : 1, nullptr, nullptr, s, &ca,
                                                    B.r0, fmt));
                r0_ready = true;
            } else {
                RUN(CAT_ELT_FWD, 0, maxpool_forward(B.zf, B.g_full, B.zp, B.g, c.two_d ? 2 : 1,
                                                    training ? net->partials : nullptr, B.pool_amax, s));
            }
            zp_stats_ready = true;
        }
        // BN_a + PReLU_a -> r0
        if (!r0_ready) {
            if (zp_stats_ready)
                FSB_TRY(bn_finalize_from(net, s, ew_num_blocks(B.g), B.g, B.bn_a, P[P_BNA_W], P[P_BNA_B], RM[B_A], RV[B_A],
                                         cnt(B_A), training, CAT_ELT_FWD));
            else
                FSB_TRY(bn_forward_stats(net, s, B.zp, B.g, B.bn_a, P[P_BNA_W], P[P_BNA_B], RM[B_A], RV[B_A], cnt(B_A),
                                         training, CAT_ELT_FWD));
            RUN(CAT_ELT_FWD, 0, bn_act_forward(B.zp, B.g, B.bn_a.coef(P[P_PRELUA]), kNoRes, kNoDrop, B.r0, fmt, nullptr,
                                               nullptr, s));
        }
        // resnet block: every conv GEMM gathers the batch statistics of its output in the epilogue
        FSB_TRY(join_packs());
        int nblk = 0;
        FwdStats st = {net->partials, &B.g, &nblk};
        const FwdStats* stp = training ? &st : nullptr;
        // eval: BatchNorm is a fixed affine map (running statistics), so BN + PReLU can ride in the epilogue of the producing
        // GEMM and the activation is written straight as operand planes (the float32 pre-activation is never materialised).
        // fuse_eval bit 0: conv1 (1x1 -- its epilogue is the kernel's bottleneck, so the fold roughly breaks even),
        // bit 1: conv2 (3x3, tensor-pipe bound: the fold is free).
        const bool fuse1 = !training && prec != 0 && (net->fuse_eval & 1), fuse2 = !training && prec != 0 && (net->fuse_eval & 2);
        if (fuse1) {
            FSB_TRY(bn_finalize_from(net, s, 0, B.g, B.bn1, P[P_BN1_W], P[P_BN1_B], RM[B_1], RV[B_1], cnt(B_1), 0, CAT_ELT_FWD));
            const FwdAct act1 = {B.bn1.scale, B.bn1.shift, P[P_PRELU1], B.C, B.g.mask};
            RUN(CAT_GEMM_FWD, conv_flops(B.c1, B.g), conv_gemm_fwd_act(prec, B.r0, B.pk1, B.a1, B.c1, act1, s));
        } else {
            RUN(CAT_GEMM_FWD, conv_flops(B.c1, B.g), conv_gemm_fwd(prec, B.r0, B.pk1, B.z1, B.c1, stp, s));
            FSB_TRY(bn_finalize_from(net, s, nblk, B.g, B.bn1, P[P_BN1_W], P[P_BN1_B], RM[B_1], RV[B_1], cnt(B_1), training,
                                     CAT_ELT_FWD));
            RUN(CAT_ELT_FWD, 0, bn_act_forward(B.z1, B.g, B.bn1.coef(P[P_PRELU1]), kNoRes, kNoDrop, B.a1, fmt, nullptr,
                                               nullptr, s));
        }
        if (fuse2) {
            FSB_TRY(bn_finalize_from(net, s, 0, B.g, B.bn2, P[P_BN2_W], P[P_BN2_B], RM[B_2], RV[B_2], cnt(B_2), 0, CAT_ELT_FWD));
            const FwdAct act2 = {B.bn2.scale, B.bn2.shift, P[P_PRELU2], B.C, B.g.mask};
            RUN(CAT_GEMM_FWD, conv_flops(B.c2, B.g), conv_gemm_fwd_act(prec, B.a1, B.pk2, B.a2, B.c2, act2, s));
        } else {
            RUN(CAT_GEMM_FWD, conv_flops(B.c2, B.g), conv_gemm_fwd(prec, B.a1, B.pk2, B.z2, B.c2, stp, s));
            FSB_TRY(bn_finalize_from(net, s, nblk, B.g, B.bn2, P[P_BN2_W], P[P_BN2_B], RM[B_2], RV[B_2], cnt(B_2), training,
                                     CAT_ELT_FWD));
            RUN(CAT_ELT_FWD, 0, bn_act_forward(B.z2, B.g, B.bn2.coef(P[P_PRELU2]), kNoRes, kNoDrop, B.a2, fmt, nullptr,
                                               nullptr, s));
        }
        RUN(CAT_GEMM_FWD, conv_flops(B.c3, B.g), conv_gemm_fwd(prec, B.a2, B.pk3, B.z3, B.c3, stp, s));
        FSB_TRY(bn_finalize_from(net, s, nblk, B.g, B.bn3, P[P_BN3_W], P[P_BN3_B], RM[B_3], RV[B_3], cnt(B_3), training,
                                 CAT_ELT_FWD));
        Residual res = {B.zp, B.bn_a.scale, B.bn_a.shift, P[P_PRELUA]};
        // the block output feeds the next block's input BatchNorm: gather its statistics here
        const bool next_stats = training && k + 1 < c.num_blocks;
        if (!training && prec != 0 && (net->fuse_eval & 4) && k + 1 < c.num_blocks) {
            // eval: the next block's input BatchNorm (fixed affine map) rides in this kernel: out (float32, for the head)
            // and u of block k + 1 (operand planes of its entry conv) in one pass
            BlockPlan& Bn = net->blocks[k + 1];
            const float* const* Pn = params + (size_t)(k + 1) * P_PER_BLOCK;
            FSB_TRY(bn_finalize_from(net, s, 0, Bn.g_in, Bn.bn_in, Pn[P_BNIN_W], Pn[P_BNIN_B],
                                     bn_mean[(size_t)(k + 1) * B_PER_BLOCK + B_IN], bn_var[(size_t)(k + 1) * B_PER_BLOCK + B_IN],
                                     bn_count ? bn_count[(size_t)(k + 1) * B_PER_BLOCK + B_IN] : nullptr, 0, CAT_ELT_FWD));
            const BnCoef cn = Bn.bn_in.coef(nullptr);
            RUN(CAT_ELT_FWD, 0, bn_act_forward(B.z3, B.g, B.bn3.coef(P[P_PRELU3]), res, kNoDrop, Bn.u, fmt, B.out, nullptr, s,
                                               &cn));
            u_ready = true;
        } else {
            RUN(CAT_ELT_FWD, 0, bn_act_forward(B.z3, B.g, B.bn3.coef(P[P_PRELU3]), res, kNoDrop, nullptr, fmt, B.out,
                                               next_stats ? net->partials : nullptr, s));
        }
        carried_nblk = ew_num_blocks(B.g);
        if (B.rnn_index >= 0) {
            const float* const* PR = params + (size_t)c.num_blocks * P_PER_BLOCK + (size_t)B.rnn_index * R_PER_HEAD;
            RUN(CAT_HEAD, 0, rnn_head_forward(B.rnn, B.out, B.g, PR, B.pk_rnn, net->feats, net->Ds, B.head_off, training, s));
        } else if (B.head_off >= 0) {
            RUN(CAT_ELT_FWD, 0, gmax_forward(B.out, B.g, net->feats, net->Ds, B.head_off, B.argrow, B.gmax_scratch, s));
        }
    }

    // ---- FC head (float32 CUDA-core GEMMs in every precision mode)
    {
        const float* const* P = params + net->head_param_base;
        float* const* RM = bn_mean + (size_t)c.num_blocks * B_PER_BLOCK;
        float* const* RV = bn_var + (size_t)c.num_blocks * B_PER_BLOCK;
        long long* const* CT = bn_count ? bn_count + (size_t)c.num_blocks * B_PER_BLOCK : nullptr;
        FSB_TRY(bn_forward_stats(net, s, net->feats, net->g_head, net->hbn0, P[H_BN0_W], P[H_BN0_B], RM[0], RV[0],
                                 CT ? CT[0] : nullptr, training, CAT_HEAD));
        RUN(CAT_HEAD, 0, bn_act_forward(net->feats, net->g_head, net->hbn0.coef(nullptr), kNoRes, kNoDrop, nullptr,
                                        FMT_F32, net->h0, nullptr, s));
        FSB_TRY(join_packs());
        RUN(CAT_HEAD, 0, simt_skinny_fwd(net->h0, net->pk_l1, net->z1h, net->lin1, net->head_scratch, s));
        FSB_TRY(bn_forward_stats(net, s, net->z1h, net->g_head, net->hbn2, P[H_BN2_W], P[H_BN2_B], RM[1], RV[1],
                                 CT ? CT[1] : nullptr, training, CAT_HEAD));
        Dropout dr = {training ? c.dropout_p : 0.f, dropout_seed, net->d_seed};
        RUN(CAT_HEAD, 0, bn_act_forward(net->z1h, net->g_head, net->hbn2.coef(P[H_PRELU]), kNoRes, dr, nullptr,
                                        FMT_F32, net->h1, nullptr, s));
        RUN(CAT_HEAD, 0, simt_skinny_fwd(net->h1, net->pk_l5, net->zl, net->lin5, net->head_scratch, s));
        RUN(CAT_HEAD, 0, copy2d(net->zl, n, c.n_classes, net->CsCls, logits, c.n_classes, s));
    }
    return 0;
}

// =================================================================================================
// absmax (optional): GradScale slot of the gradient tensor this BatchNorm emits (half-precision planes are written
// with its power-of-two scale; float32 outputs ignore it but the slot is still filled for a later converter)
static int bn_backward(fsb_net* net, cudaStream_t s, GradRef dA1, GradRef dA2, const float* z, const void* a_hi,
                       const Geo& g, BnBuf& bn, const float* slope, Residual res, Dropout dr, float* dgamma, float* dbeta,
                       float* dslope, void* dz, int fmt, void* dres, unsigned* dres_bits, unsigned* absmax, int cat,
                       const unsigned* extra_bits = nullptr) {
    BnCoef coef = bn.coef(slope);
    RUN(cat, 0, bn_act_bwd_reduce(dA1, dA2, z, a_hi, g, coef, res, dr, net->partials, s));
    RUN(cat, 0, bn_bwd_finalize(net->partials, bn_bwd_num_blocks(dA1, dA2, g, res, dr), g.pixels, bn.C, bn.Cs, bn.scale,
                                dgamma, dbeta, dslope, bn.c1, bn.c2, absmax, dres_bits, extra_bits, s));
    if (dz) RUN(cat, 0, bn_act_bwd_apply(dA1, dA2, z, a_hi, g, coef, res, dr, bn.c1, bn.c2, dz, fmt, dres, dres_bits,
                                         absmax, s));
    return 0;
}

static int backward_impl(fsb_net* net, const float* dlogits, const float* const* params, float* grads, cudaStream_t s);

extern "C" int fsb_net_backward(fsb_net* net, const float* dlogits, const float* const* params, float* grads,
                                void* workspace, size_t workspace_bytes, void* stream) {
    FSB_REQUIRE(net && dlogits && params && grads && workspace, "net_backward: null argument");
    if (!net->fwd_done || !net->training || net->ws != workspace) {
        set_error("net_backward: needs a preceding training-mode forward on the same workspace");
        return FSB_E_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    FSB_TRY(ensure_side_stream(net));
    KeyHash k;
    k.add(dlogits); k.add(grads); k.add(workspace); k.add(net->overlap); k.add(2);
    k.add(params, sizeof(float*) * fsb_net_num_params(net));
    return run_graphed(net, k.h, s, [&](cudaStream_t es) { return backward_impl(net, dlogits, params, grads, es); });
}

static int backward_impl(fsb_net* net, const float* dlogits, const float* const* params, float* grads, cudaStream_t s) {
    const fsb_net_config& c = net->cfg;
    // gradient planes carry hi + lo only when the backward GEMMs form three products; the forward activations read by
    // the weight-gradient GEMMs always have their hi plane
    const int prec = net->prec_b, fmt = act_fmt(prec);
    // Entry-conv dgrad of the 1D model keeps three products even in the mixed mode.  Its output du feeds a BatchNorm
    // WITHOUT activation, whose d(beta) = sum_r du[r] telescopes to border terms only (sum_r dz = 0 after the
    // BatchNorm above), while the half rounding errors of dz accumulate over ALL rows: relative error
    // ~2^-12 sqrt(rows / border rows) = 2^-12 sqrt(W / 2) in 1D (4e-2 measured at W = 3446), but only
    // 2^-12 sqrt(HW / 2(H + W)) <= 2^-12 * 5.7 in 2D.
    const int prec_e = (!c.two_d && prec == 2) ? 1 : prec, fmt_e = act_fmt(prec_e);
    const int n = net->N;
    auto G = [&](int index) { return grads + net->param_offset[index]; };
    // conv / linear biases that feed a batch-statistics BN have an analytically zero gradient
    FSB_CUDA(cudaMemsetAsync(grads, 0, (size_t)net->total_params * sizeof(float), s));
    FSB_CUDA(cudaMemsetAsync(net->gscale, 0, (size_t)c.num_blocks * B_PER_BLOCK * sizeof(unsigned), s));
    FSB_CUDA(cudaMemsetAsync(net->gscale_out, 0, ((size_t)c.num_blocks + 1) * sizeof(unsigned), s));
    FSB_TRY(ensure_side_stream(net));
    const bool overlap = net->overlap;
    cudaStream_t ws = overlap ? net->side : s;       // stream of the weight-gradient GEMMs
    net->fork_used = 0;
    // everything enqueued on `s` so far is visible to the side stream from here on
    auto fork = [&]() -> int {
        if (!overlap) return 0;
        if (net->fork_used == net->fork_events.size()) {
            cudaEvent_t e;
            FSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            net->fork_events.push_back(e);
        }
        cudaEvent_t e = net->fork_events[net->fork_used++];
        FSB_CUDA(cudaEventRecord(e, s));
        FSB_CUDA(cudaStreamWaitEvent(net->side, e, 0));
        return 0;
    };

    // ---- head
    const GradRef kNoGrad = grad_f32(nullptr);
    {
        const int hb = net->head_param_base;
        const float* const* P = params + hb;
        RUN(CAT_HEAD, 0, copy2d(dlogits, n, c.n_classes, c.n_classes, net->dzl, net->CsCls, s));
        RUN(CAT_HEAD, 0, colsum(dlogits, n, c.n_classes, c.n_classes, G(hb + H_L5_B), s));
        RUN(CAT_HEAD, 0, simt_wgrad(net->h1, net->dzl, G(hb + H_L5_W), net->wgrad_scratch, net->lin5, s));
        RUN(CAT_HEAD, 0, simt_skinny_dgrad(net->dzl, net->pk_l5, net->dh1, net->lin5, net->head_scratch, s));
        Dropout dr = {c.dropout_p, net->dropout_seed, net->d_seed};
        FSB_TRY(bn_backward(net, s, grad_f32(net->dh1), kNoGrad, net->z1h, nullptr, net->g_head, net->hbn2, P[H_PRELU],
                            kNoRes, dr, G(hb + H_BN2_W), G(hb + H_BN2_B), G(hb + H_PRELU), net->dz1h, FMT_F32, nullptr,
                            nullptr, nullptr, CAT_HEAD));
        RUN(CAT_HEAD, 0, simt_wgrad(net->h0, net->dz1h, G(hb + H_L1_W), net->wgrad_scratch, net->lin1, s));
        RUN(CAT_HEAD, 0, simt_skinny_dgrad(net->dz1h, net->pk_l1, net->dh0, net->lin1, net->head_scratch, s));
        FSB_TRY(bn_backward(net, s, grad_f32(net->dh0), kNoGrad, net->feats, nullptr, net->g_head, net->hbn0, nullptr,
                            kNoRes, kNoDrop, G(hb + H_BN0_W), G(hb + H_BN0_B), nullptr, net->dfeats, FMT_F32, nullptr,
                            nullptr, nullptr, CAT_HEAD));
    }

    // ---- conv blocks, last to first
    // Compact mode (mixed precision): every dgrad writes ONE half plane scaled by the Hoelder bound max|dZ| * L1(W);
    // dr0b and dzp are half planes scaled by the bounds their producers reduce; BatchNorm-backward recovers zhat and
    // the PReLU branch from the hi plane of the stored activation wherever that is well conditioned.
    const bool cmp = net->compact && prec == 2;
    // d_out of every block but the last is a scaled half plane too (2D, global-max heads): its producer (the next block's
    // input-BatchNorm backward) bounds it, plus max |dfeats| for the head gradient scattered into it afterwards
    unsigned* const DF = net->gscale_out + c.num_blocks;
    auto dout_half = [&](int k) { return cmp && c.two_d && c.aggregation == 0 && k >= 0 && k < c.num_blocks - 1; };
    if (cmp && c.two_d && c.aggregation == 0) RUN(CAT_HEAD, 0, absmax_bits(net->dfeats, (long long)n * net->Ds, DF, s));
    for (int k = c.num_blocks - 1; k >= 0; --k) {
        BlockPlan& B = net->blocks[k];
        const int pb = k * P_PER_BLOCK;
        const float* const* P = params + pb;
        unsigned* const GS = net->gscale + (size_t)k * B_PER_BLOCK;     // GradScale slots of dz3 / dz2 / dz1 / dzp (-> dzf); B_IN: dr0b
        const float* const L1 = net->wl1 + (size_t)k * 4;               // entry, conv1, conv2, conv3
        // gradient w.r.t. the input of a dgrad GEMM whose operand carries the GradScale of `slot`
        auto dgrad_out = [&](const void* p, const unsigned* slot, const float* l1) {
            return cmp ? grad_h16(p, slot, l1) : grad_f32((const float*)p);
        };
        if (k == c.num_blocks - 1) FSB_CUDA(cudaMemsetAsync(B.d_out, 0, (size_t)B.g.rows * B.g.Cs * 4, s));
        if (B.rnn_index >= 0) {
            const int rb = c.num_blocks * P_PER_BLOCK + B.rnn_index * R_PER_HEAD;
            float* GR[R_PER_HEAD];
            for (int i = 0; i < R_PER_HEAD; ++i) GR[i] = G(rb + i);
            RUN(CAT_HEAD, 0, rnn_head_backward(B.rnn, net->dfeats, net->Ds, B.head_off, params + rb, B.pk_rnn, GR, B.d_out,
                                               B.g, net->rnn_scratch, s));
        } else if (B.head_off >= 0) {
            if (dout_half(k))
                RUN(CAT_ELT_BWD, 0, gmax_backward_h16(net->dfeats, net->Ds, B.head_off, B.argrow, B.g, B.d_out,
                                                      net->gscale_out + k, s));
            else
                RUN(CAT_ELT_BWD, 0, gmax_backward(net->dfeats, net->Ds, B.head_off, B.argrow, B.g, B.d_out, s));
        }
        // out = prelu3(bn3(z3) + r0)
        Residual res = {B.zp, B.bn_a.scale, B.bn_a.shift, P[P_PRELUA]};
        const GradRef dout = dout_half(k) ? grad_h16(B.d_out, net->gscale_out + k) : grad_f32(B.d_out);
        if (cmp) {
            const BnCoef coef3 = B.bn3.coef(P[P_PRELU3]);
            RUN(CAT_ELT_BWD, 0, bn_res_bwd_compact_reduce(dout, B.z3, B.out, B.sign3, B.g, coef3, res, net->partials, s));
            RUN(CAT_ELT_BWD, 0, bn_bwd_finalize(net->partials, bn_res_bwd_compact_blocks(B.g), B.g.pixels, B.bn3.C, B.bn3.Cs,
                                                B.bn3.scale, G(pb + P_BN3_W), G(pb + P_BN3_B), G(pb + P_PRELU3), B.bn3.c1,
                                                B.bn3.c2, GS + B_3, GS + B_IN, nullptr, s));
            RUN(CAT_ELT_BWD, 0, bn_res_bwd_compact_apply(dout, B.z3, B.out, B.sign3, B.g, coef3, res, B.bn3.c1, B.bn3.c2, B.dz3,
                                                         GS + B_3, B.dr0b, GS + B_IN, s));
        } else {
            FSB_TRY(bn_backward(net, s, dout, kNoGrad, B.z3, nullptr, B.g, B.bn3, P[P_PRELU3], res, kNoDrop,
                                G(pb + P_BN3_W), G(pb + P_BN3_B), G(pb + P_PRELU3), B.dz3, fmt, B.dr0b, nullptr, GS + B_3,
                                CAT_ELT_BWD));
        }
        FSB_TRY(fork());
        RUN_S(ws, CAT_GEMM_WGRAD, conv_flops(B.c3, B.g),
              conv_gemm_wgrad(prec, B.a2, B.dz3, G(pb + P_C3_W), net->wgrad_scratch, B.c3, GS + B_3, ws));
        RUN(CAT_GEMM_DGRAD, conv_flops(B.c3, B.g),
            conv_gemm_dgrad(prec, B.dz3, B.pk3, B.da2, B.c3, GS + B_3, cmp ? L1 + 3 : nullptr, s));
        FSB_TRY(bn_backward(net, s, dgrad_out(B.da2, GS + B_3, L1 + 3), kNoGrad, B.z2, cmp ? B.a2 : nullptr, B.g, B.bn2,
                            P[P_PRELU2], kNoRes, kNoDrop, G(pb + P_BN2_W), G(pb + P_BN2_B), G(pb + P_PRELU2), B.dz2, fmt,
                            nullptr, nullptr, GS + B_2, CAT_ELT_BWD));
        FSB_TRY(fork());
        RUN_S(ws, CAT_GEMM_WGRAD, conv_flops(B.c2, B.g),
              conv_gemm_wgrad(prec, B.a1, B.dz2, G(pb + P_C2_W), net->wgrad_scratch, B.c2, GS + B_2, ws));
        RUN(CAT_GEMM_DGRAD, conv_flops(B.c2, B.g),
            conv_gemm_dgrad(prec, B.dz2, B.pk2, B.da1, B.c2, GS + B_2, cmp ? L1 + 2 : nullptr, s));
        FSB_TRY(bn_backward(net, s, dgrad_out(B.da1, GS + B_2, L1 + 2), kNoGrad, B.z1, cmp ? B.a1 : nullptr, B.g, B.bn1,
                            P[P_PRELU1], kNoRes, kNoDrop, G(pb + P_BN1_W), G(pb + P_BN1_B), G(pb + P_PRELU1), B.dz1, fmt,
                            nullptr, nullptr, GS + B_1, CAT_ELT_BWD));
        FSB_TRY(fork());
        RUN_S(ws, CAT_GEMM_WGRAD, conv_flops(B.c1, B.g),
              conv_gemm_wgrad(prec, B.r0, B.dz1, G(pb + P_C1_W), net->wgrad_scratch, B.c1, GS + B_1, ws));
        RUN(CAT_GEMM_DGRAD, conv_flops(B.c1, B.g),
            conv_gemm_dgrad(prec, B.dz1, B.pk1, B.dr0a, B.c1, GS + B_1, cmp ? L1 + 1 : nullptr, s));
        // r0 = prelu_a(bn_a(zp)) ; gradient = conv1 dgrad + residual branch.  dzp: float32, or (compact, 2D) a half
        // plane with the GradScale of slot B_A -- the same scale the routed dzf carries
        const bool dzp_half = cmp && c.two_d;
        const GradRef dr0b = cmp ? grad_h16(B.dr0b, GS + B_IN) : grad_f32((const float*)B.dr0b);
        FSB_TRY(bn_backward(net, s, dgrad_out(B.dr0a, GS + B_1, L1 + 1), dr0b, B.zp, cmp ? B.r0 : nullptr, B.g, B.bn_a,
                            P[P_PRELUA], kNoRes, kNoDrop, G(pb + P_BNA_W), G(pb + P_BNA_B), G(pb + P_PRELUA), B.dzp,
                            dzp_half ? FMT_H16 : FMT_F32, nullptr, nullptr, GS + B_A, CAT_ELT_BWD));
        const GradRef dzp = dzp_half ? grad_h16(B.dzp, GS + B_A) : grad_f32((const float*)B.dzp);
        if (c.two_d && k == 0) {
            if (prec != 0 && net->conv0_tc && conv0_tc_supported(B.g))
                RUN(CAT_CONV0, 2.0 * 2.0 * 2 * B.C * 9 * (double)n * c.n_features * net->frames,
                    conv0_tc_backward(prec, net->feat, n, c.n_features, net->frames, B.bn_in.scale, B.bn_in.shift,
                                      B.bn_in.mean, B.bn_in.invstd, P[P_CONV_W], B.dzp, dzp_half ? 1 : 0, net->conv0_amax,
                                      GS + B_A, B.g, G(pb + P_CONV_W), G(pb + P_CONV_B), G(pb + P_BNIN_W), G(pb + P_BNIN_B),
                                      net->conv0_scratch, s));
            else {
                FSB_REQUIRE(!dzp_half, "net_backward: the CUDA-core block-0 conv needs FSB200_COMPACT_BWD=0");
                RUN(CAT_CONV0, 2.0 * 2.0 * 2 * B.C * 9 * (double)n * c.n_features * net->frames,
                    conv0_backward(net->feat, n, c.n_features, net->frames, B.bn_in.scale, B.bn_in.shift, B.bn_in.mean,
                                   B.bn_in.invstd, P[P_CONV_W], P[P_CONV_B], (const float*)B.dzp, net->conv0_amax, B.g,
                                   G(pb + P_CONV_W), G(pb + P_CONV_B), G(pb + P_BNIN_W), G(pb + P_BNIN_B),
                                   net->conv0_scratch, s));
            }
        } else {
            if (cmp)
                RUN(CAT_ELT_BWD, 0, maxpool_backward_amax(dzp, B.g, B.pool_amax, B.g_full, c.two_d ? 2 : 1, B.dzf, fmt_e,
                                                          GS + B_A, s));
            else
                RUN(CAT_ELT_BWD, 0, maxpool_backward((const float*)B.dzp, B.g, B.zf, B.g_full, c.two_d ? 2 : 1, B.dzf, fmt_e,
                                                     GS + B_A, s));
            FSB_TRY(fork());
            RUN_S(ws, CAT_GEMM_WGRAD, conv_flops(B.entry, B.g_in),
                  conv_gemm_wgrad(prec, B.u, B.dzf, G(pb + P_CONV_W), net->wgrad_scratch, B.entry, GS + B_A, ws));
            // 1D: the entry dgrad keeps three products and a float32 output (d(beta_in) telescopes, see above)
            const bool du_half = cmp && c.two_d;
            RUN(CAT_GEMM_DGRAD, conv_flops(B.entry, B.g_in),
                conv_gemm_dgrad(prec_e, B.dzf, B.pk_entry, B.du, B.entry, GS + B_A, du_half ? L1 + 0 : nullptr, s));
            float* dprev = k > 0 ? net->blocks[k - 1].d_out : nullptr;
            const GradRef du = du_half ? grad_h16(B.du, GS + B_A, L1 + 0) : grad_f32((const float*)B.du);
            const bool prev_half = dout_half(k - 1);      // d_out of block k - 1: half plane bounded by this BN + the head scatter
            FSB_TRY(bn_backward(net, s, du, kNoGrad, B.x_in, du_half ? B.u : nullptr, B.g_in, B.bn_in, nullptr, kNoRes,
                                kNoDrop, G(pb + P_BNIN_W), G(pb + P_BNIN_B), nullptr, dprev, prev_half ? FMT_H16 : FMT_F32,
                                nullptr, nullptr, prev_half ? net->gscale_out + (k - 1) : nullptr, CAT_ELT_BWD,
                                prev_half ? DF : nullptr));
        }
    }
    if (overlap) {      // join: the caller's stream continues only after the last weight gradient has landed
        FSB_CUDA(cudaEventRecord(net->join_event, net->side));
        FSB_CUDA(cudaStreamWaitEvent(s, net->join_event, 0));
    }
    return 0;
}

// =================================================================================================
extern "C" int fsb_net_read_activation(fsb_net* net, int which, float* dst, long long cap, long long* numel,
                                       void* workspace, void* stream) {
    FSB_REQUIRE(net && dst && numel, "read_activation: null argument");
    if (!net->fwd_done || net->ws != workspace) {
        set_error("read_activation: no forward on this workspace");
        return FSB_E_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const fsb_net_config& c = net->cfg;
    if (which == 0) {
        long long cnt = (long long)net->N * c.n_features * net->frames;
        FSB_REQUIRE(cap >= cnt, "read_activation: destination too small");
        *numel = cnt;
        if (c.two_d) {
            FSB_CUDA(cudaMemcpyAsync(dst, net->feat, cnt * 4, cudaMemcpyDeviceToDevice, s));
            return 0;
        }
        return pf_to_nchw(net->blocks[0].x_in, net->blocks[0].g_in, dst, s);
    }
    if (which == 100) {
        long long cnt = (long long)net->N * net->D;
        FSB_REQUIRE(cap >= cnt, "read_activation: destination too small");
        *numel = cnt;
        return copy2d(net->feats, net->N, net->D, net->Ds, dst, net->D, s);
    }
    if (which >= 300) {
        // debugging taps (float32 back end only): 300 + 10*block + {0 zp, 1 r0, 2 z1, 3 dz1, 4 dr0a, 5 dr0b, 6 da1, 7 dzp}
        FSB_REQUIRE(c.precision == 0, "read_activation: internal taps need the float32 back end");
        int kb = (which - 300) / 10, j = (which - 300) % 10;
        FSB_REQUIRE(kb >= 0 && kb < c.num_blocks && j <= 7, "read_activation: unknown tap %d", which);
        const BlockPlan& Bk = net->blocks[kb];
        const float* src[8] = {Bk.zp, (const float*)Bk.r0, Bk.z1, (const float*)Bk.dz1, (const float*)Bk.dr0a,
                               (const float*)Bk.dr0b, (const float*)Bk.da1, (const float*)Bk.dzp};
        FSB_REQUIRE(src[j] != nullptr, "read_activation: tap %d not available (training only)", which);
        long long cnt = Bk.g.pixels * Bk.g.C;
        FSB_REQUIRE(cap >= cnt, "read_activation: destination too small");
        *numel = cnt;
        return pf_to_nchw(src[j], Bk.g, dst, s);
    }
    int k = which - 1;
    FSB_REQUIRE(k >= 0 && k < c.num_blocks, "read_activation: unknown tensor %d", which);
    const BlockPlan& B = net->blocks[k];
    long long cnt = B.g.pixels * B.g.C;
    FSB_REQUIRE(cap >= cnt, "read_activation: destination too small");
    *numel = cnt;
    return pf_to_nchw(B.out, B.g, dst, s);
}

extern "C" int fsb_net_set_graphs(fsb_net* net, int on) {
    FSB_REQUIRE(net, "set_graphs: null handle");
    net->graphs = on != 0;
    if (!net->graphs) drop_graphs(net);
    return 0;
}

extern "C" int fsb_net_set_overlap(fsb_net* net, int on) {
    FSB_REQUIRE(net, "set_overlap: null handle");
    net->overlap = on != 0;
    return 0;
}

extern "C" int fsb_net_set_profiling(fsb_net* net, int on) {
    net->profiling = on != 0;
    return 0;
}

extern "C" int fsb_net_get_timings(fsb_net* net, int cap, const char** names, float* ms, double* flops, int* count) {
    FSB_REQUIRE(cap >= CAT_COUNT, "get_timings: capacity must be >= %d", (int)CAT_COUNT);
    for (int i = 0; i < CAT_COUNT; ++i) { names[i] = kCatNames[i]; ms[i] = 0.f; flops[i] = 0.0; }
    for (const fsb_net::Rec& r : net->recs) {
        float t = 0.f;
        FSB_CUDA(cudaEventSynchronize(net->ev_pool[r.e1]));
        FSB_CUDA(cudaEventElapsedTime(&t, net->ev_pool[r.e0], net->ev_pool[r.e1]));
        ms[r.cat] += t;
        flops[r.cat] += r.flops;
    }
    *count = CAT_COUNT;
    return 0;
}

// =================================================================================================
// unit-level conv entry points (NCHW in / out) over the same PF pipeline
extern "C" size_t fsb_conv_workspace_bytes(int n, int cin, int cout, int h, int w, int kh, int kw) {
    Geo gi = make_geo(n, h, w, cin, kh == 3 ? 1 : 0, kw == 3 ? 1 : 0);
    Geo go = make_geo(n, h, w, cout, gi.padH, gi.padW);
    ConvGeom c = make_conv_geom(gi, cin, cout, kh, kw);
    size_t pk = std::max(packed_weight_bytes(0, c), packed_weight_bytes(1, c));
    size_t wg = std::max(wgrad_scratch_bytes(0, c), wgrad_scratch_bytes(1, c));     // the tensor-core sizes do not depend on the product count
    return 2 * plane_bytes(gi) + 2 * plane_bytes(go) + pk + wg + 4096;
}

static int conv_unit_setup(int n, int cin, int cout, int h, int w, int kh, int kw, Geo& gi, Geo& go, ConvGeom& c) {
    FSB_REQUIRE((kh == 1 || kh == 3) && (kw == 1 || kw == 3), "conv: kernel must be 1 or 3 per dimension");
    gi = make_geo(n, h, w, cin, kh == 3 ? 1 : 0, kw == 3 ? 1 : 0);
    go = make_geo(n, h, w, cout, gi.padH, gi.padW);
    c = make_conv_geom(gi, cin, cout, kh, kw);
    return 0;
}

extern "C" int fsb_conv_forward(const float* x, const float* w, const float* b, int n, int cin, int cout, int h,
                                int wd, int kh, int kw, int precision, float* y, void* workspace, size_t ws_bytes,
                                void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    Geo gi, go;
    ConvGeom c;
    FSB_REQUIRE(precision >= 0 && precision <= 3, "conv: precision must be 0..3");
    if (precision == 3) precision = 1;      // mixed: three-product forward
    FSB_TRY(conv_unit_setup(n, cin, cout, h, wd, kh, kw, gi, go, c));
    FSB_REQUIRE(ws_bytes >= fsb_conv_workspace_bytes(n, cin, cout, h, wd, kh, kw), "conv: workspace too small");
    FSB_TRY(fsb_device_ok());
    Bump bump{(char*)workspace, 0};
    void* A = bump.take_bytes(plane_bytes(gi));
    float* Z = (float*)bump.take_bytes(plane_bytes(go));
    void* pk = bump.take_bytes(packed_weight_bytes(precision, c));
    int fmt = act_fmt(precision);
    FSB_CUDA(cudaMemsetAsync(A, 0, plane_bytes(gi), s));
    FSB_TRY(nchw_to_pf(x, gi, A, fmt, s));
    FSB_TRY(pack_weights(precision, w, b, c, pk, s));
    FSB_TRY(conv_gemm_fwd(precision, A, pk, Z, c, nullptr, s));
    return pf_to_nchw(Z, go, y, s);
}

extern "C" int fsb_conv_backward(const float* x, const float* w, const float* dy, int n, int cin, int cout, int h,
                                 int wd, int kh, int kw, int precision, float* dx, float* dw, float* db,
                                 void* workspace, size_t ws_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    Geo gi, go;
    ConvGeom c;
    FSB_REQUIRE(precision >= 0 && precision <= 3, "conv: precision must be 0..3");
    if (precision == 3) precision = 2;      // mixed: single-pass backward
    FSB_TRY(conv_unit_setup(n, cin, cout, h, wd, kh, kw, gi, go, c));
    FSB_REQUIRE(ws_bytes >= fsb_conv_workspace_bytes(n, cin, cout, h, wd, kh, kw), "conv: workspace too small");
    FSB_TRY(fsb_device_ok());
    Bump bump{(char*)workspace, 0};
    void* A = bump.take_bytes(plane_bytes(gi));
    void* dZ = bump.take_bytes(plane_bytes(go));
    float* dA = (float*)bump.take_bytes(plane_bytes(gi));
    void* pk = bump.take_bytes(packed_weight_bytes(precision, c));
    void* scratch = bump.take_bytes(wgrad_scratch_bytes(precision, c));
    int fmt = act_fmt(precision);
    FSB_CUDA(cudaMemsetAsync(A, 0, plane_bytes(gi), s));
    FSB_CUDA(cudaMemsetAsync(dZ, 0, plane_bytes(go), s));
    FSB_TRY(nchw_to_pf(x, gi, A, fmt, s));
    FSB_TRY(nchw_to_pf(dy, go, dZ, fmt, s));
    FSB_TRY(pack_weights(precision, w, nullptr, c, pk, s));
    FSB_TRY(conv_gemm_dgrad(precision, dZ, pk, dA, c, nullptr, nullptr, s));
    FSB_TRY(pf_to_nchw(dA, gi, dx, s));
    FSB_TRY(conv_gemm_wgrad(precision, A, dZ, dw, scratch, c, nullptr, s));
    // bias gradient: sum of dy over (n, h, w) -- dy is NCHW here
    if (db) {
        float* dyf = (float*)dZ;
        if (fmt != FMT_F32) {
            FSB_CUDA(cudaMemsetAsync(dZ, 0, plane_bytes(go), s));
            FSB_TRY(nchw_to_pf(dy, go, dZ, FMT_F32, s));
        }
        FSB_TRY(colsum(dyf, go.rows, cout, go.Cs, db, s));
    }
    return 0;
}
