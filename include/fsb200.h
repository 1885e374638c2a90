/*
 * fsb200.h -- C ABI of libfsb200.so: the B200 (sm_100a) implementation of the one hot path of
 * ex4sperans/freesound-classification:
 *     raw PCM -> framed Hann rFFT -> mel -> log -> freq-encoded 2D CNN (or 1D CNN on raw STFT)
 *     forward / backward -> LSEP loss -> Adam-amsgrad, (+ MixUp batch assembly).
 *
 * The reference is pure Python on top of torch 1.0.1 (no FFI of its own), so every entry point
 * below replaces a *torch-op call site* of the reference; the file:line of that call site (relative
 * to the reference root) is cited on each declaration.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only: raw DEVICE pointers, explicit sizes/strides, `void* stream` = cudaStream_t;
 *   - every function returns 0 on success, a non-zero code otherwise (positive = cudaError_t,
 *     negative = FSB_E_*); nothing throws, nothing calls exit(); `fsb_last_error()` gives text;
 *   - calls only enqueue work: no device-memory allocation, no host synchronisation.  All device memory
 *     (incl. workspace, size from the *_workspace_bytes query) belongs to the caller.  Work is ordered
 *     on the caller's stream; fsb_net_forward / fsb_net_backward additionally fork part of it (weight
 *     packing, weight-gradient GEMMs) onto one internal low-priority stream per plan and join it back
 *     with events before they return control of the stream, so the caller sees plain stream semantics
 *     (FSB200_NO_OVERLAP=1 at plan creation, or fsb_net_set_overlap(net, 0), keeps everything on the caller's
 *     stream);
 *   - one host thread per device (one process per GPU under torchrun); handles are not thread-safe.
 */
#ifndef FSB200_H
#define FSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSB_E_INVALID   (-1)   /* bad argument / unsupported shape            */
#define FSB_E_WORKSPACE (-2)   /* workspace too small                         */
#define FSB_E_STATE     (-3)   /* call order violated (e.g. backward w/o fwd) */
#define FSB_E_NODEVICE  (-4)   /* no sm_100 device / driver entry point missing */

int         fsb_version(void);
const char* fsb_last_error(void);
/* 0 if the current device can run the library (compute capability 10.x), FSB_E_NODEVICE otherwise */
int         fsb_device_ok(void);

/* ------------------------------------------------------------------------------------------------
 * Feature extraction (K-feat).  Replaces ops/utils.py:110-127 (`compute_torch_stft`: torch.stft
 * center/reflect/periodic-Hann/onesided + magnitude), networks/classifiers.py:574-579 (mel
 * projection by `F.conv1d` with the librosa filterbank + `log(x + 1e-4)`) and :571-572
 * (`log(|STFT| + 1e-4)` for stft_* descriptors) with ONE fused kernel.
 *   pcm        (N, T) float32, row stride `pcm_stride` elements
 *   mode       0 = magnitude |STFT| (n_fft/2+1 rows), 1 = log(|STFT|+eps), 2 = log(mel+eps)
 *   fb_*       mode 2 only: banded filterbank -- row m covers bins [fb_start[m], fb_start[m]+fb_len[m])
 *              with weights fb_vals[fb_off[m] ...] (float32; built on the host from the dense matrix)
 *   out        element (n, f, t) at out[n*out_sn + f*out_sf + t*out_st]; frames = 1 + T/hop
 *   tables     device scratch of fsb_feat_table_bytes(n_fft) bytes, filled by fsb_feat_init_tables
 * ------------------------------------------------------------------------------------------------ */
size_t fsb_feat_table_bytes(int n_fft);
int    fsb_feat_init_tables(int n_fft, void* tables, void* stream);
int    fsb_feat_forward(const float* pcm, int n, long long pcm_stride, int t, int n_fft, int hop,
                        int mode, float eps, int n_mel, const float* fb_vals, const int* fb_off,
                        const int* fb_start, const int* fb_len, const void* tables, float* out,
                        long long out_sn, long long out_sf, long long out_st, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LSEP loss.  Replaces networks/losses.py:47-58 (6 elementwise torch kernels over (N,C,C)).
 *   scores, targets (N, C) float32 contiguous; loss (N) per-sample; general pairwise form
 *   (mask t_j < t_i), no max-shift, like the reference: a score gap above ~88 overflows float32.  The kernel then
 *   returns +inf where the reference's masked product yields NaN (inf * 0) -- both are non-finite, the value differs.
 *   backward: dscores[n,:] = dloss[n] * dL_n/ds
 * ------------------------------------------------------------------------------------------------ */
int fsb_lsep_forward(const float* scores, const float* targets, int n, int c, float* loss, void* stream);
int fsb_lsep_backward(const float* scores, const float* targets, const float* dloss, int n, int c,
                      float* dscores, void* stream);
/* networks/losses.py:25-44 (`lsep_loss_stable`): same loss shifted by max_{i,j}(s_j - s_i); finite for any gap */
int fsb_lsep_stable_forward(const float* scores, const float* targets, int n, int c, float* loss, void* stream);
int fsb_lsep_stable_backward(const float* scores, const float* targets, const float* dloss, int n, int c,
                             float* dscores, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Adam(amsgrad=True) multi-tensor step.  Replaces ops/training.py:10 (`torch.optim.Adam`,
 * per-tensor Python loop) -- torch 2.x semantics:
 *   g += wd*p; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g; vmax = max(vmax, v);
 *   p -= lr/(1-b1^t) * m / (sqrt(vmax)/sqrt(1-b2^t) + eps)
 * `table` is a DEVICE array of n_tensors records {p, g, m, v, vmax (float*), n (int64)} (6 x 8 bytes),
 * `block_map` a DEVICE int32 array of 2*n_blocks entries {tensor index, chunk index}; chunk = 8192
 * elements (fsb_adam_chunk()).  grad_scale multiplies g first (1/world_size after a SUM allreduce).
 * ------------------------------------------------------------------------------------------------ */
int fsb_adam_chunk(void);
int fsb_adam_amsgrad_step(const void* table, const int* block_map, int n_blocks, int step, float lr,
                          float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-device MixUp, equal-length branch.  Replaces ops/transforms.py:44-65 -> ops/audio.py:32-41
 * (`(a + b) / 2`, labels `clip(l1 + l2, 0, 1)`) for batches resident in HBM.
 *   partner[i] < 0  => sample i is left unmixed.
 * ------------------------------------------------------------------------------------------------ */
int fsb_mixup_equal(const float* pcm, const float* labels, const int* partner, int n, long long t,
                    int c, float* pcm_out, float* labels_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batch assembly over a device-resident PCM pool: SampleLongAudio crop (ops/transforms.py:292-309) ->
 * MixUp, both branches (ops/transforms.py:44-65, ops/audio.py:32-52) -> zero-pad collate
 * (ops/padding.py:8-32) in one pass.  Random draws are made by the caller (reference RNG order) and passed per
 * output row as 56-byte records (DEVICE array `rows`, n entries):
 *     int64  a_off, b_off     first pool sample of the cropped primary / partner clip
 *     double alpha, one_minus unequal lengths: scale of the longer clip, scale of the shorter clip (1 - alpha formed
 *                             in float64); both are rounded to float32 and multiplied in float32, as numpy does
 *     int32  a_len, b_len     lengths after cropping; b_len < 0 = row is not mixed
 *     int32  a_label, b_label rows of label_pool (n_clips, c)
 *     int32  mix_offset, pad  unequal lengths: start of the shorter clip inside the longer one
 * equal lengths -> (a + b) / 2; unequal -> alpha * longer with the window OVERWRITTEN by (1 - alpha) * shorter
 * (the reference's `=+` assigns); labels clip(l1 + l2, 0, 1).  out (n, t_out) is padded with pad_value.
 * ------------------------------------------------------------------------------------------------ */
int fsb_assemble_batch(const float* pool, const float* label_pool, const void* rows, int n, int c,
                       long long t_out, float pad_value, float* out, float* labels_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * lwlrap on the device.  Replaces ops/utils.py:17-26 (sklearn label_ranking_average_precision_score with
 * sample_weight = #positives, rows without positives dropped; ties share the worst rank).
 *   truth, scores (n, c) float32; scratch: fsb_lwlrap_scratch_bytes(n) bytes;
 *   out: 3 doubles {sum_rows sum_j L_j/rank_j, sum_rows #positives, their ratio = lwlrap}.  accumulate != 0 adds this
 *   batch to out[0], out[1] first (whole-set lwlrap over many batches, networks/classifiers.py:741-747).
 * ------------------------------------------------------------------------------------------------ */
size_t fsb_lwlrap_scratch_bytes(int n);
int    fsb_lwlrap(const float* truth, const float* scores, int n, int c, int accumulate, void* scratch,
                  double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-network plan (feature kernel + conv blocks + heads + FC head, forward and backward).
 * Replaces TwoDimensionalCNNClassificationModel.forward (networks/classifiers.py:563-607), its
 * autograd backward (:679) and HierarchicalCNNClassificationModel.forward (:176-217).
 *
 * fsb_net_config: plain-old-data description of the network (networks/classifiers.py:497-549).
 * ------------------------------------------------------------------------------------------------ */
#define FSB_MAX_BLOCKS 8

typedef struct fsb_net_config {
    int two_d;                 /* 1: TwoDimensionalCNN (freq-encoding channel), 0: Hierarchical 1D */
    int feat_mode;             /* 1 = log-STFT, 2 = log-mel (see fsb_feat_forward)                */
    int n_fft, hop;
    int n_features;            /* mel bins (2D: image height) or STFT bins (1D: input channels)   */
    int num_blocks;
    int depth[FSB_MAX_BLOCKS]; /* int(growth**k * base)                                            */
    int start_deep_supervision_on;
    int n_classes;
    float dropout_p;
    int precision;             /* GEMM back end of the conv layers:
                                  0 = fp32 CUDA-core GEMMs (cross-check path),
                                  1 = tcgen05, split-half operands x = hi + lo, three products, f32 accumulate
                                      (fp32-grade, forward AND backward),
                                  2 = tcgen05 single half pass everywhere (fast, logits NOT parity grade),
                                  3 = mixed (default of the Python layer): forward as 1, backward GEMMs (dgrad,
                                      wgrad) as 2 with per-tensor power-of-two gradient scaling              */
    int aggregation;           /* deep-supervision heads: 0 = global max (AdaptiveMaxPool), 1 = "rnn" (frequency mean ->
                                  LayerNorm -> bidirectional GRU(128) final states; 2D model only,
                                  networks/classifiers.py:514-522, 592-597)                                  */
} fsb_net_config;

typedef struct fsb_net fsb_net;   /* opaque */

/* filterbank arrays are HOST pointers here (copied into the plan's tables at first use) */
int    fsb_net_create(const fsb_net_config* cfg, const float* fb_vals, const int* fb_off,
                      const int* fb_start, const int* fb_len, int fb_nnz, fsb_net** out);
void   fsb_net_destroy(fsb_net* net);
/* number of parameter tensors / BN buffers the pointer tables must hold, in the canonical order
 * documented in DESIGN.md (== `named_parameters()` order of the reference module tree) */
int    fsb_net_num_params(const fsb_net* net);
int    fsb_net_num_bn(const fsb_net* net);
long long fsb_net_param_numel(const fsb_net* net, int index);
size_t fsb_net_workspace_bytes(const fsb_net* net, int n, int t, int training);

/*  signal     (N, T) float32 device, row stride `signal_stride`
 *  params     HOST array of fsb_net_num_params() device pointers (float32, torch layouts)
 *  bn_mean/bn_var  HOST arrays of fsb_net_num_bn() device pointers (running stats; updated in place
 *             when training != 0 with momentum 0.1 / unbiased variance), bn_count likewise (int64)
 *  training   0: running stats, no dropout, no tape; 1: batch stats, dropout(seed), tape kept in
 *             the workspace for fsb_net_backward
 *  logits     (N, n_classes) float32 device                                                       */
int fsb_net_forward(fsb_net* net, const float* signal, int n, int t, long long signal_stride,
                    const float* const* params, float* const* bn_mean, float* const* bn_var,
                    long long* const* bn_count, int training, unsigned long long dropout_seed,
                    void* workspace, size_t workspace_bytes, float* logits, void* stream);
/* Same forward pass on log features computed earlier with fsb_feat_forward (mode 1 / 2, layout (N, n_features, frames),
 * frames = 1 + t / hop): the feature kernel is skipped.  Lets an ensemble of fold models (predict_2d_cnn.py:111-118
 * averages 5 folds) share ONE feature extraction per batch -- the features do not depend on the weights.            */
int fsb_net_forward_features(fsb_net* net, const float* features, int n, int t, const float* const* params,
                             float* const* bn_running_mean, float* const* bn_running_var,
                             long long* const* bn_num_batches, int training, unsigned long long dropout_seed,
                             void* workspace, size_t workspace_bytes, float* logits, void* stream);

/*  dlogits (N, n_classes); grads: one flat float32 device buffer holding every parameter gradient
 *  back to back in canonical order (fully overwritten).                                           */
int fsb_net_backward(fsb_net* net, const float* dlogits, const float* const* params,
                     float* grads_flat, void* workspace, size_t workspace_bytes, void* stream);

/* debugging / parity taps: copy an internal activation of the last forward into an NCHW (2D) or NCW
 * (1D) float32 device buffer.  which: 0 = input features (N,F,frames) ; 1+k = output of block k;
 * 100 = concatenated head input (N, D).  Returns the element count through *numel.                */
int fsb_net_read_activation(fsb_net* net, int which, float* dst, long long dst_capacity,
                            long long* numel, void* workspace, void* stream);

/* per-phase device timings of the last forward+backward (CUDA events on the launch stream);
 * enable with fsb_net_set_profiling(net, 1).  names/ms are HOST arrays of capacity `cap`.         */
int fsb_net_set_profiling(fsb_net* net, int on);
/* side-stream overlap of weight packing / weight-gradient GEMMs (default on unless FSB200_NO_OVERLAP was set when the
 * plan was created); results are bit-identical either way */
int fsb_net_set_overlap(fsb_net* net, int on);
/* CUDA-graph replay (default on unless FSB200_GRAPHS=0 was set when the plan was created): the launch sequence of a
 * forward / backward call is captured the second time the same pointers and shapes are seen and replayed afterwards.
 * Results are bit-identical to eager launches.                                                                      */
int fsb_net_set_graphs(fsb_net* net, int on);
int fsb_net_get_timings(fsb_net* net, int cap, const char** names, float* ms, double* flops, int* count);
/* number of kernels this library launched since the counter was last reset (bench `gpu_launches`) */
long long fsb_launch_count(int reset);

/* ------------------------------------------------------------------------------------------------
 * Unit-level conv entry point used by the parity tests: y = conv(x, w) + b for NCHW float32
 * tensors through the same padded-flat NHWC pipeline and GEMM kernels the plan uses
 * (kh x kw in {1x1, 3x3, 1x3}, stride 1, "same" zero padding; networks/classifiers.py:526-531,
 * :75-80).  precision as in fsb_net_config.  grad entry point returns dx, dw, db for a given dy.
 * workspace >= fsb_conv_workspace_bytes(...).
 * ------------------------------------------------------------------------------------------------ */
size_t fsb_conv_workspace_bytes(int n, int cin, int cout, int h, int w, int kh, int kw);
int fsb_conv_forward(const float* x, const float* w, const float* b, int n, int cin, int cout, int h,
                     int wd, int kh, int kw, int precision, float* y, void* workspace,
                     size_t workspace_bytes, void* stream);
int fsb_conv_backward(const float* x, const float* w, const float* dy, int n, int cin, int cout, int h,
                      int wd, int kh, int kw, int precision, float* dx, float* dw, float* db,
                      void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_H */
