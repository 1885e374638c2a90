"""Host-side runtime over libfsb200: device-buffer ownership (torch tensors), stream hand-off and
autograd wiring.  PyTorch is plumbing here (memory, streams, autograd tape); all arithmetic is in
the CUDA library.  Nothing in this module computes on the CPU and nothing falls back."""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import NetConfig, check, lib

# GEMM back ends of the conv layers (fsb_net_config.precision).  "bf16x3" / "bf16" are the round-1 names of the
# three-product / single-pass modes (the operand format is IEEE half since round 2) and stay accepted.
PRECISIONS = {"fp32": 0, "fp16x3": 1, "fp16": 2, "mixed": 3, "bf16x3": 1, "bf16": 2}


def default_precision():
    """GEMM back end: `FSB200_PRECISION` =
    mixed  (default) tcgen05 tensor cores; forward with split-half operands and three products (fp32-grade: logits
           within 1e-3 of the reference), backward GEMMs single pass with per-tensor gradient scaling;
    fp16x3 three products in forward AND backward; fp32 CUDA-core GEMMs (cross-check); fp16 single pass everywhere."""
    return os.environ.get("FSB200_PRECISION", "mixed")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("%s must be a CUDA tensor: this package has no CPU path" % name)


# ----------------------------------------------------------------------------------------------
# filterbank: dense (n_mel, n_bins) -> banded (vals, off, start, len)
# ----------------------------------------------------------------------------------------------
def band_filterbank(fb):
    fb = np.asarray(fb, dtype=np.float32)
    vals, off, start, length = [], [], [], []
    for row in fb:
        nz = np.flatnonzero(row)
        if nz.size == 0:
            s, e = 0, 1
        else:
            s, e = int(nz[0]), int(nz[-1]) + 1
        off.append(len(vals))
        start.append(s)
        length.append(e - s)
        vals.extend(row[s:e].tolist())
    return (np.asarray(vals, np.float32), np.asarray(off, np.int32), np.asarray(start, np.int32),
            np.asarray(length, np.int32))


class FeatureExtractor:
    """Stand-alone K-feat launcher (used by `ops.utils.compute_torch_stft` and tests)."""

    _tables = {}

    def __init__(self, n_fft, hop, filterbank=None, device="cuda"):
        self.n_fft, self.hop = int(n_fft), int(hop)
        self.device = torch.device(device)
        key = (self.n_fft, str(self.device))
        if key not in FeatureExtractor._tables:
            nbytes = lib().fsb_feat_table_bytes(self.n_fft)
            tab = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                check(lib().fsb_feat_init_tables(self.n_fft, _ptr(tab), _stream()), "feat_init_tables")
            FeatureExtractor._tables[key] = tab
        self.tables = FeatureExtractor._tables[key]
        self.fb = None
        if filterbank is not None:
            vals, off, start, length = band_filterbank(filterbank)
            self.n_mel = len(off)
            self.fb = tuple(torch.from_numpy(a).to(self.device) for a in (vals, off, start, length))

    def __call__(self, audio, mode):
        """audio (N, T) float32 CUDA -> (N, F, frames); mode 0 |STFT|, 1 log|STFT|, 2 log-mel."""
        require_cuda(audio, "audio")
        if audio.dim() != 2:
            raise ValueError("audio must be (N, T)")
        audio = audio.float()
        if audio.stride(1) != 1:
            audio = audio.contiguous()
        n, t = audio.shape
        frames = 1 + t // self.hop
        f_out = self.n_mel if mode == 2 else self.n_fft // 2 + 1
        if mode == 2 and self.fb is None:
            raise ValueError("mel mode needs a filterbank")
        out = torch.empty((n, f_out, frames), dtype=torch.float32, device=audio.device)
        fb = self.fb if mode == 2 else (None, None, None, None)
        with torch.cuda.device(audio.device):
            check(lib().fsb_feat_forward(_ptr(audio), n, audio.stride(0), t, self.n_fft, self.hop, mode, 1e-4,
                                         self.n_mel if mode == 2 else 0, _ptr(fb[0]), _ptr(fb[1]), _ptr(fb[2]),
                                         _ptr(fb[3]), _ptr(self.tables), _ptr(out), f_out * frames, frames, 1,
                                         _stream()), "feat_forward")
        return out


# ----------------------------------------------------------------------------------------------
# LSEP
# ----------------------------------------------------------------------------------------------
class _LsepFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, targets, stable):
        require_cuda(scores, "input")
        s = scores.contiguous().float()
        t = targets.contiguous().float()
        n, c = s.shape
        loss = torch.empty(n, dtype=torch.float32, device=s.device)
        fn = lib().fsb_lsep_stable_forward if stable else lib().fsb_lsep_forward
        with torch.cuda.device(s.device):
            check(fn(_ptr(s), _ptr(t), n, c, _ptr(loss), _stream()), "lsep_forward")
        ctx.save_for_backward(s, t)
        ctx.stable = stable
        return loss

    @staticmethod
    def backward(ctx, dloss):
        s, t = ctx.saved_tensors
        n, c = s.shape
        dloss = dloss.contiguous().float()
        ds = torch.empty_like(s)
        fn = lib().fsb_lsep_stable_backward if ctx.stable else lib().fsb_lsep_backward
        with torch.cuda.device(s.device):
            check(fn(_ptr(s), _ptr(t), _ptr(dloss), n, c, _ptr(ds), _stream()), "lsep_backward")
        return ds, None, None


def lsep_per_sample(scores, targets, stable=False):
    return _LsepFunction.apply(scores, targets, stable)


# ----------------------------------------------------------------------------------------------
# lwlrap on the device
# ----------------------------------------------------------------------------------------------
class DeviceLwlrap:
    """Label-weighted label-ranking average precision without leaving the GPU (`fsb_lwlrap`).

    `update(truth, scores)` adds one batch to the running {numerator, weight}; `compute()` reads the whole-set value
    (one 24-byte device->host copy); `batch(truth, scores)` returns the per-batch value as a 0-dim float64 DEVICE
    tensor (no sync) -- `train_epoch` copies it to pinned memory asynchronously."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.total = torch.zeros(3, dtype=torch.float64, device=self.device)
        self._scratch = None

    def _run(self, truth, scores, out, accumulate):
        require_cuda(scores, "scores")
        t = truth.to(self.device).contiguous().float()
        s = scores.contiguous().float()
        n, c = s.shape
        need = lib().fsb_lwlrap_scratch_bytes(n)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().fsb_lwlrap(_ptr(t), _ptr(s), n, c, 1 if accumulate else 0, _ptr(self._scratch), _ptr(out),
                                   _stream()), "lwlrap")

    def reset(self):
        self.total.zero_()

    def update(self, truth, scores):
        self._run(truth, scores, self.total, True)

    def batch(self, truth, scores):
        out = torch.empty(3, dtype=torch.float64, device=self.device)
        self._run(truth, scores, out, False)
        return out[2]

    def compute(self):
        return float(self.total.cpu()[2])


# ----------------------------------------------------------------------------------------------
# network plan
# ----------------------------------------------------------------------------------------------
BLOCK_PARAM_SUFFIXES = (
    "0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias", "4.weight",
    "5.conv1.weight", "5.conv1.bias", "5.bn1.weight", "5.bn1.bias",
    "5.conv2.weight", "5.conv2.bias", "5.bn2.weight", "5.bn2.bias",
    "5.conv3.weight", "5.conv3.bias", "5.bn3.weight", "5.bn3.bias",
    "5.prelu1.weight", "5.prelu2.weight", "5.prelu3.weight")
HEAD_PARAM_NAMES = ("0.weight", "0.bias", "1.weight", "1.bias", "2.weight", "2.bias", "3.weight",
                    "5.weight", "5.bias")
BLOCK_BN_PREFIXES = ("0", "3", "5.bn1", "5.bn2", "5.bn3")
HEAD_BN_PREFIXES = ("0", "2")


RNN_PARAM_SUFFIXES = ("0.weight", "0.bias",
                      "1.weight_ih_l0", "1.weight_hh_l0", "1.bias_ih_l0", "1.bias_hh_l0",
                      "1.weight_ih_l0_reverse", "1.weight_hh_l0_reverse", "1.bias_ih_l0_reverse", "1.bias_hh_l0_reverse")


def canonical_param_names(num_blocks, n_rnn=0):
    """named_parameters() order of the reference's module tree: conv_modules, rnns, output_transform."""
    names = []
    for k in range(num_blocks):
        names += ["conv_modules.%d.%s" % (k, s) for s in BLOCK_PARAM_SUFFIXES]
    for i in range(n_rnn):
        names += ["rnns.%d.%s" % (i, s) for s in RNN_PARAM_SUFFIXES]
    names += ["output_transform.%s" % s for s in HEAD_PARAM_NAMES]
    return names


def canonical_bn_prefixes(num_blocks):
    names = []
    for k in range(num_blocks):
        names += ["conv_modules.%d.%s" % (k, s) for s in BLOCK_BN_PREFIXES]
    names += ["output_transform.%s" % s for s in HEAD_BN_PREFIXES]
    return names


class NetPlan:
    """Owns an `fsb_net` handle plus its workspace tensor."""

    def __init__(self, two_d, features, depths, start_deep_supervision_on, n_classes, dropout_p, filterbank=None,
                 precision=None, device="cuda", aggregation="max"):
        parts = features.split("_")
        if parts[0] not in ("mel", "stft"):
            raise ValueError("unsupported feature descriptor %r (mel_* / stft_* only)" % features)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("freesound-classification_b200 runs on CUDA (sm_100a) only; got device %r" % (device,))
        self.precision = precision or default_precision()
        cfg = NetConfig()
        cfg.two_d = 1 if two_d else 0
        cfg.feat_mode = 2 if parts[0] == "mel" else 1
        cfg.n_fft, cfg.hop = int(parts[1]), int(parts[2])
        cfg.n_features = int(parts[3]) if parts[0] == "mel" else cfg.n_fft // 2 + 1
        cfg.num_blocks = len(depths)
        for i, d in enumerate(depths):
            cfg.depth[i] = int(d)
        cfg.start_deep_supervision_on = int(start_deep_supervision_on)
        cfg.n_classes = int(n_classes)
        cfg.dropout_p = float(dropout_p)
        cfg.precision = PRECISIONS[self.precision]
        cfg.aggregation = {"max": 0, "rnn": 1}[aggregation]
        self.cfg = cfg
        handle = ctypes.c_void_p()
        if cfg.feat_mode == 2:
            vals, off, start, length = band_filterbank(filterbank)
            self._fb_keep = (vals, off, start, length)
            check(lib().fsb_net_create(ctypes.byref(cfg), vals.ctypes.data, off.ctypes.data, start.ctypes.data,
                                       length.ctypes.data, len(vals), ctypes.byref(handle)), "net_create")
        else:
            check(lib().fsb_net_create(ctypes.byref(cfg), None, None, None, None, 0, ctypes.byref(handle)),
                  "net_create")
        self.handle = handle
        self.num_params = lib().fsb_net_num_params(handle)
        self.num_bn = lib().fsb_net_num_bn(handle)
        self.param_numel = [lib().fsb_net_param_numel(handle, i) for i in range(self.num_params)]
        self.total_params = sum(self.param_numel)
        self.workspace = None
        self._ws_key = None
        self.forward_id = 0             # generation of the activations currently held by the workspace
        self.grad_flat = None
        self.grad_views = None
        self._param_ptrs = (ctypes.c_void_p * self.num_params)()
        self._mean_ptrs = (ctypes.c_void_p * self.num_bn)()
        self._var_ptrs = (ctypes.c_void_p * self.num_bn)()
        self._cnt_ptrs = (ctypes.c_void_p * self.num_bn)()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().fsb_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _ensure_workspace(self, n, t, training):
        key = (n, t, bool(training))
        if key != self._ws_key:
            need = lib().fsb_net_workspace_bytes(self.handle, n, t, 1 if training else 0)
            if need == 0:
                check(-1, "net_workspace_bytes")
            if self.workspace is None or self.workspace.numel() < need:
                self.workspace = None          # release before growing
                self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_key = key
        return self.workspace

    def set_pointers(self, params, bn_means, bn_vars, bn_counts):
        for i, p in enumerate(params):
            if p.numel() != self.param_numel[i] or not p.is_contiguous() or p.dtype != torch.float32:
                raise RuntimeError("parameter %d: expected %d contiguous float32 elements" % (i, self.param_numel[i]))
            self._param_ptrs[i] = p.data_ptr()
        for i in range(self.num_bn):
            self._mean_ptrs[i] = bn_means[i].data_ptr()
            self._var_ptrs[i] = bn_vars[i].data_ptr()
            self._cnt_ptrs[i] = bn_counts[i].data_ptr()

    def forward(self, signal, training, dropout_seed=0):
        """signal (N, T) float32 CUDA; pointers must have been set.  Returns logits (N, C)."""
        require_cuda(signal, "signal")
        if signal.stride(1) != 1 or signal.dtype != torch.float32:
            signal = signal.float().contiguous()
        n, t = signal.shape
        ws = self._ensure_workspace(n, t, training)
        self.forward_id += 1            # any earlier forward's activations are gone from here on
        logits = torch.empty((n, self.cfg.n_classes), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().fsb_net_forward(self.handle, _ptr(signal), n, t, signal.stride(0), self._param_ptrs,
                                        self._mean_ptrs, self._var_ptrs, self._cnt_ptrs, 1 if training else 0,
                                        int(dropout_seed) & 0xFFFFFFFFFFFFFFFF, _ptr(ws), ws.numel(), _ptr(logits),
                                        _stream()), "net_forward")
        return logits

    def forward_features(self, features, t, training=False, dropout_seed=0):
        """Forward pass on log features `(N, n_features, frames)` computed earlier (`FeatureExtractor`, mode 1 / 2) for
        clips of `t` samples: the feature kernel is skipped (fold ensembles share one extraction per batch)."""
        require_cuda(features, "features")
        features = features.float().contiguous()
        n = features.shape[0]
        frames = 1 + t // self.cfg.hop
        if tuple(features.shape) != (n, self.cfg.n_features, frames):
            raise ValueError("features must be (N, %d, %d) for clips of %d samples, got %r"
                             % (self.cfg.n_features, frames, t, tuple(features.shape)))
        ws = self._ensure_workspace(n, t, training)
        self.forward_id += 1
        logits = torch.empty((n, self.cfg.n_classes), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().fsb_net_forward_features(self.handle, _ptr(features), n, t, self._param_ptrs, self._mean_ptrs,
                                                 self._var_ptrs, self._cnt_ptrs, 1 if training else 0,
                                                 int(dropout_seed) & 0xFFFFFFFFFFFFFFFF, _ptr(ws), ws.numel(),
                                                 _ptr(logits), _stream()), "net_forward_features")
        return logits

    def backward(self, dlogits, fresh=False):
        """Runs the backward plan; returns the flat gradient (canonical parameter order).  The plan owns ONE
        persistent flat buffer (stable pointers keep the fused-Adam table and NCCL buffers cached); `fresh`
        asks for a temporary buffer instead (gradient accumulation into existing `.grad`)."""
        dlogits = dlogits.contiguous().float()
        if fresh:
            grads = torch.empty(self.total_params, dtype=torch.float32, device=self.device)
        else:
            if self.grad_flat is None:
                self.grad_flat = torch.empty(self.total_params, dtype=torch.float32, device=self.device)
                self.grad_views = None
            grads = self.grad_flat
        ws = self.workspace
        with torch.cuda.device(self.device):
            check(lib().fsb_net_backward(self.handle, _ptr(dlogits), self._param_ptrs, _ptr(grads), _ptr(ws),
                                         ws.numel(), _stream()), "net_backward")
        return grads

    def read_activation(self, which, shape):
        dst = torch.empty(shape, dtype=torch.float32, device=self.device)
        numel = ctypes.c_longlong(0)
        with torch.cuda.device(self.device):
            check(lib().fsb_net_read_activation(self.handle, which, _ptr(dst), dst.numel(), ctypes.byref(numel),
                                                _ptr(self.workspace), _stream()), "net_read_activation")
        if numel.value != dst.numel():
            raise RuntimeError("activation %d has %d elements, expected shape %r" % (which, numel.value, shape))
        return dst

    def set_profiling(self, on):
        lib().fsb_net_set_profiling(self.handle, 1 if on else 0)

    def set_graphs(self, on):
        check(lib().fsb_net_set_graphs(self.handle, 1 if on else 0), "net_set_graphs")

    def set_overlap(self, on):
        check(lib().fsb_net_set_overlap(self.handle, 1 if on else 0), "net_set_overlap")

    def timings(self):
        cap = 16
        names = (ctypes.c_char_p * cap)()
        ms = (ctypes.c_float * cap)()
        flops = (ctypes.c_double * cap)()
        count = ctypes.c_int(0)
        check(lib().fsb_net_get_timings(self.handle, cap, names, ms, flops, ctypes.byref(count)), "net_get_timings")
        return {names[i].decode(): (ms[i], flops[i]) for i in range(count.value)}


class _NetFunction(torch.autograd.Function):
    """logits = net(signal; params): the whole forward/backward is two library calls.

    backward() writes every parameter gradient into the plan's persistent flat buffer and installs views of
    it as `p.grad` directly (autograd's AccumulateGrad would clone each of the ~120 views); when a parameter
    already holds a gradient (accumulation_steps > 1) the step's gradient goes to a temporary buffer and
    is added.  Consequences, all checked here rather than silently wrong:
      * the activations live in ONE plan-owned workspace, so only the most recent training forward can be
        differentiated -- a backward through an older forward raises;
      * parameter gradients are delivered through `.grad`, not through autograd's return values
        (`torch.autograd.grad(..., params)` is not supported); parameters with `requires_grad=False` are skipped."""

    @staticmethod
    def forward(ctx, plan, signal, dropout_seed, *params):
        ctx.plan = plan
        ctx.params = params
        out = plan.forward(signal, True, dropout_seed)
        ctx.forward_id = plan.forward_id
        return out

    @staticmethod
    def backward(ctx, dlogits):
        plan, params = ctx.plan, ctx.params
        if ctx.forward_id != plan.forward_id:
            raise RuntimeError(
                "backward through a stale forward: the plan keeps the activations of its most recent forward only "
                "(forward #%d was overwritten by #%d); call backward before the next forward of this model"
                % (ctx.forward_id, plan.forward_id))
        wanted = [p.requires_grad for p in params]
        accumulate = any(p.grad is not None for p, w in zip(params, wanted) if w)
        flat = plan.backward(dlogits, fresh=accumulate)
        if accumulate:
            off = 0
            for p, n, w in zip(params, plan.param_numel, wanted):
                if w:
                    g = flat[off:off + n].view(p.shape)
                    if p.grad is None:
                        p.grad = g.clone()
                    else:
                        p.grad.add_(g)
                off += n
        else:
            if plan.grad_views is None or len(plan.grad_views) != len(params):
                views, off = [], 0
                for p, n in zip(params, plan.param_numel):
                    views.append(flat[off:off + n].view(p.shape))
                    off += n
                plan.grad_views = views
            for p, g, w in zip(params, plan.grad_views, wanted):
                if w:
                    p.grad = g
            plan.last_flat_grad = flat
        return (None, None, None) + (None,) * len(params)


def net_apply(plan, signal, dropout_seed, params):
    return _NetFunction.apply(plan, signal, dropout_seed, *params)


# ----------------------------------------------------------------------------------------------
# unit-level conv (tests)
# ----------------------------------------------------------------------------------------------
def conv_forward(x, w, b, precision="fp32"):
    require_cuda(x, "x")
    n, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    nbytes = lib().fsb_conv_workspace_bytes(n, cin, cout, h, wd, kh, kw)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    y = torch.empty((n, cout, h, wd), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().fsb_conv_forward(_ptr(x.contiguous()), _ptr(w.contiguous()), _ptr(b), n, cin, cout, h, wd, kh, kw,
                                     PRECISIONS[precision], _ptr(y), _ptr(ws), nbytes, _stream()), "conv_forward")
    return y


def conv_backward(x, w, dy, precision="fp32"):
    n, cin, h, wd = x.shape
    cout, _, kh, kw = w.shape
    nbytes = lib().fsb_conv_workspace_bytes(n, cin, cout, h, wd, kh, kw)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    dx = torch.empty_like(x)
    dw = torch.empty_like(w)
    db = torch.empty(cout, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().fsb_conv_backward(_ptr(x.contiguous()), _ptr(w.contiguous()), _ptr(dy.contiguous()), n, cin, cout,
                                      h, wd, kh, kw, PRECISIONS[precision], _ptr(dx), _ptr(dw), _ptr(db), _ptr(ws),
                                      nbytes, _stream()), "conv_backward")
    return dx, dw, db


def mixup_equal(pcm, labels, partner):
    """On-device MixUp (equal-length branch): pcm (N, T), labels (N, C), partner (N) int32 (-1 = keep)."""
    require_cuda(pcm, "pcm")
    pcm = pcm.contiguous().float()
    labels = labels.contiguous().float()
    partner = partner.to(device=pcm.device, dtype=torch.int32).contiguous()
    out = torch.empty_like(pcm)
    lout = torch.empty_like(labels)
    with torch.cuda.device(pcm.device):
        check(lib().fsb_mixup_equal(_ptr(pcm), _ptr(labels), _ptr(partner), pcm.shape[0], pcm.shape[1],
                                    labels.shape[1], _ptr(out), _ptr(lout), _stream()), "mixup_equal")
    return out, lout


def launch_count(reset=False):
    return lib().fsb_launch_count(1 if reset else 0)
