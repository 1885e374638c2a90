"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32 / numpy) of the reference's hot
path: raw PCM -> STFT -> mel -> log -> frequency-encoded 2D CNN (and 1D CNN on raw STFT) ->
LSEP loss, plus Adam-amsgrad, the 1-cycle LR schedule, lwlrap, MixUp and the bucketing sampler.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker.  The product
(`freesound-classification_b200/`) never imports it and has no CPU fallback.

PARITY PIN: the reference ships no tests, golden vectors or fixtures, so
this restatement is pinned against outputs of the reference's own modules executed in the
build container through `oracle/reference_shim.py`; `oracle/make_golden.py` is the generating
script and `tests/golden/*.npz` the committed vectors (`tests/test_oracle_golden.py` checks
this file against them).  Every function cites the reference file:line it follows
(paths relative to the reference root).
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# feature descriptors  (ops/utils.py:85-107, ops/transforms.py:154-203)
# --------------------------------------------------------------------------------------
def parse_descriptor(descriptor):
    """`"mel_2048_1024_128"` -> ("mel", 2048, 1024, 128); `"stft_256_128"` -> ("stft", 256, 128, None)."""
    name, *args = descriptor.split("_")
    if name == "mel":
        n_fft, hop, n_mel = (int(a) for a in args)
        return name, n_fft, hop, n_mel
    if name == "stft":
        n_fft, hop = (int(a) for a in args[:2])
        return name, n_fft, hop, None
    return name, None, None, None


# --------------------------------------------------------------------------------------
# librosa 0.6.3 `filters.mel` (third-party, pinned at requirements.txt:35; call site
# ops/utils.py:94-97 with fmin=5, fmax=None, htk=False, norm=1)
# --------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz,
                    min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """Slaney-scale, area-normalised triangular filters, float64 `(n_mels, 1 + n_fft//2)`."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def make_mel_filterbanks(descriptor, sr=44100):
    """ops/utils.py:85-99: float32 `(n_mel, n_fft//2+1)`, fmin=5."""
    _, n_fft, _, n_mel = parse_descriptor(descriptor)
    return mel_filterbank(sr, n_fft, n_mel, fmin=5, fmax=None).astype(np.float32)


# --------------------------------------------------------------------------------------
# STFT magnitude  (ops/utils.py:110-127; torch.stft defaults of torch 1.0.1: center=True,
# pad_mode="reflect", normalized=False, onesided=True, periodic Hann)
# --------------------------------------------------------------------------------------
def stft_magnitude(audio, n_fft, hop):
    """audio `(N, T)` float32 -> `(N, n_fft//2+1, 1 + T//hop)` float32 magnitude.

    Restated with explicit reflect padding, framing and `torch.fft.rfft` (not `torch.stft`).
    """
    audio = torch.as_tensor(audio, dtype=torch.float32)
    n, t = audio.shape
    half = n_fft // 2
    assert t > half, "reflect padding needs T > n_fft/2"
    padded = F.pad(audio.unsqueeze(1), (half, half), mode="reflect").squeeze(1)
    frames = padded.unfold(-1, n_fft, hop)                       # (N, n_frames, n_fft)
    k = torch.arange(n_fft, dtype=torch.float64)
    window = (0.5 - 0.5 * torch.cos(2 * math.pi * k / n_fft)).to(torch.float32).to(audio.device)
    spec = torch.fft.rfft(frames * window, dim=-1)               # (N, n_frames, F)
    mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2)
    return mag.transpose(1, 2).contiguous()


def features(signal, descriptor, filterbank=None):
    """networks/classifiers.py:565-579 (2D) / :178-192 (1D): `(N,T,1)` PCM -> log features
    `(N, n_features, frames)`: `log(FB @ |STFT| + 1e-4)` for mel, `log(|STFT| + 1e-4)` for stft."""
    name, n_fft, hop, n_mel = parse_descriptor(descriptor)
    signal = torch.as_tensor(signal, dtype=torch.float32)
    if signal.dim() == 3:
        signal = signal.squeeze(-1)
    mag = stft_magnitude(signal, n_fft, hop)
    if name == "stft":
        return torch.log(mag + 1e-4)
    if filterbank is None:
        filterbank = torch.from_numpy(make_mel_filterbanks(descriptor))
    mel = torch.matmul(torch.as_tensor(filterbank, dtype=torch.float32).to(mag.device), mag)
    return torch.log(mel + 1e-4)


def scipy_style_stft(audio, window_size, hop_size, log=True, eps=1e-4):
    """ops/audio.py:10-19: `scipy.signal.stft(audio, nperseg=window_size, noverlap=hop_size)`
    -> magnitude (optionally log).  scipy semantics restated in numpy: periodic Hann window
    scaled by 1/sum(window), hop = nperseg - noverlap, zero boundary extension of nperseg/2 on
    both sides and zero padding of the tail to a whole number of hops."""
    x = np.asarray(audio, dtype=np.float64)
    nperseg, noverlap = int(window_size), int(hop_size)
    step = nperseg - noverlap
    k = np.arange(nperseg)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * k / nperseg)
    half = nperseg // 2
    x = np.concatenate([np.zeros(half), x, np.zeros(half)])
    nadd = (-(x.shape[-1] - nperseg) % step) % nperseg
    x = np.concatenate([x, np.zeros(nadd)])
    n_frames = (x.shape[-1] - noverlap) // step
    idx = np.arange(nperseg)[None, :] + step * np.arange(n_frames)[:, None]
    frames = x[idx] * win
    spec = np.fft.rfft(frames, axis=-1) / win.sum()
    s = np.abs(spec).T
    if log:
        s = np.log(s + eps)
    return s


# --------------------------------------------------------------------------------------
# network configuration helpers (networks/classifiers.py:502-513)
# --------------------------------------------------------------------------------------
def block_depths(num_conv_blocks, conv_base_depth, growth_rate):
    return [int(growth_rate ** k * conv_base_depth) for k in range(num_conv_blocks)]


def _bn(x, sd, prefix, training, stats_out=None):
    """BatchNorm (train: batch mean / biased var, eps 1e-5; eval: running stats)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        if stats_out is not None:
            dims = [0] + list(range(2, x.dim()))
            n = x.numel() // x.shape[1]
            mean = x.detach().mean(dims)
            var_unbiased = x.detach().var(dims, unbiased=True) if n > 1 else x.detach().var(dims, unbiased=False)
            stats_out[prefix] = (mean, var_unbiased)
        return F.batch_norm(x, None, None, w, b, True, 0.1, 1e-5)
    return F.batch_norm(x, rm, rv, w, b, False, 0.1, 1e-5)


def _resblock(x, sd, p, training, two_d, stats_out):
    """networks/classifiers.py:72-104 (2D) / :37-69 (1D)."""
    conv = F.conv2d if two_d else F.conv1d
    identity = x
    out = conv(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"])
    out = _bn(out, sd, p + ".bn1", training, stats_out)
    out = F.prelu(out, sd[p + ".prelu1.weight"])
    out = conv(out, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    out = _bn(out, sd, p + ".bn2", training, stats_out)
    out = F.prelu(out, sd[p + ".prelu2.weight"])
    out = conv(out, sd[p + ".conv3.weight"], sd[p + ".conv3.bias"])
    out = _bn(out, sd, p + ".bn3", training, stats_out)
    out = out + identity
    out = F.prelu(out, sd[p + ".prelu3.weight"])
    return out


def _head(feats, sd, training, dropout_p, stats_out, dropout_mask=None):
    """output_transform, networks/classifiers.py:542-549."""
    p = "output_transform"
    h = _bn(feats, sd, p + ".0", training, stats_out)
    h = F.linear(h, sd[p + ".1.weight"], sd[p + ".1.bias"])
    h = _bn(h, sd, p + ".2", training, stats_out)
    h = F.prelu(h, sd[p + ".3.weight"])
    if training and dropout_p > 0:
        if dropout_mask is not None:
            h = h * dropout_mask / (1.0 - dropout_p)
        else:
            h = F.dropout(h, dropout_p, True)
    return F.linear(h, sd[p + ".5.weight"], sd[p + ".5.bias"])


RNN_SIZE = 128      # networks/classifiers.py:509 (`rnn_size = 128`)


def _rnn_head(h, sd, prefix):
    """aggregation_type == "rnn" (networks/classifiers.py:514-522, 592-597): mean over the frequency axis,
    LayerNorm over channels, bidirectional GRU(C -> 128, batch_first) over time; the head feature is the pair of final
    hidden states [forward | backward] = `state.permute(1, 0, 2).view(N, -1)`.  GRU cell restated explicitly (torch
    gate order r, z, n): r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r * (W_hn h + b_hn)),
    h' = (1 - z) * n + z * h."""
    x = torch.mean(h, 2).permute(0, 2, 1)                                   # (N, W, C)
    x = F.layer_norm(x, (x.shape[-1],), sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], 1e-5)
    finals = []
    for suffix, reverse in (("", False), ("_reverse", True)):
        w_ih, w_hh = sd[prefix + ".1.weight_ih_l0" + suffix], sd[prefix + ".1.weight_hh_l0" + suffix]
        b_ih, b_hh = sd[prefix + ".1.bias_ih_l0" + suffix], sd[prefix + ".1.bias_hh_l0" + suffix]
        gi = F.linear(x.flip(1) if reverse else x, w_ih, b_ih)              # (N, W, 3 * 128)
        hcur = torch.zeros(x.shape[0], RNN_SIZE, dtype=x.dtype, device=x.device)
        for t in range(gi.shape[1]):
            gh = F.linear(hcur, w_hh, b_hh)
            i_r, i_z, i_n = gi[:, t].chunk(3, -1)
            h_r, h_z, h_n = gh.chunk(3, -1)
            r = torch.sigmoid(i_r + h_r)
            z = torch.sigmoid(i_z + h_z)
            cand = torch.tanh(i_n + r * h_n)
            hcur = (1 - z) * cand + z * hcur
        finals.append(hcur)
    return torch.cat(finals, -1)


def add_frequency_encoding(x):
    """networks/classifiers.py:553-561: concat channel `linspace(-1, 1, H)[h]`."""
    n, d, h, w = x.shape
    vertical = torch.linspace(-1, 1, h, dtype=x.dtype).to(x.device).view(1, 1, -1, 1).repeat(n, 1, 1, w)
    return torch.cat([x, vertical], dim=1)


def net2d_forward(sd, config, signal, training=False, stats_out=None, feats_in=None,
                  taps=None, dropout_mask=None):
    """TwoDimensionalCNNClassificationModel.forward, networks/classifiers.py:563-607
    (aggregation_type == "max").  `sd` maps state_dict keys to tensors (leaf tensors with
    requires_grad give gradients through plain autograd).  Returns logits `(N, C)`.
    `stats_out` (dict) receives per-BN batch (mean, unbiased var) in train mode."""
    net = config["network"]
    assert net["aggregation_type"] in ("max", "rnn")
    if feats_in is None:
        feats_in = features(signal, config["data"]["features"])
    h = add_frequency_encoding(feats_in.unsqueeze(1))
    if taps is not None:
        taps["input"] = h
    heads = []
    for k in range(net["num_conv_blocks"]):
        p = "conv_modules.%d" % k
        h = _bn(h, sd, p + ".0", training, stats_out)
        h = F.conv2d(h, sd[p + ".1.weight"], sd[p + ".1.bias"], padding=1)
        h = F.max_pool2d(h, kernel_size=2, stride=2)
        h = _bn(h, sd, p + ".3", training, stats_out)
        h = F.prelu(h, sd[p + ".4.weight"])
        h = _resblock(h, sd, p + ".5", training, True, stats_out)
        if taps is not None:
            taps["block%d" % k] = h
        if k >= net["start_deep_supervision_on"]:
            if net["aggregation_type"] == "rnn":
                heads.append(_rnn_head(h, sd, "rnns.%d" % (k - net["start_deep_supervision_on"])))
            else:
                heads.append(F.adaptive_max_pool2d(h, 1).squeeze(-1).squeeze(-1))
    feats = torch.cat(heads, -1)
    if taps is not None:
        taps["head_in"] = feats
    return _head(feats, sd, training, net["output_dropout"], stats_out, dropout_mask)


def net1d_forward(sd, config, signal, training=False, stats_out=None, feats_in=None,
                  taps=None, dropout_mask=None):
    """HierarchicalCNNClassificationModel.forward, networks/classifiers.py:176-217."""
    net = config["network"]
    assert net["aggregation_type"] == "max"
    if feats_in is None:
        feats_in = features(signal, config["data"]["features"])
    h = feats_in
    heads = []
    for k in range(net["num_conv_blocks"]):
        p = "conv_modules.%d" % k
        h = _bn(h, sd, p + ".0", training, stats_out)
        h = F.conv1d(h, sd[p + ".1.weight"], sd[p + ".1.bias"], padding=1)
        h = F.max_pool1d(h, kernel_size=2, stride=2)
        h = _bn(h, sd, p + ".3", training, stats_out)
        h = F.prelu(h, sd[p + ".4.weight"])
        h = _resblock(h, sd, p + ".5", training, False, stats_out)
        if taps is not None:
            taps["block%d" % k] = h
        if k >= net["start_deep_supervision_on"]:
            heads.append(F.adaptive_max_pool1d(h, 1).squeeze(-1))
    feats = torch.cat(heads, -1)
    if taps is not None:
        taps["head_in"] = feats
    return _head(feats, sd, training, net["output_dropout"], stats_out, dropout_mask)


def init_state_dict(config, two_d=True, seed=42):
    """Build the reference's module tree (same construction order => same default-init RNG
    consumption under `torch.manual_seed(seed)`, networks/classifiers.py:497-549 / :120-173)
    and return its state_dict."""
    import torch.nn as nn
    torch.manual_seed(seed)
    net, data = config["network"], config["data"]
    depths = block_depths(net["num_conv_blocks"], net["conv_base_depth"], net["growth_rate"])
    conv, bn = (nn.Conv2d, nn.BatchNorm2d) if two_d else (nn.Conv1d, nn.BatchNorm1d)
    pool = nn.MaxPool2d if two_d else nn.MaxPool1d

    class Res(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.conv1 = conv(d, d, kernel_size=1)
            self.bn1 = bn(d)
            self.conv2 = conv(d, d, kernel_size=3, padding=1)
            self.bn2 = bn(d)
            self.conv3 = conv(d, d, kernel_size=1)
            self.bn3 = bn(d)
            self.prelu1 = nn.PReLU(d)
            self.prelu2 = nn.PReLU(d)
            self.prelu3 = nn.PReLU(d)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv_modules = nn.ModuleList()
            self.rnns = nn.ModuleList()
            total = 0
            for k, d in enumerate(depths):
                cin = (2 if two_d else data["_input_dim"]) if k == 0 else depths[k - 1]
                if k >= net["start_deep_supervision_on"]:
                    if two_d and net["aggregation_type"] == "rnn":        # registered (and initialised) before the block
                        total += 2 * RNN_SIZE
                        self.rnns.append(nn.Sequential(
                            nn.LayerNorm((d,)), nn.GRU(d, RNN_SIZE, batch_first=True, bidirectional=True)))
                        continue_max = False
                    else:
                        continue_max = True
                    if continue_max:
                        total += d
                self.conv_modules.append(nn.Sequential(
                    bn(cin), conv(cin, d, kernel_size=3, padding=1), pool(kernel_size=2, stride=2),
                    bn(d), nn.PReLU(d), Res(d)))
            self.output_transform = nn.Sequential(
                nn.BatchNorm1d(total), nn.Linear(total, total), nn.BatchNorm1d(total),
                nn.PReLU(total), nn.Dropout(p=net["output_dropout"]),
                nn.Linear(total, data["_n_classes"]))

    return {k: v.clone() for k, v in Net().state_dict().items()}


# --------------------------------------------------------------------------------------
# LSEP  (networks/losses.py:47-58)
# --------------------------------------------------------------------------------------
def lsep_loss(input, target, average=True):
    """`log(1 + sum_{i,j: t_j < t_i} exp(s_j - s_i))` per sample (pairwise form, no max shift)."""
    diff = input.unsqueeze(1) - input.unsqueeze(2)            # [n, i, j] = s_j - s_i
    where = (target.unsqueeze(1) < target.unsqueeze(2)).to(input.dtype)
    lsep = torch.log(1 + (diff.exp() * where).sum(2).sum(1))
    return lsep.mean() if average else lsep


# --------------------------------------------------------------------------------------
# optimiser + schedule (ops/training.py:9-12,208-234; torch.optim.Adam amsgrad of torch 2.x,
# the oracle semantics fixed by SURVEY.md Appendix B)
# --------------------------------------------------------------------------------------
def onecycle_lr(step_index, min_lr, max_lr, max_steps):
    """LR set by the `step_index`-th call (0-based) of OneCycleScheduler.step()."""
    mid = int(round(max_steps * 0.3))
    if step_index < mid:
        r = step_index / mid
        return min_lr + r * (max_lr - min_lr)
    r = (step_index - mid) / (max_steps - mid)
    return max_lr + r * (min_lr / 1e3 - max_lr)


def adam_amsgrad_step(p, g, m, v, vmax, step, lr, beta1=0.9, beta2=0.999, eps=1e-8,
                      weight_decay=0.0):
    """One Adam(amsgrad=True) update, float32 numpy arrays updated in place; `step` is the
    1-based step count after increment."""
    f32 = np.float32
    if weight_decay != 0:
        g = g + f32(weight_decay) * p
    m[...] = m + (g - m) * f32(1 - beta1)                     # torch: lerp_(grad, 1-beta1)
    v[...] = v * f32(beta2) + g * g * f32(1 - beta2)
    np.maximum(vmax, v, out=vmax)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(vmax) / f32(math.sqrt(bc2)) + f32(eps)
    p[...] = p - f32(step_size) * (m / denom)
    return p


# --------------------------------------------------------------------------------------
# lwlrap (ops/utils.py:17-26 -> sklearn.metrics.label_ranking_average_precision_score with
# sample_weight = #positives; restated in numpy so the checker does not depend on sklearn)
# --------------------------------------------------------------------------------------
def lwlrap(truth, scores):
    truth = np.asarray(truth) > 0
    scores = np.asarray(scores, dtype=np.float64)
    total, weight = 0.0, 0.0
    for t, s in zip(truth, scores):
        npos = int(t.sum())
        if npos == 0:
            continue
        if npos == t.size:
            total += npos * 1.0
            weight += npos
            continue
        # rank with ties counted "max" (sklearn uses rankdata(-s, 'max'))
        pos = np.flatnonzero(t)
        rank_all = np.array([(s >= s[i]).sum() for i in pos], dtype=np.float64)
        rank_pos = np.array([(s[pos] >= s[i]).sum() for i in pos], dtype=np.float64)
        total += npos * float(np.mean(rank_pos / rank_all))
        weight += npos
    return total / weight


# --------------------------------------------------------------------------------------
# MixUp (ops/audio.py:32-52, equal- and unequal-length branches incl. the `=+` quirk)
# --------------------------------------------------------------------------------------
def mix_audio_and_labels(first_audio, second_audio, first_labels, second_labels, a=None, start=None):
    new_labels = np.clip(first_labels + second_labels, 0, 1)
    if a is None:
        a = np.random.uniform(0.4, 0.6)
    shorter, longer = first_audio, second_audio
    if shorter.size == longer.size:
        return (shorter + longer) / 2, new_labels
    if first_audio.size > second_audio.size:
        shorter, longer = longer, shorter
    if start is None:
        start = random.randint(0, longer.size - 1 - shorter.size)
    end = start + shorter.size
    longer = longer * a
    longer[start:end] = +shorter * (1 - a)      # reference assigns (`=+`), it does not add
    return longer, new_labels


# --------------------------------------------------------------------------------------
# collate + bucketing (ops/padding.py:8-32, :36-81)
# --------------------------------------------------------------------------------------
def pad_collate(signals, padding_value=0.0):
    """Right-pad `(T_i, 1)` arrays to the batch max with a constant; returns `(N, T_max, 1)`."""
    tmax = max(len(s) for s in signals)
    out = np.full((len(signals), tmax) + signals[0].shape[1:], padding_value, dtype=signals[0].dtype)
    for i, s in enumerate(signals):
        out[i, :len(s)] = s
    return out


def bucket_batches(lengths, max_batch_elems, buckets, rng=None):
    """BucketingSampler._create_batches with an explicit `random.Random` (None = no shuffle)."""
    lengths = np.asarray(lengths)
    binned = np.digitize(lengths, buckets)
    batches = []
    for bin_idx in range(1, len(buckets)):
        ids = list(np.nonzero(binned == bin_idx)[0])
        if rng is not None:
            rng.shuffle(ids)
        current_len, batch = 0, []
        for i in ids:
            if current_len < max_batch_elems:
                batch.append(int(i))
                current_len += lengths[i]
            else:
                batches.append(batch)
                current_len = lengths[i]
                batch = [int(i)]
        if batch:
            batches.append(batch)
    if rng is not None:
        rng.shuffle(batches)
    return batches


# --------------------------------------------------------------------------------------
# seeded synthetic inputs; shared by tests, bench and golden generation
# --------------------------------------------------------------------------------------
def synth_waveforms(n, t, seed=42, kind="structured", sr=44100):
    rng = np.random.RandomState(seed)
    if kind == "noise":
        return np.clip(0.1 * rng.randn(n, t), -1, 1).astype(np.float32)
    time = np.arange(t, dtype=np.float64) / sr
    out = np.zeros((n, t), dtype=np.float64)
    for i in range(n):
        for _ in range(rng.randint(3, 7)):
            f0 = np.exp(rng.uniform(np.log(50.0), np.log(16000.0)))
            f1 = f0 * np.exp(rng.uniform(-0.7, 0.7)) if rng.rand() < 0.5 else f0
            f1 = min(f1, 20000.0)
            amp = 10 ** rng.uniform(-1.5, -0.3)
            onset = rng.uniform(0, 0.8) * time[-1]
            dur = rng.uniform(0.1, 1.0) * time[-1]
            env = np.clip((time - onset) / 0.01, 0, 1) * np.exp(-np.maximum(time - onset, 0) / dur)
            phase = 2 * np.pi * (f0 * time + 0.5 * (f1 - f0) * time ** 2 / max(time[-1], 1e-9))
            out[i] += amp * env * np.sin(phase + rng.uniform(0, 2 * np.pi))
        out[i] += 10 ** (rng.uniform(-50, -30) / 20) * rng.randn(t)
    return np.clip(out, -1, 1).astype(np.float32)


def synth_labels(n, n_classes=80, seed=42):
    rng = np.random.RandomState(seed + 1)
    labels = np.zeros((n, n_classes), dtype=np.float32)
    for i in range(n):
        k = rng.randint(1, 4)
        labels[i, rng.choice(n_classes, size=k, replace=False)] = 1.0
    return labels
