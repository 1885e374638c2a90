"""Feature ops + metric with the reference's signatures (ops/utils.py of the reference).

`compute_torch_stft` runs the fused sm_100a feature kernel (no torch.stft / cuFFT);
`make_mel_filterbanks` restates `librosa.filters.mel` (librosa 0.6.3: Slaney scale, area
normalisation) because the matrix is a constructor-time constant of the model.
"""
import json

import numpy as np
import torch


def lwlrap(truth, scores):
    """Label-weighted label-ranking average precision (reference ops/utils.py:17-26).  Host version (sklearn); the
    training / evaluation loops use the device kernel behind `fsb200.runtime.DeviceLwlrap`.
    Scores are widened to float64 first: the scipy pinned by the reference (1.2.1) ranks in float64 whatever the
    input, current scipy ranks float32 scores in float32 and rounds every L/rank term to 2^-24."""
    from sklearn.metrics import label_ranking_average_precision_score
    truth, scores = np.asarray(truth), np.asarray(scores, dtype=np.float64)
    weight = np.sum(truth > 0, axis=1)
    keep = np.flatnonzero(weight > 0)
    return label_ranking_average_precision_score(truth[keep, :] > 0, scores[keep, :], sample_weight=weight[keep])


def load_json(file):
    with open(file, "r") as f:
        return json.load(f)


def get_class_names_from_classmap(classmap):
    inverse = {v: k for k, v in classmap.items()}
    return [inverse[label] for label in sorted(classmap.values())]


def _slaney_hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    lin = f / (200.0 / 3)
    log_region = f >= 1000.0
    out = lin.copy()
    out[log_region] = 15.0 + np.log(f[log_region] / 1000.0) / (np.log(6.4) / 27.0)
    return out


def _slaney_mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=np.float64))
    out = m * (200.0 / 3)
    log_region = m >= 15.0
    out[log_region] = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m[log_region] - 15.0))
    return out


def mel_matrix(sr, n_fft, n_mels, fmin, fmax=None):
    """float64 (n_mels, n_fft//2+1) triangular filters; edges equally spaced on the Slaney mel axis,
    each filter scaled by 2 / (f_right - f_left)."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    bins = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    lo, hi = _slaney_hz_to_mel(fmin)[0], _slaney_hz_to_mel(fmax)[0]
    edges = _slaney_mel_to_hz(np.linspace(lo, hi, n_mels + 2))
    left, centre, right = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    rising = (bins[None, :] - left) / (centre - left)
    falling = (right - bins[None, :]) / (right - centre)
    tri = np.clip(np.minimum(rising, falling), 0.0, None)
    return tri * (2.0 / (right - left))


def parse_features(descriptor):
    name, *args = descriptor.split("_")
    return name, [int(a) for a in args]


def make_mel_filterbanks(descriptor, sr=44100):
    """`"mel_<n_fft>_<hop>_<n_mel>"` -> float32 (n_mel, n_fft//2+1), fmin = 5 Hz (reference :85-99)."""
    _, (n_fft, hop_size, n_mel) = parse_features(descriptor)
    return mel_matrix(sr, n_fft, n_mel, fmin=5, fmax=None).astype(np.float32)


def is_mel(descriptor):
    return descriptor.startswith("mel")


def is_stft(descriptor):
    return descriptor.startswith("stft")


def compute_torch_stft(audio, descriptor):
    """audio (N, T) CUDA float32 -> |STFT| (N, n_fft//2+1, 1 + T//hop): centred, reflect padded,
    periodic Hann, one-sided (reference :110-127), computed by the fused feature kernel."""
    from fsb200.runtime import FeatureExtractor
    _, args = parse_features(descriptor)
    n_fft, hop_size = args[0], args[1]
    if not (isinstance(audio, torch.Tensor) and audio.is_cuda):
        raise RuntimeError("compute_torch_stft: audio must be a CUDA tensor (no CPU path in this package)")
    return FeatureExtractor(n_fft, hop_size, device=audio.device)(audio, 0)


def compute_log_features(audio, descriptor, filterbank=None):
    """Fused `log(FB @ |STFT| + 1e-4)` (mel_*) or `log(|STFT| + 1e-4)` (stft_*): what the models'
    forward computes at networks/classifiers.py:565-579."""
    from fsb200.runtime import FeatureExtractor
    name, args = parse_features(descriptor)
    if name == "mel":
        fb = make_mel_filterbanks(descriptor) if filterbank is None else filterbank
        return FeatureExtractor(args[0], args[1], filterbank=fb, device=audio.device)(audio, 2)
    return FeatureExtractor(args[0], args[1], device=audio.device)(audio, 1)
