"""Audio numpy ops with the reference's signatures (ops/audio.py of the reference)."""
import random

import numpy as np


def compute_stft(audio, window_size, hop_size, log=True, eps=1e-4):
    """`scipy.signal.stft(audio, nperseg=window_size, noverlap=hop_size)` magnitude (reference :10-19):
    periodic Hann scaled by 1/sum(window), hop = window_size - hop_size, zero boundary extension.
    numpy in / numpy out like the reference (it is a DataLoader-side transform), but the transform
    itself runs on the GPU through the fused feature kernel: the scipy framing of the zero-extended
    signal equals the kernel's centred framing shifted by window_size/hop frames."""
    import torch
    from fsb200.runtime import FeatureExtractor
    n = int(window_size)
    step = n - int(hop_size)
    if step <= 0 or n % step != 0 or n & (n - 1):
        raise ValueError("compute_stft: window_size must be a power of two and a multiple of the hop")
    x = np.asarray(audio, dtype=np.float32)
    half = n // 2
    ext = x.size + 2 * half
    nadd = (-(ext - n) % step) % n
    total = ext + nadd
    n_frames = (total - int(hop_size)) // step
    shift = n // step                        # kernel frame t' = t + n/step starts at sample t*step
    y = np.zeros(half + total + n, dtype=np.float32)
    y[2 * half:2 * half + x.size] = x
    mag = FeatureExtractor(n, step)(torch.from_numpy(y).cuda()[None], 0)[0]
    s = (mag[:, shift:shift + n_frames] / (n / 2.0)).cpu().numpy()
    if log:
        s = np.log(s + eps)
    return s


def trim_audio(audio):
    import librosa
    audio, interval = librosa.effects.trim(audio, top_db=60)
    return audio


def read_audio(file):
    import librosa
    audio, sr = librosa.load(file, sr=None)
    return audio, sr


def mix_audio_and_labels(first_audio, second_audio, first_labels, second_labels):
    """MixUp with OR-ed labels (reference :32-52), including the unequal-length branch whose
    `=+` ASSIGNS the scaled shorter clip into the longer one."""
    new_labels = np.clip(first_labels + second_labels, 0, 1)
    a = np.random.uniform(0.4, 0.6)
    shorter, longer = first_audio, second_audio
    if shorter.size == longer.size:
        return (shorter + longer) / 2, new_labels
    if first_audio.size > second_audio.size:
        shorter, longer = longer, shorter
    start = random.randint(0, longer.size - 1 - shorter.size)
    end = start + shorter.size
    longer *= a
    longer[start:end] = shorter * (1 - a)
    return longer, new_labels


def shuffle_audio(audio, chunk_length=0.5, sr=None):
    from sklearn.utils import gen_even_slices
    n_chunks = int((audio.size / sr) / chunk_length)
    if n_chunks in (0, 1):
        return audio
    slices = list(gen_even_slices(audio.size, n_chunks))
    random.shuffle(slices)
    return np.concatenate([audio[s] for s in slices])


def cutout(audio, area=0.25):
    area = int(audio.size * area)
    start = random.randrange(audio.size)
    audio[start:start + area] = 0
    return audio
