"""dict -> dict sample transforms with the reference's signatures (ops/transforms.py of the
reference).  The transforms on the accelerated path are implemented here (`AudioFeatures`,
`MixUp`, collate-side helpers); file / sox based augmentations (`LoadAudio`,
`AudioAugmentation`, `FlipAudio`, `ShuffleAudio`, `CutOut`, `SampleSegment`, `STFT`) are out of
scope and are forwarded to a reference checkout when one is on sys.path.
"""
import importlib.util
import os
import random

import numpy as np

from ops.audio import mix_audio_and_labels

SAMPLE_RATE = 44100


class Augmentation:
    """Base class: `Compose.switch_off_augmentations` sets `p = 0` on every instance."""
    pass


class MapLabels:

    def __init__(self, class_map, drop_raw=True):
        self.class_map = class_map

    def __call__(self, dataset, **inputs):
        labels = np.zeros(len(self.class_map), dtype=np.float32)
        for c in inputs["raw_labels"]:
            labels[self.class_map[c]] = 1.0
        transformed = dict(inputs)
        transformed["labels"] = labels
        transformed.pop("raw_labels")
        return transformed


class MixUp(Augmentation):
    """With probability p mix the sample with `dataset.random_clean_sample()` (reference :44-65)."""

    def __init__(self, p):
        self.p = p

    def __call__(self, dataset, **inputs):
        transformed = dict(inputs)
        if np.random.uniform() < self.p:
            partner = dataset.random_clean_sample()
            audio, labels = mix_audio_and_labels(
                inputs["audio"], partner["audio"], inputs["labels"], partner["labels"])
            transformed["audio"] = audio
            transformed["labels"] = labels
        return transformed


class AudioFeatures:
    """Emits `signal = audio[:, None]` (raw PCM): feature extraction itself happens inside the
    model on the GPU.  The reference additionally computes a scipy STFT for mel features and throws
    it away (reference :222-228); that dead work is skipped -- the emitted dict is identical."""

    eps = 1e-4

    def __init__(self, descriptor, verbose=True):
        name, *args = descriptor.split("_")
        self.feature_type = name
        if name == "stft":
            n_fft, hop_size = args
            self.n_fft, self.hop_size = int(n_fft), int(hop_size)
            self.n_features = self.n_fft // 2 + 1
            self.padding_value = 0.0
            if verbose:
                print("\nUsing STFT features with params:\n", "n_fft: {}, hop_size: {}".format(n_fft, hop_size))
        elif name == "mel":
            n_fft, hop_size, n_mel = args
            self.n_fft, self.hop_size, self.n_mel = int(n_fft), int(hop_size), int(n_mel)
            self.n_features = self.n_mel
            self.padding_value = 0.0
            if verbose:
                print("\nUsing mel features with params:\n",
                      "n_fft: {}, hop_size: {}, n_mel: {}".format(n_fft, hop_size, n_mel))
        elif name == "raw":
            self.n_features = 1
            self.padding_value = 0.0
            if verbose:
                print("\nUsing raw waveform features.")

    def __call__(self, dataset, **inputs):
        transformed = dict(inputs)
        if self.feature_type in ("stft", "mel", "raw"):
            transformed["signal"] = np.expand_dims(inputs["audio"], -1)
        return transformed


class SampleLongAudio:

    def __init__(self, max_length):
        self.max_length = max_length

    def __call__(self, dataset, **inputs):
        transformed = dict(inputs)
        if (inputs["audio"].size / inputs["sr"]) > self.max_length:
            max_length = self.max_length * inputs["sr"]
            start = np.random.randint(0, inputs["audio"].size - max_length)
            transformed["audio"] = inputs["audio"][start:start + max_length]
        return transformed


class OneOf:

    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, dataset, **inputs):
        return random.choice(self.transforms)(**inputs)


class DropFields:

    def __init__(self, fields):
        self.to_drop = fields

    def __call__(self, dataset, **inputs):
        return {name: value for name, value in inputs.items() if name not in self.to_drop}


class RenameFields:

    def __init__(self, mapping):
        self.mapping = mapping

    def __call__(self, dataset, **inputs):
        transformed = dict(inputs)
        for old, new in self.mapping.items():
            transformed[new] = transformed.pop(old)
        return transformed


class Compose:

    def __init__(self, transforms):
        self.transforms = transforms

    def switch_off_augmentations(self):
        for t in self.transforms:
            if isinstance(t, Augmentation) or type(t).__mro__[-2].__name__ == "Augmentation":
                t.p = 0.0

    def __call__(self, dataset=None, **inputs):
        for t in self.transforms:
            inputs = t(dataset=dataset, **inputs)
        return inputs


class Identity:

    def __call__(self, dataset=None, **inputs):
        return inputs


_FORWARDED = ("LoadAudio", "AudioAugmentation", "FlipAudio", "ShuffleAudio", "CutOut", "SampleSegment", "STFT")
_reference_module = None


def __getattr__(name):
    """Out-of-scope transforms come from the reference checkout (if any) later on sys.path."""
    global _reference_module
    if name not in _FORWARDED:
        raise AttributeError("module 'ops.transforms' has no attribute %r" % name)
    if _reference_module is None:
        import ops
        here = os.path.dirname(os.path.abspath(__file__))
        for directory in ops.__path__:
            candidate = os.path.join(directory, "transforms.py")
            if os.path.abspath(directory) != here and os.path.isfile(candidate):
                spec = importlib.util.spec_from_file_location("ops._reference_transforms", candidate)
                module = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(module)
                _reference_module = module
                break
        else:
            raise AttributeError(
                "ops.transforms.%s is a file/sox based augmentation outside the accelerated path; put a "
                "checkout of the reference after this package on sys.path to use it" % name)
    return getattr(_reference_module, name)
