"""TEST INFRASTRUCTURE ONLY -- import shim that lets the *unmodified* reference modules
(`/root/reference/networks/classifiers.py`, `networks/losses.py`, `ops/utils.py`,
`ops/training.py`, `ops/padding.py`, `ops/audio.py`) import and run on this container's
torch 2.x CPU.  Used only by `oracle/make_golden.py` (golden-vector generation, in the
build container where `/root/reference` exists) and by `bench.py --impl reference` when a
copy of the reference travelled under `baseline/_ref`.  Never imported by the product.

Shim recipe follows SURVEY.md section 8(c):
  1. empty stub modules for tensorboardX / pretrainedmodels / umap / matplotlib / librosa /
     pysndfx (absent from the image; none of them does arithmetic on the hot path except
     `librosa.filters.mel`, see 2);
  2. `librosa.filters.mel` := `oracle.restate.mel_filterbank` (numpy float64 restatement of
     librosa 0.6.3, the version pinned in the reference's requirements.txt:35);
  3. `torch.stft` wrapped so that the torch-1.0.1 call in ops/utils.py:118-123 (no
     `return_complex`) yields the legacy real `(N, F, frames, 2)` layout it sums over.

Because the reference's packages are called `ops` and `networks` -- the same names the
product mirrors -- the shim loads them under a private prefix via a sys.path swap and
removes them from `sys.modules` afterwards.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_CANDIDATES = ("/root/reference",
                        os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                     "baseline", "_ref"))


def find_reference_root():
    for root in REFERENCE_CANDIDATES:
        if os.path.isfile(os.path.join(root, "networks", "classifiers.py")):
            return root
    return None


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def _install_stubs():
    from oracle import restate

    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    stub("tensorboardX", SummaryWriter=_Dummy)
    pm = stub("pretrainedmodels")
    pm.models = stub("pretrainedmodels.models", resnet18=None, resnet34=None)
    stub("umap")
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    stub("pysndfx")
    lib = stub("librosa")

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
        return restate.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)

    lib.filters = stub("librosa.filters", mel=mel)
    lib.effects = stub("librosa.effects")


_orig_stft = torch.stft


def _legacy_stft(input, n_fft, hop_length=None, win_length=None, window=None,
                 center=True, pad_mode="reflect", normalized=False, onesided=None,
                 return_complex=None):
    if return_complex is None:
        out = _orig_stft(input, n_fft, hop_length=hop_length, win_length=win_length,
                         window=window, center=center, pad_mode=pad_mode,
                         normalized=normalized, onesided=onesided, return_complex=True)
        return torch.view_as_real(out)
    return _orig_stft(input, n_fft, hop_length=hop_length, win_length=win_length,
                      window=window, center=center, pad_mode=pad_mode,
                      normalized=normalized, onesided=onesided,
                      return_complex=return_complex)


class ReferenceModules:
    """Namespace holding the reference's modules, loaded in isolation."""

    def __init__(self, root=None):
        root = root or find_reference_root()
        if root is None:
            raise RuntimeError("reference sources not found (looked in %s)"
                               % (REFERENCE_CANDIDATES,))
        self.root = root
        _install_stubs()
        saved = {k: v for k, v in sys.modules.items()
                 if k in ("ops", "networks", "datasets") or
                 k.startswith(("ops.", "networks.", "datasets."))}
        for k in saved:
            del sys.modules[k]
        sys.path.insert(0, root)
        torch.stft = _legacy_stft
        try:
            self.utils = importlib.import_module("ops.utils")
            self.training = importlib.import_module("ops.training")
            self.padding = importlib.import_module("ops.padding")
            self.audio = importlib.import_module("ops.audio")
            self.losses = importlib.import_module("networks.losses")
            self.classifiers = importlib.import_module("networks.classifiers")
        finally:
            sys.path.remove(root)
            for k in [k for k in sys.modules
                      if k in ("ops", "networks", "datasets") or
                      k.startswith(("ops.", "networks.", "datasets."))]:
                del sys.modules[k]
            sys.modules.update(saved)
        # torch.stft stays wrapped: the reference calls it at run time, and the wrapper is
        # transparent for callers that pass return_complex explicitly.


class AttrDict(dict):
    """`experiment.config` stand-in: nested dict with attribute access (what `mag` gives)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return AttrDict(v) if isinstance(v, dict) else v


class FakeExperiment:
    def __init__(self, config, root="/tmp/fsb200_experiment"):
        self.config = AttrDict(config)
        self.root = root
        self.checkpoints = os.path.join(root, "checkpoints")
        self.predictions = os.path.join(root, "predictions")

    def register_directory(self, name):
        path = os.path.join(self.root, name)
        os.makedirs(path, exist_ok=True)
        setattr(self, name, path)

    def register_result(self, *a, **k):
        pass


def make_config(features="mel_2048_1024_128", num_conv_blocks=5, conv_base_depth=100,
                growth_rate=1.5, start_deep_supervision_on=1, output_dropout=0.0,
                n_classes=80, input_dim=None, aggregation_type="max",
                scheduler="1cycle_0.0001_0.005", weight_decay=0.0, accumulation_steps=1,
                learning_rate=0.001, optimizer="adam"):
    if input_dim is None:
        parts = features.split("_")
        input_dim = int(parts[3]) if parts[0] == "mel" else int(parts[1]) // 2 + 1
    return dict(
        data=dict(features=features, _input_dim=input_dim, _n_classes=n_classes),
        network=dict(num_conv_blocks=num_conv_blocks,
                     start_deep_supervision_on=start_deep_supervision_on,
                     conv_base_depth=conv_base_depth, growth_rate=growth_rate,
                     output_dropout=output_dropout, aggregation_type=aggregation_type),
        train=dict(accumulation_steps=accumulation_steps, learning_rate=learning_rate,
                   optimizer=optimizer, scheduler=scheduler, weight_decay=weight_decay,
                   _save_every=1000, switch_off_augmentations_on=1000),
    )
