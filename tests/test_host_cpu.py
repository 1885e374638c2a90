"""CPU-only checks: the C-ABI library loads and exports every declared symbol, and the host-side
logic of the drop-in package (filterbank, schedules, collate/bucketing, MixUp, module tree) matches
the oracle / golden vectors.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import restate
from oracle.reference_shim import FakeExperiment, make_config


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def test_library_exports_every_declared_symbol():
    import fsb200
    header = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    handle = ctypes.CDLL(fsb200.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), "libfsb200.so does not export %s" % name
    assert sorted(fsb200.EXPORTS) == declared
    lib = fsb200.lib()
    assert lib.fsb_version() >= 100
    assert lib.fsb_adam_chunk() == 65536


def test_mel_matrix_matches_oracle_and_bands():
    from fsb200.runtime import band_filterbank
    from ops.utils import make_mel_filterbanks
    fb = make_mel_filterbanks("mel_2048_1024_128")
    ref = restate.make_mel_filterbanks("mel_2048_1024_128")
    assert fb.shape == (128, 1025) and fb.dtype == np.float32
    assert np.abs(fb - ref).max() < 1e-9
    vals, off, start, length = band_filterbank(fb)
    dense = np.zeros_like(fb)
    for m in range(128):
        dense[m, start[m]:start[m] + length[m]] = vals[off[m]:off[m] + length[m]]
    assert np.array_equal(dense, fb)
    assert len(vals) < 2200          # banded storage ~ nnz (2014), not 128 x 1025


def test_onecycle_and_make_scheduler():
    from ops.training import OneCycleScheduler, make_scheduler, make_step

    class Opt:
        param_groups = [dict(lr=0.1)]

    g = load("adam.npz")
    sched = make_scheduler("1cycle_0.0001_0.005", max_steps=10)(Opt())
    assert isinstance(sched, OneCycleScheduler)
    for i in range(10):
        make_step(sched, step=i + 1)
        assert abs(Opt.param_groups[0]["lr"] - g["lrs"][i]) < 1e-15


def test_collate_and_bucketing_match_reference():
    from ops.padding import BucketingSampler, make_collate_fn
    g = load("padding.npz")
    batch, off = [], 0
    for i, l in enumerate(g["lens"]):
        batch.append(dict(signal=g["raw"][off:off + l].reshape(l, 1).copy(), labels=np.float32([i])))
        off += l
    out = make_collate_fn({"signal": 0.0})(batch)
    assert np.array_equal(out["signal"].numpy(), g["collated"])

    class DS:
        lengths = g["ds_lengths"]

    random.seed(6)
    bs = BucketingSampler(DS(), max_batch_elems=64 * 441000 // 8,
                          buckets=[0, 5 * 44100, 10 * 44100, 20 * 44100, 31 * 44100])
    assert np.array_equal(np.array([len(b) for b in bs.batches]), g["batch_sizes"])
    assert np.array_equal(np.array([i for b in bs.batches for i in b]), g["batch_flat"])
    assert len(bs) == len(g["batch_sizes"])


def test_mixup_numpy_matches_reference():
    from ops.audio import mix_audio_and_labels
    g = load("mixup.npz")
    mixed, labels = mix_audio_and_labels(g["a1"].copy(), g["a2"].copy(), g["l1"], g["l2"])
    assert np.array_equal(mixed, g["mixed"]) and np.array_equal(labels, g["labels"])
    np.random.seed(9)
    random.seed(9)
    mixed_u, labels_u = mix_audio_and_labels(g["a1"].copy(), g["a3"].copy(), g["l1"], g["l2"])
    assert np.array_equal(mixed_u, g["mixed_unequal"]) and np.array_equal(labels_u, g["labels_unequal"])


def test_transforms_host_side():
    from ops.transforms import AudioFeatures, Compose, DropFields, MixUp, RenameFields
    af = AudioFeatures("mel_2048_1024_128", verbose=False)
    assert af.n_features == 128 and af.padding_value == 0.0
    assert AudioFeatures("stft_256_128", verbose=False).n_features == 129
    audio = np.arange(10, dtype=np.float32)
    t = Compose([af, RenameFields({"sr": "rate"}), DropFields(("audio",))])
    out = t(dataset=None, audio=audio, sr=44100)
    assert out["signal"].shape == (10, 1) and "audio" not in out and out["rate"] == 44100
    mix = MixUp(p=1.0)
    t2 = Compose([mix])
    t2.switch_off_augmentations()
    assert mix.p == 0.0


@pytest.mark.parametrize("name,cls,cfg", [
    ("net2d_small.npz", "TwoDimensionalCNNClassificationModel", dict(conv_base_depth=8, growth_rate=1.5)),
    ("net1d_small.npz", "HierarchicalCNNClassificationModel",
     dict(features="stft_256_128", conv_base_depth=8, growth_rate=1.5)),
])
def test_module_tree_matches_reference(name, cls, cfg):
    import networks.classifiers as nc
    from fsb200.runtime import canonical_bn_prefixes, canonical_param_names
    g = load(name)
    torch.manual_seed(42)
    model = getattr(nc, cls)(FakeExperiment(make_config(**cfg)), device="cpu")
    sd = model.state_dict()
    ref_keys = [k[3:] for k in g.files if k.startswith("sd/")]
    assert list(sd.keys()) == ref_keys                       # same keys, same order
    for k in ref_keys:
        assert tuple(sd[k].shape) == g["sd/" + k].shape, k
    chk = np.array([float(v.double().sum()) for k, v in sorted(sd.items())])
    assert np.array_equal(chk, g["init_checksum"])           # same default init under seed 42
    names = canonical_param_names(5)
    assert names == [n for n, _ in model.named_parameters()]
    assert len(canonical_bn_prefixes(5)) == 27
    assert "filterbanks" not in sd
    # no CPU fallback: forward on a CPU model must fail loudly, not compute
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 40000, 1))
    model.load_state_dict({k: torch.from_numpy(g["sd/" + k]) for k in ref_keys})


def test_rnn_aggregation_is_rejected():
    import networks.classifiers as nc
    with pytest.raises(NotImplementedError):
        nc.TwoDimensionalCNNClassificationModel(
            FakeExperiment(make_config(conv_base_depth=8, aggregation_type="rnn")), device="cpu")
