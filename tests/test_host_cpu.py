"""CPU-only checks: the C-ABI library loads and exports every declared symbol, and the host-side
logic of the drop-in package (filterbank, schedules, collate/bucketing, MixUp, module tree) matches
the oracle / golden vectors.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import random
import re
import sys

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
    assert lib.fsb_adam_chunk() == 8192


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
    from ops.transforms import AudioFeatures, Compose, MixUp, SampleLongAudio
    af = AudioFeatures("mel_2048_1024_128", verbose=False)
    assert af.n_features == 128 and af.padding_value == 0.0
    assert AudioFeatures("stft_256_128", verbose=False).n_features == 129
    audio = np.arange(10, dtype=np.float32)
    out = Compose([af])(dataset=None, audio=audio, sr=44100)
    assert out["signal"].shape == (10, 1) and np.array_equal(out["signal"][:, 0], audio) and out["sr"] == 44100
    mix = MixUp(p=1.0)
    t2 = Compose([mix])
    t2.switch_off_augmentations()
    assert mix.p == 0.0
    # SampleLongAudio: same draw as the reference (np.random.randint(0, size - window)), window of max_length * sr
    long_audio = np.arange(50, dtype=np.float32)
    np.random.seed(3)
    start = np.random.randint(0, 50 - 2 * 10)
    np.random.seed(3)
    cut = SampleLongAudio(max_length=2)(dataset=None, audio=long_audio, sr=10)
    assert np.array_equal(cut["audio"], long_audio[start:start + 20])
    assert SampleLongAudio(max_length=10)(dataset=None, audio=long_audio, sr=10)["audio"] is long_audio


def test_out_of_scope_transforms_forward_to_a_reference_checkout():
    """Names outside the accelerated path resolve to the reference's own classes when a checkout sits later on
    sys.path, and raise a clear AttributeError otherwise."""
    import ops.transforms as T
    from oracle.reference_shim import find_reference_root
    root = find_reference_root()
    if root is None:
        with pytest.raises(AttributeError):
            T.RenameFields
        return
    import ops
    saved, saved_mod = list(ops.__path__), T._reference_module
    try:
        ops.__path__.append(os.path.join(root, "ops"))
        T._reference_module = None
        try:
            renamed = T.RenameFields({"sr": "rate"})(dataset=None, sr=1, audio=2)
        except Exception as exc:       # the reference module imports librosa & co at module scope
            pytest.skip("reference ops/transforms.py is not importable here: %r" % (exc,))
        assert renamed == {"rate": 1, "audio": 2}
    finally:
        ops.__path__[:] = saved
        T._reference_module = saved_mod


@pytest.mark.parametrize("name,cls,cfg", [
    ("net2d_small.npz", "TwoDimensionalCNNClassificationModel", dict(conv_base_depth=8, growth_rate=1.5)),
    ("net1d_small.npz", "HierarchicalCNNClassificationModel",
     dict(features="stft_256_128", conv_base_depth=8, growth_rate=1.5)),
    ("net2d_rnn_small.npz", "TwoDimensionalCNNClassificationModel",
     dict(conv_base_depth=8, growth_rate=1.5, aggregation_type="rnn", start_deep_supervision_on=3)),
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
    names = canonical_param_names(5, len(model.rnns))
    assert names == [n for n, _ in model.named_parameters()]
    assert len(canonical_bn_prefixes(5)) == 27
    assert "filterbanks" not in sd
    # no CPU fallback: forward on a CPU model must fail loudly, not compute
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 40000, 1))
    model.load_state_dict({k: torch.from_numpy(g["sd/" + k]) for k in ref_keys})


def test_unsupported_aggregation_is_rejected():
    """rnn heads exist for the 2D model only (as in the reference); anything else must fail loudly."""
    import networks.classifiers as nc
    with pytest.raises(NotImplementedError):
        nc.HierarchicalCNNClassificationModel(
            FakeExperiment(make_config(features="stft_256_128", conv_base_depth=8, aggregation_type="rnn")), device="cpu")
    with pytest.raises(NotImplementedError):
        nc.TwoDimensionalCNNClassificationModel(
            FakeExperiment(make_config(conv_base_depth=8, aggregation_type="attention")), device="cpu")


def _dp_worker(rank, world_size, port, out_dir):
    """world-size-2 data-parallel step on CPU (gloo): per-rank oracle gradients on a batch shard, one flat all-reduce."""
    import torch.distributed as dist
    from fsb200 import dist as fdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert fdist.world() == (rank, world_size)
        cfg = make_config(conv_base_depth=4, growth_rate=1.5, num_conv_blocks=5)
        sd = restate.init_state_dict(cfg, two_d=True, seed=42)
        names = [k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
        params = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
        n, t = 4, 40000
        wav = torch.from_numpy(restate.synth_waveforms(n, t, seed=5))[..., None]
        labels = torch.from_numpy(restate.synth_labels(n, 80, seed=5))
        b, e = fdist.shard_range(n, rank, world_size)
        out = restate.net2d_forward(params, cfg, wav[b:e], training=True)
        restate.lsep_loss(out, labels[b:e], average=False).mean().backward()
        grads = [params[k].grad for k in names]
        local = torch.cat([g.reshape(-1) for g in grads]).clone()
        # path 1: gradients are views of one flat buffer -> a single in-place collective
        flat = local.clone()
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        scale = fdist.allreduce_gradients(views, flat=flat)
        # path 2: separately allocated gradients -> packed, reduced, copied back
        scale2 = fdist.allreduce_gradients(grads, flat=None)
        assert scale == scale2 == 1.0 / world_size
        assert torch.equal(torch.cat([g.reshape(-1) for g in grads]), flat)
        np.save(os.path.join(out_dir, "local_%d.npy" % rank), local.numpy())
        np.save(os.path.join(out_dir, "summed_%d.npy" % rank), flat.numpy())
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    from fsb200 import dist as fdist
    assert [fdist.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert fdist.shard_range(3, 1, 2) == (2, 3)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    l0, l1 = np.load(tmp_path / "local_0.npy"), np.load(tmp_path / "local_1.npy")
    s0, s1 = np.load(tmp_path / "summed_0.npy"), np.load(tmp_path / "summed_1.npy")
    assert np.array_equal(s0, s1)                       # every rank holds the same summed gradient
    assert np.allclose(s0, l0 + l1, rtol=0, atol=1e-6 * np.abs(s0).max())
    assert np.abs(l0 - l1).max() > 0                    # the shards really differed


def test_bucketed_inference_packing():
    from fsb200.inference import pack_batches, pad_batch
    rng = np.random.RandomState(3)
    lengths = rng.randint(1000, 9000, size=57)
    buckets = [2000, 4000, 6000, 8000]
    batches, dropped = pack_batches(lengths, buckets, max_batch_elems=15000)
    covered = sorted(i for b in batches for i in b)
    assert sorted(covered + dropped) == list(range(57))                  # every clip is placed or reported
    assert all(lengths[i] < 2000 or lengths[i] >= 8000 for i in dropped)
    for b in batches:
        bins = set(np.digitize(lengths[b], buckets))
        assert len(bins) == 1                                             # one length bucket per batch
        assert sum(lengths[i] for i in b[:-1]) < 15000 + max(lengths[b])  # closed by the first clip after the limit
    clips = [np.full(n, k, np.float32) for k, n in enumerate(lengths)]
    x = pad_batch(clips, batches[0])
    assert x.shape == (len(batches[0]), max(lengths[batches[0]]), 1) and x.dtype == np.float32
    for row, i in enumerate(batches[0]):
        assert (x[row, :lengths[i], 0] == i).all() and (x[row, lengths[i]:, 0] == 0).all()


@pytest.mark.parametrize("script", ["train_2d_cnn.py", "predict_2d_cnn.py", "evaluate_2d_cnn.py",
                                    "train_hierarchical_cnn.py", "finetune_hierarchical_cnn.py"])
def test_reference_entry_scripts_resolve_their_imports_against_this_package(script):
    """Drop-in boundary at import level: every name the reference's CLI scripts of the hot path import from `ops.*` and
    `networks.*` (e.g. train_2d_cnn.py:15-24, predict_2d_cnn.py:13-22) resolves with this package FIRST on sys.path --
    to the accelerated implementation where this package provides one, by forwarding to the reference checkout that
    follows on the path for host-side plumbing that is out of scope (third-party packages the reference imports at module
    scope and this container lacks -- pysndfx, librosa, iterstrat ... -- are stubbed as in oracle/reference_shim.py).
    (The scripts themselves cannot be executed here: they need `mag`, the competition data and a GPU, and the checkout
    does not exist on the GPU box.)"""
    import ast
    import importlib
    import types
    from oracle import reference_shim
    root = reference_shim.find_reference_root()
    if root is None or not os.path.isfile(os.path.join(root, script)):
        pytest.skip("no reference checkout reachable")
    tree = ast.parse(open(os.path.join(root, script)).read())
    wanted = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in ("ops", "networks"):
            wanted.extend((node.module, alias.name) for alias in node.names)
    assert wanted, "the script imports nothing from ops / networks?"
    import networks
    import ops
    import ops.transforms as T
    saved_ops, saved_net, saved_mod = list(ops.__path__), list(networks.__path__), T._reference_module
    saved_modules = dict(sys.modules)
    own, forwarded, missing = [], [], []
    try:
        reference_shim._install_stubs()
        for name in ("iterstrat", "iterstrat.ml_stratifiers", "scipy.io.wavfile", "soundfile", "pydub"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["iterstrat.ml_stratifiers"].MultilabelStratifiedKFold = object
        # the reference checkout FOLLOWS this package on the path, as in INTEGRATION.md level 1
        ops.__path__.append(os.path.join(root, "ops"))
        networks.__path__.append(os.path.join(root, "networks"))
        T._reference_module = None
        for module, name in wanted:
            try:
                mod = importlib.import_module(module)
                obj = getattr(mod, name)
            except Exception as exc:
                missing.append((module, name, repr(exc)))
                continue
            src = getattr(sys.modules.get(getattr(obj, "__module__", ""), None), "__file__", "") or getattr(mod, "__file__", "")
            (forwarded if src.startswith(root) else own).append((module, name))
    finally:
        ops.__path__[:] = saved_ops
        networks.__path__[:] = saved_net
        T._reference_module = saved_mod
        for k in [k for k in sys.modules if k not in saved_modules]:
            del sys.modules[k]
    assert not missing, missing
    # the hot-path names are this package's own implementations, never the checkout's
    must_own = {("networks.classifiers", "TwoDimensionalCNNClassificationModel"),
                ("networks.classifiers", "HierarchicalCNNClassificationModel"), ("ops.padding", "make_collate_fn"),
                ("ops.utils", "lwlrap"), ("ops.transforms", "AudioFeatures"), ("ops.transforms", "MixUp")}
    assert (must_own & set(wanted)) <= set(own), (sorted(must_own & set(wanted)), own)
    print("\n%s: %d names from this package, %d forwarded to the checkout" % (script, len(own), len(forwarded)))
