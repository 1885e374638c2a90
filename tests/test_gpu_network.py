"""GPU parity of the whole network (forward, backward, running statistics, training loop) against the
golden vectors produced by the reference and against the CPU oracle at canonical width."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import restate
from oracle.reference_shim import FakeExperiment, make_config

pytestmark = pytest.mark.gpu

PRECISIONS = ["fp32", "fp16x3", "mixed"]


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def build(cls_name, cfg, sd=None, precision="fp32", tmp=None):
    import networks.classifiers as nc
    os.environ["FSB200_PRECISION"] = precision
    torch.manual_seed(42)
    model = getattr(nc, cls_name)(FakeExperiment(make_config(**cfg), root=tmp or "/tmp/fsb200_exp"), device="cuda")
    if sd is not None:
        model.load_state_dict(sd)
    return model


CASES = [
    ("net2d_small.npz", "TwoDimensionalCNNClassificationModel", dict(conv_base_depth=8, growth_rate=1.5)),
    ("net2d_pow2.npz", "TwoDimensionalCNNClassificationModel",
     dict(conv_base_depth=8, growth_rate=2.0, start_deep_supervision_on=2)),
    ("net1d_small.npz", "HierarchicalCNNClassificationModel",
     dict(features="stft_256_128", conv_base_depth=8, growth_rate=1.5)),
    # aggregation_type="rnn": LayerNorm + bidirectional GRU(128) heads (networks/classifiers.py:514-522, 592-597)
    ("net2d_rnn_small.npz", "TwoDimensionalCNNClassificationModel",
     dict(conv_base_depth=8, growth_rate=1.5, aggregation_type="rnn", start_deep_supervision_on=3)),
]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name,cls,cfg", CASES)
def test_network_matches_reference_golden(name, cls, cfg, precision):
    from networks.losses import lsep_loss
    from ops.utils import lwlrap
    g = load(name)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    model = build(cls, cfg, sd, precision)
    wav = restate.synth_waveforms(int(g["n"]), int(g["t"]), seed=int(g["seed"]), kind=str(g["wave_kind"]))
    labels_np = restate.synth_labels(int(g["n"]), 80, seed=int(g["seed"]))
    signal = torch.from_numpy(wav)[..., None].cuda()
    labels = torch.from_numpy(labels_np).cuda()

    # ---- eval mode (running statistics): tight gate
    model.eval()
    with torch.no_grad():
        le = model(signal)["class_logits"].cpu().numpy()
    assert rel_err(le, g["logits_eval"]) < 1e-3            # north-star tolerance
    assert rel_err(le, g["logits_eval"]) < (2e-5 if precision == "fp32" else 3e-4)
    assert np.array_equal(np.argsort(-le, 1)[:, :3], np.argsort(-g["logits_eval"], 1)[:, :3])   # top-3 labels

    # ---- train mode: batch statistics over a 3-4 clip batch (ill-conditioned; 1e-3 gate, see oracle test)
    model.train()
    lt = model(signal)["class_logits"]
    assert rel_err(lt.detach().cpu().numpy(), g["logits_train"]) < 1e-3
    per = lsep_loss(lt, labels, average=False)
    assert rel_err(per.detach().cpu().numpy(), g["per_sample_loss"]) < 1e-3
    per.mean().backward()
    gmax = max(np.abs(g[k]).max() for k in g.files if k.startswith("grad/"))
    for k, p in model.named_parameters():
        ref = g["grad/" + k]
        got = p.grad.cpu().numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max() + 1e-3 * gmax, k
    after = model.state_dict()
    for k in g.files:
        if k.startswith("sd_after/"):
            ref = g[k]
            got = after[k[9:]].cpu().numpy()
            if "num_batches" in k:
                assert np.array_equal(got, ref), k
            else:
                assert rel_err(got, ref) < 1e-4, k
    got_lw = lwlrap(labels_np, torch.sigmoid(lt).detach().cpu().numpy())
    assert round(got_lw, 4) == round(float(g["lwlrap_train"]), 4)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_canonical_width_against_oracle(precision):
    """Canonical 2D config (base 100, growth 1.5, 5 blocks, 80 classes) on 8 x 1.5 s clips."""
    cfg = dict()
    config = make_config(**cfg)
    n, t = 8, 66150
    model = build("TwoDimensionalCNNClassificationModel", cfg, None, precision)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    wav = restate.synth_waveforms(n, t, seed=21)
    labels_np = restate.synth_labels(n, 80, seed=21)
    signal = torch.from_numpy(wav)[..., None]
    taps = {}
    with torch.no_grad():
        ref_eval = restate.net2d_forward(sd, config, signal, training=False, taps=taps)
    model.eval()
    with torch.no_grad():
        got_eval = model(signal.cuda())["class_logits"].cpu()
    plan = model._plan
    feats = plan.read_activation(0, (n, 128, 1 + t // 1024)).cpu()
    assert rel_err(np.exp(feats.numpy()), np.exp(taps["input"][:, 0].numpy())) < 1e-5
    for k in range(5):
        ref_k = taps["block%d" % k]
        got_k = plan.read_activation(1 + k, tuple(ref_k.shape)).cpu()
        assert rel_err(got_k.numpy(), ref_k.numpy()) < (1e-4 if precision == "fp32" else 1e-3), k
    assert rel_err(got_eval.numpy(), ref_eval.numpy()) < (1e-4 if precision == "fp32" else 1e-3)
    assert torch.equal(got_eval.argmax(1), ref_eval.argmax(1))

    params = {k: (v.clone().requires_grad_() if v.dtype.is_floating_point and "running" not in k else v.clone())
              for k, v in sd.items()}
    ref_train = restate.net2d_forward(params, config, signal, training=True)
    restate.lsep_loss(ref_train, torch.from_numpy(labels_np), average=False).mean().backward()
    from networks.losses import lsep_loss
    model.train()
    got_train = model(signal.cuda())["class_logits"]
    lsep_loss(got_train, torch.from_numpy(labels_np).cuda(), average=False).mean().backward()
    assert rel_err(got_train.detach().cpu().numpy(), ref_train.detach().numpy()) < 1e-3
    assert torch.equal(got_train.argmax(1).cpu(), ref_train.argmax(1))
    # Gradients.  With 1.5 s clips the last blocks see 8 x 4 x 2 values per channel, so ONE PReLU / max-pool
    # decision that lands on the other side of its kink (|bn(z)| below the ~1e-5 float32 forward difference)
    # moves a whole channel's gradient by ~1/64 and everything upstream by a fraction of a percent.  That is
    # measured behaviour of two valid float32 evaluations, not an arithmetic defect (tools/grad_debug2.py: the
    # difference sits in a single channel of one d(beta) while d(gamma) and the slope gradient agree to 3e-5),
    # so the gate is flip-robust: tight in the L2 norm per tensor, looser on the single worst element.
    # Conditioning.  tools/conditioning.py runs the ORACLE in float64 on this very problem and perturbs its input
    # features by 1e-5 relative: logits move by 4e-4 and every gradient tensor by 3-5e-2 in the L2 norm (train-mode
    # BatchNorm over 64 values per channel of an untrained net amplifies a forward difference ~4000x).  The float32
    # back end differs from the oracle by ~1e-6 in the forward pass and is held to 2e-2; the bf16x3 tensor-core back
    # end keeps ~2^-16 per product (forward logits ~1e-4 here), so its gradients are held to the oracle's own
    # sensitivity at that forward difference.  The GEMMs themselves are checked to 1e-4 (forward, dgrad, wgrad) in
    # test_gpu_kernels.py::test_conv_forward_backward.
    l2_gate = 2e-2 if precision == "fp32" else 8e-2
    gmax = max(float(p.grad.abs().max()) for p in params.values() if p.requires_grad)
    for k, p in model.named_parameters():
        ref = params[k].grad.numpy().astype(np.float64)
        got = p.grad.cpu().numpy().astype(np.float64)
        if ref.size >= 64:             # the norm is only flip-robust for tensors with many elements
            l2 = np.sqrt(((got - ref) ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-3 * gmax)
            assert l2 <= l2_gate, (k, l2)
        loose = 0.25 if ref.size < 8 else 0.1       # BN_in of block 0 has two elements: one flip is 1/2 of it
        if precision != "fp32":
            loose = 0.3                             # see the conditioning note above (single worst element)
        assert np.abs(got - ref).max() <= loose * np.abs(ref).max() + 2e-3 * gmax, k


def test_eval_mode_batch_independence_and_padding_at_full_size():
    """Size-independent properties at the bench workload (64 x 10 s, canonical width): eval-mode logits of a
    clip do not depend on its batch neighbours, and a zero-padded copy of a shorter clip matches running
    that padded clip alone."""
    model = build("TwoDimensionalCNNClassificationModel", dict(), None, "fp32")
    model.eval()
    wav = torch.from_numpy(restate.synth_waveforms(4, 441000, seed=5, kind="noise"))
    batch = wav.repeat(16, 1)[:, :, None].cuda()                 # 64 clips
    batch[7, 300000:] = 0.0                                       # a zero-padded shorter clip
    with torch.no_grad():
        full = model(batch)["class_logits"]
        alone = model(batch[4:8])["class_logits"]
    assert full.shape == (64, 80) and torch.isfinite(full).all()
    assert torch.allclose(full[4:8], alone, rtol=1e-5, atol=1e-6)
    assert torch.allclose(full[0], full[4 * 3], rtol=1e-5, atol=1e-6)     # identical clips, identical logits
    assert not torch.allclose(full[7], full[3], rtol=1e-3, atol=1e-3)


def test_training_loop_matches_oracle_steps(tmp_path):
    """Three optimiser steps through the public loop (`make_optimizer` + forward/LSEP/backward/step with
    the 1-cycle schedule) track the oracle's Adam-amsgrad trajectory."""
    from networks.losses import lsep_loss
    from ops.training import make_step
    cfg = dict(conv_base_depth=8, growth_rate=1.5)
    config = make_config(**cfg)
    model = build("TwoDimensionalCNNClassificationModel", cfg, None, "fp32", tmp=str(tmp_path))
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    n, t = 16, 40000
    wav = restate.synth_waveforms(n, t, seed=31)
    labels_np = restate.synth_labels(n, 80, seed=31)
    signal = torch.from_numpy(wav)[..., None]
    labels = torch.from_numpy(labels_np)

    # oracle: plain torch autograd + restated Adam
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    state = {k: [np.zeros(sd[k].numel(), np.float32) for _ in range(3)] for k in names}
    ref_losses = []
    for step in range(3):
        params = {k: (v.clone().requires_grad_() if k in names else v) for k, v in sd.items()}
        stats = {}
        out = restate.net2d_forward(params, config, signal, training=True, stats_out=stats)
        loss = restate.lsep_loss(out, labels, average=False).mean()
        loss.backward()
        ref_losses.append(loss.item())
        lr = restate.onecycle_lr(step, 0.0001, 0.005, 30)
        for k in names:
            p = sd[k].numpy().reshape(-1)
            restate.adam_amsgrad_step(p, params[k].grad.numpy().reshape(-1), *state[k], step + 1, lr)
        for prefix, (mean, var) in stats.items():
            sd[prefix + ".running_mean"] = 0.9 * sd[prefix + ".running_mean"] + 0.1 * mean
            sd[prefix + ".running_var"] = 0.9 * sd[prefix + ".running_var"] + 0.1 * var

    model.make_optimizer(max_steps=30)
    model.train()
    losses = []
    for step in range(3):
        make_step(model.scheduler, step=step + 1)
        out = model(signal.cuda())["class_logits"]
        loss = lsep_loss(out, labels.cuda(), average=False).mean()
        loss.backward()
        model.optimizer.step()
        model.optimizer.zero_grad()
        losses.append(loss.item())
    assert np.allclose(losses, ref_losses, rtol=2e-3)
    assert losses[2] < losses[0]
    got = model.state_dict()
    for k in names:
        if k.endswith("bias") and (".1.bias" in k or "conv" in k):
            continue      # biases feeding a batch-stat BN: zero gradient, Adam amplifies float noise
        assert np.abs(got[k].cpu().numpy() - sd[k].numpy()).max() < 2e-3 * max(1.0, np.abs(sd[k].numpy()).max()), k


def test_fit_validate_predict_roundtrip(tmp_path):
    """Public API: fit_validate -> checkpoint -> load_best_model -> predict, on synthetic loaders."""
    import torch.utils.data as data
    from ops.padding import make_collate_fn
    cfg = dict(conv_base_depth=8, growth_rate=1.5)
    model = build("TwoDimensionalCNNClassificationModel", cfg, None, "fp32", tmp=str(tmp_path))

    class DS(data.Dataset):
        def __init__(self, n, seed):
            rng = np.random.RandomState(seed)
            self.lengths = rng.randint(36000, 44000, size=n)
            self.wav = [restate.synth_waveforms(1, int(l), seed=seed + i, kind="noise")[0] for i, l in enumerate(self.lengths)]
            self.labels = restate.synth_labels(n, 80, seed=seed)
            self.transform = type("T", (), {"switch_off_augmentations": lambda self: None})()

        def __len__(self):
            return len(self.wav)

        def __getitem__(self, i):
            return dict(signal=self.wav[i][:, None], labels=self.labels[i], is_noisy=np.float32(0))

    collate = make_collate_fn({"signal": 0.0})
    train = data.DataLoader(DS(16, 1), batch_size=8, collate_fn=collate)
    valid = data.DataLoader(DS(8, 2), batch_size=4, collate_fn=collate)
    model.experiment.register_directory("checkpoints")
    scores = model.fit_validate(train, valid, epochs=2, fold=0, log_interval=1)
    assert len(scores) == 2 and all(0.0 <= s <= 1.0 for s in scores)
    assert os.path.isfile(os.path.join(model.experiment.checkpoints, "fold_0", "best_model.pth"))
    before = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_best_model(0)
    probs = model.predict(valid, n_tta=2)
    assert probs.shape == (8, 80) and probs.dtype == np.float32
    assert (probs >= 0).all() and (probs <= 1).all()
    assert set(before) == set(model.state_dict())


def test_full_size_determinism_and_backward_linearity():
    """Size-independent properties at the bench workload (64 x 10 s, canonical width, bf16x3 tensor-core back end, side
    stream overlap on): two training passes from the same state are bit-identical (every reduction has a fixed order),
    and the backward pass is exactly linear in the incoming gradient (scaling dlogits by 2 doubles every gradient
    bit for bit: all backward arithmetic is products and sums of the gradient, and 2x is exact in float32 / bf16)."""
    cfg = dict(output_dropout=0.0)
    model = build("TwoDimensionalCNNClassificationModel", cfg, None, "bf16x3")
    model.train()
    gen = torch.Generator(device="cuda").manual_seed(7)
    signal = 0.1 * torch.randn(64, 441000, generator=gen, device="cuda")
    dlogits = torch.randn(64, 80, generator=gen, device="cuda") / 64

    def run(scale):
        for p in model.parameters():
            p.grad = None
        out = model(signal[..., None])["class_logits"]
        out.backward(dlogits * scale)
        torch.cuda.synchronize()
        return out.detach().clone(), torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()

    sd = {k: v.clone() for k, v in model.state_dict().items()}
    l1, g1 = run(1.0)
    model.load_state_dict(sd)              # running statistics back to the same state
    l2, g2 = run(1.0)
    assert torch.isfinite(l1).all() and torch.isfinite(g1).all()
    assert torch.equal(l1, l2)
    assert torch.equal(g1, g2)
    model.load_state_dict(sd)
    l3, g3 = run(2.0)
    assert torch.equal(l1, l3)
    assert torch.equal(g3, 2.0 * g1)
    assert float(g1.abs().max()) > 0


def test_1d_model_full_size_properties():
    """BASELINE.json configs[2] at full size (1D CNN on raw STFT win 256 / hop 128, 64 x 10 s clips, canonical width):
    finite outputs and gradients, bit-identical repeat (fixed reduction orders), and eval-mode logits of a clip that do
    not depend on its batch neighbours."""
    cfg = dict(features="stft_256_128", output_dropout=0.0)
    model = build("HierarchicalCNNClassificationModel", cfg, None, "bf16x3")
    gen = torch.Generator(device="cuda").manual_seed(11)
    signal = 0.1 * torch.randn(64, 441000, generator=gen, device="cuda")
    dlogits = torch.randn(64, 80, generator=gen, device="cuda") / 64
    sd = {k: v.clone() for k, v in model.state_dict().items()}

    def run():
        model.load_state_dict(sd)
        model.train()
        for p in model.parameters():
            p.grad = None
        out = model(signal[..., None])["class_logits"]
        out.backward(dlogits)
        torch.cuda.synchronize()
        return out.detach().clone(), torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()

    l1, g1 = run()
    l2, g2 = run()
    assert l1.shape == (64, 80) and torch.isfinite(l1).all() and torch.isfinite(g1).all()
    assert torch.equal(l1, l2) and torch.equal(g1, g2)
    assert float(g1.abs().max()) > 0
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        full = model(signal[..., None])["class_logits"]
        part = model(signal[5:9, :, None])["class_logits"]
    assert rel_err(part.cpu().numpy(), full[5:9].cpu().numpy()) < 1e-5


def test_bucketed_inference_matches_oracle_on_identical_padded_batches():
    """BASELINE.json configs[4] semantics at small size: variable-length clips, length-bucketed batches, zero padding
    per batch; every clip's probabilities equal the oracle's on the SAME padded batch (padding changes the global-max
    features, so batches must be identical), and clips outside the buckets are reported as NaN rows."""
    from fsb200.inference import pack_batches, pad_batch, predict_bucketed
    cfg = dict(conv_base_depth=8, growth_rate=1.5)
    model = build("TwoDimensionalCNNClassificationModel", cfg, None, "bf16x3")
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    config = make_config(**cfg)
    rng = np.random.RandomState(5)
    lengths = rng.randint(33 * 1024, 90000, size=11).tolist() + [20000]          # the last clip is below buckets[0]
    clips = [restate.synth_waveforms(1, n, seed=100 + k)[0] for k, n in enumerate(lengths)]
    buckets = [33 * 1024, 50000, 70000, 100000]
    got, stats = predict_bucketed(model, clips, buckets, max_batch_elems=150000, return_stats=True)
    assert got.shape == (12, 80) and np.isnan(got[11]).all() and stats["dropped"] == 1
    assert stats["padding_overhead"] >= 0.0 and stats["real_samples"] == sum(lengths[:11])
    batches, _ = pack_batches(lengths, buckets, 150000)
    for indices in batches:
        batch = torch.from_numpy(pad_batch(clips, indices))
        with torch.no_grad():
            ref = torch.sigmoid(restate.net2d_forward(sd, config, batch, training=False)).numpy()
        assert rel_err(got[indices], ref) < 1e-3
        assert np.array_equal(np.argsort(-got[indices], 1)[:, :3], np.argsort(-ref, 1)[:, :3])
