"""GPU parity tests of the individual C-ABI entry points (through ctypes) against the CPU oracle
(`oracle/restate.py`) and the golden vectors generated from the reference."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import restate

pytestmark = pytest.mark.gpu


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ------------------------------------------------------------------------------------------ features
def test_stft_magnitude_matches_reference_golden():
    from ops.utils import compute_torch_stft
    g = load("features.npz")
    wav = restate.synth_waveforms(int(g["n"]), int(g["t"]), seed=int(g["seed"]))
    mag = compute_torch_stft(torch.from_numpy(wav).cuda(), "mel_2048_1024_128").cpu().numpy()
    assert mag.shape == g["stft_mag_2048"].shape
    # tolerance: 1e-3 relative is the north-star gate; fp32 FFT reaches ~1e-6 of the frame maximum
    assert rel_err(mag, g["stft_mag_2048"]) < 5e-6


def test_logmel_and_logstft_match_reference_golden():
    from ops.utils import compute_log_features
    g = load("features.npz")
    wav = torch.from_numpy(restate.synth_waveforms(int(g["n"]), int(g["t"]), seed=int(g["seed"]))).cuda()
    logmel = compute_log_features(wav, "mel_2048_1024_128").cpu().numpy()
    assert logmel.shape == g["logmel"].shape
    assert np.abs(logmel - g["logmel"]).max() < 1e-3 * np.abs(g["logmel"]).max()
    assert rel_err(np.exp(logmel), np.exp(g["logmel"])) < 1e-5
    logstft = compute_log_features(wav, "stft_256_128").cpu().numpy()
    assert logstft.shape == g["logstft_256"].shape
    assert rel_err(np.exp(logstft), np.exp(g["logstft_256"])) < 1e-5


@pytest.mark.parametrize("descriptor,n,t", [
    ("mel_2048_1024_128", 3, 1025), ("mel_2048_1024_128", 2, 44100 * 4), ("mel_1024_512_64", 2, 30000),
    ("stft_256_128", 3, 12001), ("stft_512_256", 2, 9000), ("stft_128_64", 2, 5000), ("mel_2048_1024_128", 64, 1024 * 33 + 5)])
def test_features_match_oracle(descriptor, n, t):
    from ops.utils import compute_log_features, compute_torch_stft
    wav = restate.synth_waveforms(n, t, seed=n + t, kind="structured")
    if n == 3:
        wav[1, t // 2:] = 0.0          # a zero-padded (shorter) clip inside the batch
    ref = restate.features(torch.from_numpy(wav)[..., None], descriptor).numpy()
    got = compute_log_features(torch.from_numpy(wav).cuda(), descriptor).cpu().numpy()
    assert got.shape == ref.shape
    assert rel_err(np.exp(got), np.exp(ref)) < 1e-5
    assert np.abs(got - ref).max() < 1e-3 * np.abs(ref).max()
    name, n_fft, hop, _ = restate.parse_descriptor(descriptor)
    mag = compute_torch_stft(torch.from_numpy(wav).cuda(), descriptor).cpu().numpy()
    assert rel_err(mag, restate.stft_magnitude(torch.from_numpy(wav), n_fft, hop).numpy()) < 5e-6


def test_feature_kernel_rejects_short_clips_and_cpu_tensors():
    from ops.utils import compute_torch_stft
    with pytest.raises(RuntimeError):
        compute_torch_stft(torch.zeros(1, 1024).cuda(), "mel_2048_1024_128")   # T must exceed n_fft/2
    with pytest.raises(RuntimeError):
        compute_torch_stft(torch.zeros(1, 4096), "mel_2048_1024_128")          # no CPU path


def test_scipy_style_compute_stft():
    from ops.audio import compute_stft
    g = load("features.npz")
    wav = restate.synth_waveforms(2, 40000, seed=7)
    s = compute_stft(wav[0][:8000], 256, 128, log=True)
    assert s.shape == g["scipy_stft"].shape
    assert rel_err(np.exp(s), np.exp(g["scipy_stft"])) < 1e-5


# ------------------------------------------------------------------------------------------ LSEP
def test_lsep_matches_reference_golden():
    from networks.losses import lsep_loss
    g = load("lsep.npz")
    s = torch.from_numpy(g["scores"]).cuda().requires_grad_()
    t = torch.from_numpy(g["targets"]).cuda()
    per = lsep_loss(s, t, average=False)
    assert rel_err(per.detach().cpu().numpy(), g["per_sample"]) < 1e-5
    per.mean().backward()
    assert rel_err(s.grad.cpu().numpy(), g["grad_mean"]) < 1e-5
    assert abs(lsep_loss(s, t).item() - float(g["mean"])) < 1e-5
    assert per[5].item() == 0.0                       # row without positives


def test_lsep_general_targets_and_overflow_like_reference():
    from networks.losses import lsep_loss
    gen = torch.Generator().manual_seed(0)
    s = torch.randn(33, 80, generator=gen) * 3
    t = torch.randint(0, 3, (33, 80), generator=gen).float()       # non-binary targets: pairwise form
    ref = restate.lsep_loss(s, t, average=False)
    got = lsep_loss(s.cuda(), t.cuda(), average=False).cpu()
    assert rel_err(got.numpy(), ref.numpy()) < 1e-5
    big = torch.zeros(1, 80)
    big[0, 0], big[0, 1] = -100.0, 100.0
    tb = torch.zeros(1, 80)
    tb[0, 0] = 1.0
    # no max-shift, like the reference: exp(200) overflows.  The reference's masked product turns the
    # overflow into inf * 0 = nan for the masked pairs; the factorised kernel reports +inf (never finite).
    assert torch.isinf(lsep_loss(big.cuda(), tb.cuda(), average=False)).all()
    assert not torch.isfinite(restate.lsep_loss(big, tb, average=False)).any()


# ------------------------------------------------------------------------------------------ Adam
def test_fused_adam_matches_reference_golden():
    from ops.training import OPTIMIZERS, make_scheduler, make_step
    g = load("adam.npz")
    p = torch.nn.Parameter(torch.from_numpy(g["params"][0].copy()).cuda())
    opt = OPTIMIZERS["adam"]([p], 0.001, weight_decay=float(g["weight_decay"]))
    sched = make_scheduler("1cycle_0.0001_0.005", max_steps=10)(opt)
    for step in range(10):
        make_step(sched, step=step + 1)
        p.grad = torch.from_numpy(g["grads"][step]).cuda()
        opt.step()
        assert np.abs(p.detach().cpu().numpy() - g["params"][step + 1]).max() < 2e-6


def test_fused_adam_multi_tensor_matches_torch():
    from ops.training import FusedAdam
    gen = torch.Generator().manual_seed(1)
    shapes = [(70000,), (3, 5), (1,), (129, 33)]
    ps = [torch.randn(s, generator=gen) for s in shapes]
    ours = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    theirs = [torch.nn.Parameter(p.clone()) for p in ps]
    o1 = FusedAdam(ours, lr=0.01, weight_decay=0.0)
    o2 = torch.optim.Adam(theirs, lr=0.01, amsgrad=True)
    for _ in range(5):
        for a, b in zip(ours, theirs):
            gr = torch.randn(a.shape, generator=gen)
            a.grad, b.grad = gr.cuda(), gr.clone()
        o1.step()
        o2.step()
    for a, b in zip(ours, theirs):
        assert (a.detach().cpu() - b.detach()).abs().max() < 2e-6


# ------------------------------------------------------------------------------------------ conv GEMMs
CONV_CASES = [(2, 5, 7, 6, 9, 3, 3), (3, 19, 33, 5, 11, 1, 1), (2, 37, 21, 1, 50, 1, 3), (1, 100, 150, 8, 13, 3, 3),
              (2, 16, 16, 4, 130, 3, 3),
              # wide tiles (N = 240 / 256 / 2 x 176): one activation box per tap, single accumulator set
              (1, 225, 250, 6, 9, 3, 3), (2, 337, 240, 5, 7, 3, 3), (1, 506, 337, 4, 13, 1, 1)]


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("n,cin,cout,h,w,kh,kw", CONV_CASES)
def test_conv_forward_backward(precision, n, cin, cout, h, w, kh, kw):
    from fsb200.runtime import conv_backward, conv_forward
    gen = torch.Generator().manual_seed(n * 1000 + cin)
    x = torch.randn(n, cin, h, w, generator=gen)
    wt = torch.randn(cout, cin, kh, kw, generator=gen) / (cin * kh * kw) ** 0.5
    b = torch.randn(cout, generator=gen)
    dy = torch.randn(n, cout, h, w, generator=gen)
    xr, wr = x.clone().requires_grad_(), wt.clone().requires_grad_()
    ref = F.conv2d(xr, wr, b, padding=(kh // 2, kw // 2))
    ref.backward(dy)
    tol = 2e-5 if precision == "fp32" else 1e-4          # bf16x3 keeps ~2^-16 per product
    y = conv_forward(x.cuda(), wt.cuda(), b.cuda(), precision).cpu()
    assert rel_err(y.numpy(), ref.detach().numpy()) < tol
    dx, dw, db = conv_backward(x.cuda(), wt.cuda(), dy.cuda(), precision)
    assert rel_err(dx.cpu().numpy(), xr.grad.numpy()) < tol
    assert rel_err(dw.cpu().numpy(), wr.grad.numpy()) < tol
    assert rel_err(db.cpu().numpy(), dy.sum((0, 2, 3)).numpy()) < 1e-5


# ------------------------------------------------------------------------------------------ MixUp
def test_device_mixup_equal_length_branch():
    from fsb200.runtime import mixup_equal
    g = load("mixup.npz")
    pcm = torch.from_numpy(np.stack([g["a1"], g["a2"], g["a1"]])).cuda()
    labels = torch.from_numpy(np.stack([g["l1"], g["l2"], g["l1"]])).cuda()
    out, lab = mixup_equal(pcm, labels, torch.tensor([1, -1, 1]))
    assert np.array_equal(out[0].cpu().numpy(), g["mixed"])           # (a + b) / 2, bit exact
    assert np.array_equal(lab[0].cpu().numpy(), g["labels"])
    assert np.array_equal(out[1].cpu().numpy(), g["a2"]) and np.array_equal(lab[1].cpu().numpy(), g["l2"])
