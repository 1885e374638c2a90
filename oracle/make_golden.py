"""TEST INFRASTRUCTURE ONLY -- generates `tests/golden/*.npz` by executing the reference's OWN
modules (through `oracle/reference_shim.py`) on seeded synthetic inputs.  Runs only where
`/root/reference` is mounted (the build container); the vectors it writes are committed so
the GPU box -- which has no reference checkout -- can check both the oracle restatement and
the CUDA path against them.

    python -m oracle.make_golden            # from the repo root
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import restate  # noqa: E402
from oracle.reference_shim import FakeExperiment, ReferenceModules, make_config  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def randomize_state(model, seed):
    """Make every parameter / buffer non-default (deterministically) so BN affine terms,
    running statistics and PReLU slopes are all exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.state_dict().items():
            if name.endswith("num_batches_tracked"):
                continue
            if name.endswith("running_var"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif name.endswith("running_mean"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            elif ".bn" in name or name.split(".")[-2] in ("0", "3", "2") and p.dim() == 1 and "weight" in name:
                # BN gamma / PReLU slope (1-D weights)
                p.copy_(p + 0.2 * torch.rand(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(p + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(p * (1.0 + 0.1 * torch.randn(p.shape, generator=g)))


def model_case(ref, cls_name, config, n, t, seed, out_name, wave_kind="structured"):
    torch.manual_seed(42)
    exp = FakeExperiment(config)
    model = getattr(ref.classifiers, cls_name)(exp, device="cpu")
    init_sd = {k: v.clone() for k, v in model.state_dict().items()}
    randomize_state(model, seed)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}

    wav = restate.synth_waveforms(n, t, seed=seed, kind=wave_kind)
    labels = restate.synth_labels(n, config["data"]["_n_classes"], seed=seed)
    signal = torch.from_numpy(wav).unsqueeze(-1)

    model.eval()
    with torch.no_grad():
        logits_eval = model(signal)["class_logits"].clone()

    model.train()
    model.zero_grad()
    logits_train = model(signal)["class_logits"]
    per_sample = ref.losses.lsep_loss(logits_train, torch.from_numpy(labels), average=False)
    loss = per_sample.mean()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    sd1 = model.state_dict()

    out = dict(
        n=n, t=t, seed=seed, wave_kind=wave_kind,
        logits_eval=logits_eval.numpy(), logits_train=logits_train.detach().numpy(),
        per_sample_loss=per_sample.detach().numpy(), loss=loss.item(),
        lwlrap_train=ref.utils.lwlrap(labels, torch.sigmoid(logits_train).detach().numpy()),
        init_checksum=np.array([float(v.double().sum()) for k, v in sorted(init_sd.items())]),
    )
    for k, v in sd0.items():
        out["sd/" + k] = v.numpy()
    for k, v in grads.items():
        out["grad/" + k] = v.numpy()
    for k, v in sd1.items():
        if "running" in k or "num_batches" in k:
            out["sd_after/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, out_name), **out)
    print(out_name, "loss", loss.item(), "logits", logits_train.abs().max().item())


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ReferenceModules()
    print("reference root:", ref.root)

    # ---- 1. feature path: compute_torch_stft + mel conv1d + log (ops/utils.py:110-127,
    #         networks/classifiers.py:565-579)
    wav = restate.synth_waveforms(2, 40000, seed=7)
    sig = torch.from_numpy(wav)
    mag = ref.utils.compute_torch_stft(sig, "mel_2048_1024_128")
    fb = ref.utils.make_mel_filterbanks("mel_2048_1024_128")
    mel = torch.log(torch.nn.functional.conv1d(mag, torch.from_numpy(fb).unsqueeze(-1)) + 1e-4)
    mag256 = ref.utils.compute_torch_stft(sig, "stft_256_128")
    # independent pin of the restated librosa filterbank: torchaudio's slaney/slaney filters
    import torchaudio
    ta = torchaudio.functional.melscale_fbanks(1025, 5.0, 22050.0, 128, 44100,
                                               norm="slaney", mel_scale="slaney").T.numpy()
    nz = np.nonzero(ta)
    np.savez_compressed(
        os.path.join(OUT, "features.npz"),
        n=2, t=40000, seed=7,
        stft_mag_2048=mag.numpy(), logmel=mel.numpy(),
        logstft_256=torch.log(mag256 + 1e-4).numpy(),
        fb_checksum=np.array([fb.astype(np.float64).sum(), np.count_nonzero(fb)]),
        ta_rows=nz[0].astype(np.int16), ta_cols=nz[1].astype(np.int16), ta_vals=ta[nz],
        scipy_stft=ref.audio.compute_stft(wav[0][:8000], 256, 128, log=True),
    )

    # ---- 2. models (small widths so the vectors stay small; same topology as canonical)
    cfg2d = make_config(conv_base_depth=8, growth_rate=1.5)
    model_case(ref, "TwoDimensionalCNNClassificationModel", cfg2d, 4, 40000, 11, "net2d_small.npz")
    cfg2d_b = make_config(conv_base_depth=8, growth_rate=2.0, start_deep_supervision_on=2,
                          num_conv_blocks=5)
    model_case(ref, "TwoDimensionalCNNClassificationModel", cfg2d_b, 3, 36000, 12,
               "net2d_pow2.npz", wave_kind="noise")
    cfg1d = make_config(features="stft_256_128", conv_base_depth=8, growth_rate=1.5)
    model_case(ref, "HierarchicalCNNClassificationModel", cfg1d, 4, 12000, 13, "net1d_small.npz")
    # aggregation_type="rnn" heads: LayerNorm + bidirectional GRU(128) over time (networks/classifiers.py:514-522)
    cfg2d_rnn = make_config(conv_base_depth=8, growth_rate=1.5, aggregation_type="rnn", start_deep_supervision_on=3)
    model_case(ref, "TwoDimensionalCNNClassificationModel", cfg2d_rnn, 4, 40000, 14, "net2d_rnn_small.npz")

    # ---- 3. LSEP (networks/losses.py:47-58)
    g = torch.Generator().manual_seed(3)
    s = (2.0 * torch.randn(6, 80, generator=g)).requires_grad_()
    tgt = torch.from_numpy(restate.synth_labels(6, 80, seed=3))
    tgt[5] = 0.0                                     # a row without positives
    per = ref.losses.lsep_loss(s, tgt, average=False)
    per.mean().backward()
    np.savez_compressed(os.path.join(OUT, "lsep.npz"), scores=s.detach().numpy(),
                        targets=tgt.numpy(), per_sample=per.detach().numpy(),
                        grad_mean=s.grad.numpy(),
                        mean=ref.losses.lsep_loss(s, tgt).item())

    # ---- 4. Adam-amsgrad + OneCycle (ops/training.py:9-12,208-234)
    g = torch.Generator().manual_seed(4)
    p = torch.nn.Parameter(torch.randn(257, generator=g))
    opt = ref.training.OPTIMIZERS["adam"]([p], 0.001, weight_decay=0.01)
    sched = ref.training.make_scheduler("1cycle_0.0001_0.005", max_steps=10)(opt)
    ps, gs, lrs = [p.detach().clone().numpy()], [], []
    for step in range(10):
        ref.training.make_step(sched, step=step + 1)
        lrs.append(opt.param_groups[0]["lr"])
        grad = torch.randn(257, generator=g) * (0.1 if step % 3 else 3.0)
        p.grad = grad.clone()
        opt.step()
        gs.append(grad.numpy())
        ps.append(p.detach().clone().numpy())
    np.savez_compressed(os.path.join(OUT, "adam.npz"), params=np.stack(ps), grads=np.stack(gs),
                        lrs=np.array(lrs), weight_decay=0.01)

    # ---- 5. lwlrap (ops/utils.py:17-26)
    rng = np.random.RandomState(5)
    truth = restate.synth_labels(40, 80, seed=5)
    truth[3] = 0
    scores = rng.rand(40, 80).astype(np.float32)
    scores[7, :10] = 0.5                              # ties
    np.savez_compressed(os.path.join(OUT, "lwlrap.npz"), truth=truth, scores=scores,
                        value=ref.utils.lwlrap(truth, scores))

    # ---- 6. collate + bucketing (ops/padding.py)
    rng = np.random.RandomState(6)
    lens = [5, 9, 3, 9]
    batch = [dict(signal=rng.randn(l, 1).astype(np.float32), labels=np.float32([i])) for i, l in enumerate(lens)]
    raw = [b["signal"].copy() for b in batch]
    coll = ref.padding.make_collate_fn({"signal": 0.0})(batch)

    class DS:
        lengths = rng.randint(1, 31, size=200) * 44100

    random.seed(6)
    bs = ref.padding.BucketingSampler(DS(), max_batch_elems=64 * 441000 // 8,
                                      buckets=[0, 5 * 44100, 10 * 44100, 20 * 44100, 31 * 44100])
    flat = np.array([i for b in bs.batches for i in b] , dtype=np.int64)
    sizes = np.array([len(b) for b in bs.batches], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "padding.npz"),
                        raw=np.concatenate([r.ravel() for r in raw]), lens=np.array(lens),
                        collated=coll["signal"].numpy(), ds_lengths=DS.lengths,
                        batch_flat=flat, batch_sizes=sizes)

    # ---- 7. MixUp (ops/audio.py:32-52)
    rng = np.random.RandomState(8)
    a1, a2 = rng.randn(100).astype(np.float32), rng.randn(100).astype(np.float32)
    l1, l2 = restate.synth_labels(2, 80, seed=8)
    mixed, ml = ref.audio.mix_audio_and_labels(a1.copy(), a2.copy(), l1, l2)
    a3 = rng.randn(60).astype(np.float32)
    np.random.seed(9)
    random.seed(9)
    mixed_u, ml_u = ref.audio.mix_audio_and_labels(a1.copy(), a3.copy(), l1, l2)
    np.random.seed(9)
    random.seed(9)
    a_val = np.random.uniform(0.4, 0.6)
    start_val = random.randint(0, 100 - 1 - 60)
    np.savez_compressed(os.path.join(OUT, "mixup.npz"), a1=a1, a2=a2, a3=a3, l1=l1, l2=l2,
                        mixed=mixed, labels=ml, mixed_unequal=mixed_u, labels_unequal=ml_u,
                        a=a_val, start=start_val)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
