"""How ill-conditioned is the canonical-width parity problem (default 8 untrained clips x 1.5 s; pass n and t for other
sizes, e.g. `python tools/conditioning.py 1e-5 64 441000` for the bench workload -- needs ~35 GB of host memory)?  Runs the CPU ORACLE in
float64, in float32, and in float64 with the input features perturbed by `noise` relative, and prints how far
logits and gradients move.  Measured in the build container (seed 21): float32 vs float64 logits 3e-6; 1e-5 feature
noise -> logits 4e-4, gradient tensors 3-5e-2 (L2).  This sets the gradient gates of
tests/test_gpu_network.py::test_canonical_width_against_oracle and of tests/test_gpu_fullsize.py.
usage: python tools/conditioning.py [noise [n t [seeds...]]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import restate  # noqa: E402
from oracle.reference_shim import make_config  # noqa: E402

noise = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5
config = make_config(output_dropout=0.0)
n, t = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 66150)
seeds = [int(a) for a in sys.argv[4:]] or [21, 22]
sd = restate.init_state_dict(config, two_d=True, seed=42)
print("n = %d clips x t = %d samples, feature noise %.0e" % (n, t, noise))
for seed in seeds:
    wav = restate.synth_waveforms(n, t, seed=seed)
    labels = torch.from_numpy(restate.synth_labels(n, 80, seed=seed))
    feats = restate.features(torch.from_numpy(wav)[..., None], config["data"]["features"])

    def run(dtype, eps=0.0):
        params = {k: (v.clone().to(dtype).requires_grad_() if v.dtype.is_floating_point and "running" not in k
                      else (v.clone().to(dtype) if v.dtype.is_floating_point else v.clone())) for k, v in sd.items()}
        f = feats.to(dtype)
        if eps:
            g = torch.Generator().manual_seed(0)
            f = f * (1 + eps * torch.randn(f.shape, generator=g).to(dtype))
        out = restate.net2d_forward(params, config, None, training=True, feats_in=f)
        restate.lsep_loss(out, labels.to(dtype), average=False).mean().backward()
        return out.detach().double(), {k: p.grad.double() for k, p in params.items() if getattr(p, "grad", None) is not None}

    o64, g64 = run(torch.float64)
    o32, g32 = run(torch.float32)
    on, gn = run(torch.float64, noise)

    def l2(a, b):
        return float((a - b).norm() / b.norm())

    print("seed %d  logits: float32 vs float64 %.2e ; feature noise %.0e vs clean %.2e" % (
        seed, float((o32 - o64).abs().max() / o64.abs().max()), noise, float((on - o64).abs().max() / o64.abs().max())))
    for k in ["conv_modules.0.1.weight", "conv_modules.0.3.bias", "conv_modules.2.5.bn3.bias",
              "conv_modules.4.5.conv1.weight", "output_transform.1.weight"]:
        print("   %-34s float32 %.2e   noise %.2e" % (k, l2(g32[k], g64[k]), l2(gn[k], g64[k])))
    big = [k for k in g64 if g64[k].numel() >= 64 and float(g64[k].norm()) > 0]
    print("   worst tensor (>= 64 elements): float32 %.2e   noise %.2e" % (
        max(l2(g32[k], g64[k]) for k in big), max(l2(gn[k], g64[k]) for k in big)))
