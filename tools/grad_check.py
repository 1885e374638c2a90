"""Per-parameter gradient error of the CUDA path against the CPU oracle at canonical width (8 x 1.5 s clips):
L2 error per tensor and how many elements carry it (a PReLU / max-pool decision flipping in one channel shows up
as ONE outlier element of a d(beta), not as a spread).  usage: python tools/grad_check.py [precision] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "freesound-classification_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

os.environ["FSB200_PRECISION"] = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 21
from networks.classifiers import TwoDimensionalCNNClassificationModel  # noqa: E402
from networks.losses import lsep_loss  # noqa: E402
from oracle import restate  # noqa: E402
from oracle.reference_shim import FakeExperiment, make_config  # noqa: E402

n, t = 8, 66150
config = make_config()
torch.manual_seed(42)
model = TwoDimensionalCNNClassificationModel(FakeExperiment(config), device="cuda:0")
sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
wav = restate.synth_waveforms(n, t, seed=seed)
labels_np = restate.synth_labels(n, 80, seed=seed)
signal = torch.from_numpy(wav)[..., None]
params = {k: (v.clone().requires_grad_() if v.dtype.is_floating_point and "running" not in k else v.clone())
          for k, v in sd.items()}
ref = restate.net2d_forward(params, config, signal, training=True)
restate.lsep_loss(ref, torch.from_numpy(labels_np), average=False).mean().backward()
model.train()
got = model(signal.cuda())["class_logits"]
lsep_loss(got, torch.from_numpy(labels_np).cuda(), average=False).mean().backward()
print("logits rel err %.3e" % float((got.detach().cpu() - ref.detach()).abs().max() / ref.detach().abs().max()))
gmax = max(float(p.grad.abs().max()) for p in params.values() if p.requires_grad)
worst = []
for k, p in model.named_parameters():
    r = params[k].grad.numpy().astype(np.float64)
    g = p.grad.cpu().numpy().astype(np.float64)
    d = np.abs(g - r)
    l2 = np.sqrt((d ** 2).sum()) / max(np.sqrt((r ** 2).sum()), 1e-3 * gmax)
    big = int((d > 0.02 * np.abs(r).max() + 1e-4 * gmax).sum())
    worst.append((l2, k, r.size, big, float(d.max() / max(np.abs(r).max(), 1e-30))))
for l2, k, size, big, mx in sorted(worst, reverse=True)[:12]:
    print("%-40s l2 %.3e  max-elem %.3e  outliers %d / %d" % (k, l2, mx, big, size))
