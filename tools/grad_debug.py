"""Debug: per-parameter gradient error of the canonical-width test (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "freesound-classification_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
os.environ["FSB200_PRECISION"] = sys.argv[1] if len(sys.argv) > 1 else "fp32"
from oracle import restate
from oracle.reference_shim import FakeExperiment, make_config
from networks.classifiers import TwoDimensionalCNNClassificationModel
from networks.losses import lsep_loss
n, t = int(os.environ.get("N", 8)), int(os.environ.get("T", 66150))
config = make_config()
torch.manual_seed(42)
model = TwoDimensionalCNNClassificationModel(FakeExperiment(config), device="cuda:0")
sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
wav = restate.synth_waveforms(n, t, seed=21, kind=os.environ.get("KIND", "structured")); labels_np = restate.synth_labels(n, 80, seed=21)
signal = torch.from_numpy(wav)[..., None]
for dt in (torch.float32, torch.float64):
    params = {k: ((v.to(dt) if v.dtype.is_floating_point else v).clone().requires_grad_() if v.dtype.is_floating_point and "running" not in k else (v.to(dt) if v.dtype.is_floating_point else v).clone()) for k, v in sd.items()}
    try:
        feats = restate.features(signal, config["data"]["features"]).to(dt)
        ref = restate.net2d_forward(params, config, None, training=True, feats_in=feats)
        restate.lsep_loss(ref, torch.from_numpy(labels_np).to(dt), average=False).mean().backward()
    except Exception as e:
        print("oracle in", dt, "failed:", e); continue
    if dt == torch.float32:
        p32 = params; ref32 = ref
    else:
        p64 = params; ref64 = ref
model.train()
got = model(signal.cuda())["class_logits"]
lsep_loss(got, torch.from_numpy(labels_np).cuda(), average=False).mean().backward()
print("logits: ours-vs-f32 %.2e" % float((got.detach().cpu() - ref32.detach()).abs().max() / ref32.detach().abs().max()))
have64 = "p64" in globals()
if have64:
    print("logits: ours-vs-f64 %.2e   f32-vs-f64 %.2e" % (float((got.detach().cpu().double() - ref64.detach()).abs().max() / ref64.detach().abs().max()), float((ref32.detach().double() - ref64.detach()).abs().max() / ref64.detach().abs().max())))
gmax = max(float(p.grad.abs().max()) for p in p32.values() if p.requires_grad)
for k, p in model.named_parameters():
    r32 = p32[k].grad.numpy(); g = p.grad.cpu().numpy()
    tol = 1e-2 * np.abs(r32).max() + 1e-3 * gmax
    line = "%-40s max|ref| %.2e rel %.2e ours-f32 %.3f" % (k, np.abs(r32).max(), np.abs(g - r32).max() / max(np.abs(r32).max(), 1e-30), np.abs(g - r32).max() / tol)
    if have64:
        r64 = p64[k].grad.numpy()
        line += "  ours-f64 %.3f  f32-f64 %.3f" % (np.abs(g - r64).max() / tol, np.abs(r32 - r64).max() / tol)
    if np.abs(g - r32).max() / tol > float(os.environ.get("SHOW", 0.5)):
        print(line)
