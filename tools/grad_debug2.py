"""Debug: internal gradients of the LAST block vs autograd of the oracle (float32 back end)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "freesound-classification_b200"))
import numpy as np, torch, torch.nn.functional as F
os.environ["FSB200_PRECISION"] = "fp32"
from oracle import restate
from oracle.reference_shim import FakeExperiment, make_config
from networks.classifiers import TwoDimensionalCNNClassificationModel
from networks.losses import lsep_loss
n, t = 8, 66150
config = make_config()
torch.manual_seed(42)
model = TwoDimensionalCNNClassificationModel(FakeExperiment(config), device="cuda:0")
sd = {k: (v.detach().cpu().clone().requires_grad_() if v.dtype.is_floating_point and "running" not in k else v.detach().cpu().clone()) for k, v in model.state_dict().items()}
wav = restate.synth_waveforms(n, t, seed=21, kind="noise"); labels_np = restate.synth_labels(n, 80, seed=21)
signal = torch.from_numpy(wav)[..., None]
labels = torch.from_numpy(labels_np)
# oracle with intermediate taps on the last block
K = 4
feats = restate.features(signal, config["data"]["features"])
h = restate.add_frequency_encoding(feats.unsqueeze(1))
heads = []
inter = {}
for k in range(5):
    p = "conv_modules.%d" % k
    h = restate._bn(h, sd, p + ".0", True)
    h = F.conv2d(h, sd[p + ".1.weight"], sd[p + ".1.bias"], padding=1)
    h = F.max_pool2d(h, 2, 2)
    if k == K:
        h = h.detach().requires_grad_(); inter["zp"] = h
    hb = restate._bn(h, sd, p + ".3", True)
    r0 = F.prelu(hb, sd[p + ".4.weight"])
    if k == K:
        r0.retain_grad(); inter["r0"] = r0
    q = p + ".5"
    z1 = F.conv2d(r0, sd[q + ".conv1.weight"], sd[q + ".conv1.bias"])
    if k == K:
        z1.retain_grad(); inter["z1"] = z1
    o = restate._bn(z1, sd, q + ".bn1", True)
    a1 = F.prelu(o, sd[q + ".prelu1.weight"])
    if k == K:
        a1.retain_grad(); inter["a1"] = a1
    o = F.conv2d(a1, sd[q + ".conv2.weight"], sd[q + ".conv2.bias"], padding=1)
    o = restate._bn(o, sd, q + ".bn2", True)
    o = F.prelu(o, sd[q + ".prelu2.weight"])
    o = F.conv2d(o, sd[q + ".conv3.weight"], sd[q + ".conv3.bias"])
    o = restate._bn(o, sd, q + ".bn3", True)
    o = o + r0
    h = F.prelu(o, sd[q + ".prelu3.weight"])
    if k >= 1:
        heads.append(F.adaptive_max_pool2d(h, 1).squeeze(-1).squeeze(-1))
logits = restate._head(torch.cat(heads, -1), sd, True, 0.0, None)
restate.lsep_loss(logits, labels, average=False).mean().backward()
model.train()
got = model(signal.cuda())["class_logits"]
lsep_loss(got, labels.cuda(), average=False).mean().backward()
plan = model._plan
shape = tuple(inter["zp"].shape)
def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())
print("logits", rel(got.detach().cpu(), logits.detach()))
for j, name, ref in [(0, "zp", inter["zp"].detach()), (1, "r0", inter["r0"].detach()), (2, "z1", inter["z1"].detach()),
                     (3, "dz1", inter["z1"].grad), (6, "da1", inter["a1"].grad), (7, "dzp", inter["zp"].grad)]:
    ours = plan.read_activation(300 + 10 * K + j, shape).cpu()
    print("%-5s rel err %.3e   max|ref| %.3e" % (name, rel(ours, ref), float(ref.abs().max())))
dr0 = plan.read_activation(300 + 10 * K + 4, shape).cpu() + plan.read_activation(300 + 10 * K + 5, shape).cpu()
print("dr0 (a+b) rel err %.3e" % rel(dr0, inter["r0"].grad))
ref_sd = {k: v for k, v in sd.items()}
named = dict(model.named_parameters())
for name in ["conv_modules.4.5.bn1.bias", "conv_modules.4.5.bn2.bias", "conv_modules.4.5.bn1.weight"]:
    g = named[name].grad.cpu().numpy()
    r = sd[name].grad.numpy() if sd[name].grad is not None else None
    if r is None:
        print(name, "no oracle grad (sd tensors are not leaves)"); continue
    d = np.abs(g - r)
    idx = np.argsort(-d)[:12]
    print(name, "worst idx", idx.tolist(), "err", d[idx].round(6).tolist(), "ref", r[idx].round(5).tolist())
    print("   err>1e-5 count", int((d > 1e-5).sum()), "of", d.size, " first few err", d[:8].round(7).tolist())
