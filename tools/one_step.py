"""Run a few canonical training steps (for ncu / compute-sanitizer): python tools/one_step.py [steps] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "freesound-classification_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

os.environ.setdefault("FSB200_PRECISION", "mixed")
from networks.classifiers import TwoDimensionalCNNClassificationModel  # noqa: E402
from networks.losses import lsep_loss  # noqa: E402
from ops.training import make_step  # noqa: E402
from fsb200.experiment import StandaloneExperiment  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(42)
model = TwoDimensionalCNNClassificationModel(StandaloneExperiment(bench.model_config("2d", 0.5)), device="cuda:0")
model.make_optimizer(max_steps=100)
model.train()
x = torch.from_numpy(bench.synth_batch(B, 0)).cuda()
y = torch.from_numpy(bench.synth_labels(B, 0)).cuda()
for i in range(steps):
    make_step(model.scheduler, step=i + 1)
    out = model(x[..., None])["class_logits"]
    loss = lsep_loss(out, y, average=False).mean()
    loss.backward()
    model.optimizer.step()
    model.optimizer.zero_grad()
    torch.cuda.synchronize()
    import fsb200
    print("step", i, "loss", float(loss), "launches", fsb200.lib().fsb_launch_count(0), flush=True)
