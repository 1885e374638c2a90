"""Host-side profile of one training step (run on the GPU box): where does the Python/driver time go?"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "freesound-classification_b200"))
import torch  # noqa: E402

import bench  # noqa: E402

os.environ.setdefault("FSB200_PRECISION", "bf16x3")
from networks.classifiers import TwoDimensionalCNNClassificationModel  # noqa: E402
from networks.losses import lsep_loss  # noqa: E402
from ops.training import make_step  # noqa: E402
from oracle.reference_shim import FakeExperiment  # noqa: E402

torch.manual_seed(42)
model = TwoDimensionalCNNClassificationModel(FakeExperiment(bench.canonical_config(0.5)), device="cuda:0")
model.make_optimizer(max_steps=100)
model.train()
B = int(os.environ.get("B", "64"))
x = torch.from_numpy(bench.synth_batch(B, 0)).cuda()
y = torch.from_numpy(bench.synth_labels(B, 0)).cuda()
n = [0]


def step():
    n[0] += 1
    make_step(model.scheduler, step=n[0])
    out = model(x[..., None])["class_logits"]
    loss = lsep_loss(out, y, average=False).mean()
    loss.backward()
    model._sync_gradients()
    model.optimizer.step()
    model.optimizer.zero_grad()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("enqueue ms/step %.2f   total ms/step %.2f" % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
