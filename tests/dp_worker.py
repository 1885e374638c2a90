"""Worker of tests/test_gpu_round2.py::test_two_gpu_allreduce_matches_single_gpu_shards (launched under torchrun with
2 ranks, one GPU each).  Every rank runs forward + backward on its shard of a seeded batch and all-reduces the flat
gradient over NCCL; rank 0 then recomputes both shards alone on its GPU and checks that the all-reduced gradient equals
the sum of the two single-GPU shard gradients BIT FOR BIT (per-rank BatchNorm statistics, SURVEY.md 8e option 1), and
that a replica built under a different seed was overwritten by rank 0's parameters and buffers."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "freesound-classification_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from fsb200.experiment import StandaloneExperiment, make_config
    from networks.classifiers import TwoDimensionalCNNClassificationModel
    from networks.losses import lsep_loss
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    os.environ["FSB200_PRECISION"] = "mixed"
    cfg = make_config(conv_base_depth=16, growth_rate=1.5)
    torch.manual_seed(42 + 1000 * rank)               # deliberately different initial weights per rank
    model = TwoDimensionalCNNClassificationModel(StandaloneExperiment(cfg), device="cuda:%d" % local)
    model.make_optimizer(max_steps=10)                # broadcasts rank 0's parameters and buffers
    sd0 = [torch.zeros_like(t) for t in model.state_dict().values()]
    for t, src in zip(sd0, model.state_dict().values()):
        t.copy_(src)
        dist.broadcast(t, src=0)
    same = all(torch.equal(a, b) for a, b in zip(sd0, model.state_dict().values()))

    per = 6
    rng = np.random.RandomState(7)
    signal = torch.from_numpy((0.1 * rng.randn(world * per, 40000)).astype(np.float32)).cuda()
    labels = torch.from_numpy((rng.rand(world * per, 80) < 0.03).astype(np.float32)).cuda()
    labels[:, 0] = 1.0
    start = {k: v.clone() for k, v in model.state_dict().items()}

    def shard_gradient(r):
        model.load_state_dict(start)
        model.train()
        for p in model.parameters():
            p.grad = None
        out = model(signal[r * per:(r + 1) * per, :, None])["class_logits"]
        lsep_loss(out, labels[r * per:(r + 1) * per], average=False).mean().backward()
        return model._plan.last_flat_grad

    flat = shard_gradient(rank)
    model._sync_gradients()
    reduced = flat.clone()
    ok = True
    if rank == 0:
        total = torch.zeros_like(reduced)
        for r in range(world):
            total += shard_gradient(r)
        ok = bool(torch.equal(total, reduced)) and float(reduced.abs().max()) > 0
        print("DP_CHECK replicas_synced=%s allreduce_bit_exact=%s grad_scale=%r" % (
            same, ok, getattr(model.optimizer, "grad_scale", None)), flush=True)
    flag = torch.tensor([1.0 if (ok and same) else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
