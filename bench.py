#!/usr/bin/env python
"""Benchmark of the hot path: clips/s through fused STFT -> mel -> log -> 2D CNN forward + LSEP +
backward + Adam-amsgrad on synthetic 10 s @ 44.1 kHz clips (BASELINE.json configs[1]; batch 64 per GPU,
weak scaling with one NCCL gradient all-reduce per step for N > 1).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle port) on host cores

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = pinned host inputs copied
every step + loss read back inside the timed region; `roofline` = conv GEMM family against the measured
bf16 tensor peak; `roofline_feat` = feature kernel against measured HBM bandwidth; `cpu_baseline` = the
oracle timed on this box's host cores (bounded sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "freesound-classification_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SR = 44100
CLIP_SECONDS = 10
T = SR * CLIP_SECONDS
N_CLASSES = 80
CONV_GFLOP_PER_CLIP = 42.81      # canonical config, fwd + dgrad + wgrad, unpadded channels
FEAT_MB_PER_CLIP = 1.985         # PCM read + log-mel write


DTYPES = {
    "fp32": "f32",
    "fp16x3": "fp16x3 (split-half operands hi+lo, three tcgen05 products, f32 accumulate: fp32-grade, fwd and bwd)",
    "fp16": "fp16 (single tcgen05 pass, f32 accumulate)",
    "mixed": "fp16x3 forward (split-half operands, three tcgen05 products, f32 accumulate: fp32-grade) + fp16 single-pass "
             "dgrad/wgrad with per-tensor power-of-two gradient scaling, f32 accumulate",
}


def canonical_config(dropout):
    from oracle.reference_shim import make_config
    return make_config(features="mel_2048_1024_128", num_conv_blocks=5, conv_base_depth=100, growth_rate=1.5,
                       start_deep_supervision_on=1, output_dropout=dropout, n_classes=N_CLASSES)


def workload_config(batch, n_gpus, precision, extra=None):
    cfg = {
        "workload": "2D-CNN (5 resnet blocks, base 100, growth 1.5, 128 mel, 80 classes) STFT+mel+fwd+LSEP+bwd+"
                    "Adam-amsgrad, batch %d x 10 s @ 44.1 kHz per GPU" % batch,
        "global_batch": batch * n_gpus, "per_gpu_batch": batch, "clip_seconds": CLIP_SECONDS,
        "features": "mel_2048_1024_128", "parallelism": "dp%d" % n_gpus, "precision": precision,
        "l2": "inputs rotate over 4 distinct batches (452 MB > 126 MB L2); each step also streams >10 GB of "
              "activations, so nothing survives in L2 between timed steps",
    }
    if extra:
        cfg.update(extra)
    return cfg


def synth_batch(batch, seed):
    """white noise + one chirp per clip (fast to generate; amplitude like real audio)."""
    rng = np.random.RandomState(seed)
    x = 0.05 * rng.randn(batch, T).astype(np.float32)
    t = np.arange(T, dtype=np.float32) / SR
    for i in range(batch):
        f0, f1 = rng.uniform(80, 4000), rng.uniform(80, 8000)
        x[i] += (0.3 * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / CLIP_SECONDS))).astype(np.float32)
    return x


def synth_labels(batch, seed):
    from oracle import restate
    return restate.synth_labels(batch, N_CLASSES, seed=seed)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, sm_max, power = [], set(), None, []
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    sm_max = float(f[2])
                    power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=sm_max, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops_sustained"], tensor_burst=d["bf16_tflops"], source="measured")
    except Exception:
        return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source="fallback")


# ------------------------------------------------------------------------------------------------
def cpu_port_step_fn(batch, dropout=0.0, seed=42):
    """The oracle's CPU restatement of the same step (forward incl. feature extraction, LSEP, backward,
    Adam-amsgrad), all host threads."""
    from oracle import restate
    config = canonical_config(dropout)
    torch.set_num_threads(os.cpu_count())
    sd = restate.init_state_dict(config, two_d=True, seed=seed)
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    params = {k: (v.clone().requires_grad_() if k in names else v) for k, v in sd.items()}
    opt = torch.optim.Adam([params[k] for k in names], lr=1e-3, amsgrad=True)
    signal = torch.from_numpy(synth_batch(batch, 0))[..., None]
    labels = torch.from_numpy(synth_labels(batch, 0))

    def step():
        opt.zero_grad()
        out = restate.net2d_forward(params, config, signal, training=True)
        loss = restate.lsep_loss(out, labels, average=False).mean()
        loss.backward()
        opt.step()
        return loss.item()

    return step


def run_reference(args):
    """`--impl reference`: the reference's algorithm on host cores (oracle port; the reference itself has no
    packaging metadata so it cannot be pip-installed into baseline/_ref -- see DESIGN.md)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 8
    step = cpu_port_step_fn(batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(64, args.gpus, args.precision),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port",
                         "sample": "each step = %d of the 64 clips (10 s each, same model); %d timed steps, torch CPU fp32 "
                                   "oracle, %d threads" % (batch, args.steps, cores)},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import fsb200
    from networks.classifiers import TwoDimensionalCNNClassificationModel
    from networks.losses import lsep_loss
    from ops.training import make_step
    from oracle.reference_shim import FakeExperiment

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    os.environ["FSB200_PRECISION"] = args.precision
    batch = args.batch

    torch.manual_seed(42)
    model = TwoDimensionalCNNClassificationModel(FakeExperiment(canonical_config(0.5)), device=device)
    total_steps = 4 * (args.warmup + args.steps) + 16
    model.make_optimizer(max_steps=total_steps)
    model.train()

    n_rot = 4
    host = [torch.from_numpy(synth_batch(batch, 1000 * rank + i)).pin_memory() for i in range(n_rot)]
    host_labels = [torch.from_numpy(synth_labels(batch, 1000 * rank + i)).pin_memory() for i in range(n_rot)]
    dev = [h.to(device) for h in host]
    dev_labels = [h.to(device) for h in host_labels]
    step_counter = [0]

    def train_step(signal, labels):
        step_counter[0] += 1
        make_step(model.scheduler, step=step_counter[0])
        out = model(signal[..., None])["class_logits"]
        loss = lsep_loss(out, labels, average=False).mean()
        loss.backward()
        model._sync_gradients()
        model.optimizer.step()
        model.optimizer.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident inputs
    def resident(steps):
        for i in range(steps):
            train_step(dev[i % n_rot], dev_labels[i % n_rot])

    resident(args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    fsb200.lib().fsb_launch_count(1)
    ms_total = timed(resident, args.steps)
    launches = fsb200.lib().fsb_launch_count(0)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- end to end: pinned host -> device every step (prefetched on a copy stream), loss read back
    copy_stream = torch.cuda.Stream()
    stage = [torch.empty_like(dev[0]) for _ in range(2)]
    stage_labels = [torch.empty_like(dev_labels[0]) for _ in range(2)]
    loss_host = torch.empty(args.steps + args.warmup + 1, dtype=torch.float32).pin_memory()

    def e2e(steps):
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])
                stage[s].copy_(host[i % n_rot], non_blocking=True)
                stage_labels[s].copy_(host_labels[i % n_rot], non_blocking=True)
                ready[s].record(copy_stream)

        for s in range(2):
            consumed[s].record()
        prefetch(0)
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            s = i % 2
            torch.cuda.current_stream().wait_event(ready[s])
            loss = train_step(stage[s], stage_labels[s])
            consumed[s].record()
            loss_host[i].copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host[steps - 1])

    e2e(2)
    ms_e2e = timed(e2e, args.steps)

    value = batch * world * args.steps / (ms_total / 1e3)
    e2e_value = batch * world * args.steps / (ms_e2e / 1e3)
    h2d = host[0].numel() * 4 + host_labels[0].numel() * 4
    peaks = measured_peaks()

    # ---- per-kernel-family device time (CUDA events on the launch stream) for the roofline
    # (every rank runs these steps -- they contain the gradient all-reduce -- but only rank 0 records)
    roofline = roofline_feat = None
    phases = {}
    plan = model._plan
    acc = {}
    nprof = 3
    # the per-family times are taken with the side-stream overlap of weight-gradient GEMMs / weight packing switched
    # off, so that each family's CUDA-event time is its own device time (overlapped, the families' times add up to more
    # than the step); `value` / `e2e` above are measured with the overlap on
    plan.set_overlap(False)
    if rank == 0:
        plan.set_profiling(True)
    for i in range(nprof):
        train_step(dev[i % n_rot], dev_labels[i % n_rot])
        torch.cuda.synchronize()
        if rank == 0:
            for name, (ms, fl) in plan.timings().items():
                a = acc.setdefault(name, [0.0, 0.0])
                a[0] += ms / nprof
                a[1] += fl / nprof
    plan.set_overlap(True)
    if rank == 0:
        plan.set_profiling(False)
        phases = {k: {"ms": round(v[0], 4), "gflop": round(v[1] / 1e9, 2)} for k, v in acc.items()}
        gemm_ms = sum(acc[k][0] for k in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad"))
        gemm_fl = sum(acc[k][1] for k in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad"))
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        n_gemm_launches = 57          # 19 convs x (forward + dgrad + wgrad) tcgen05 launches per step
        roofline = {
            "bound": "tensor", "kernel": "conv_tc_kernel + wgrad_tc_kernel (tcgen05 row-shifted conv GEMMs: fwd + dgrad + "
                                         "wgrad of the 19 convs, %s)" % args.precision,
            "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor"],
            # DRAM bytes of one launch (block-0 3x3 forward, the largest layer) from the committed ncu --set full
            # capture profiles/r01_ncu_summary.txt; algorithmic bytes of that launch = 2 x 411 MB (read a, write z)
            "traffic": 778.8e6,
            "peak_source": "%s bf16 sustained (kernel timed inside a long step); bf16x3 issues 3 MMAs per algorithmic "
                           "one, so its ceiling is peak / 3" % peaks["source"],
            "algorithmic_gflop_per_step": gemm_fl / 1e9, "algorithmic_gflop_per_launch": gemm_fl / 1e9 / n_gemm_launches,
            "ms_per_step": gemm_ms, "ms_per_launch": gemm_ms / n_gemm_launches,
            "frac_of_bf16x3_ceiling": 3.0 * achieved / peaks["tensor"],
            "whole_step_frac": value / world * CONV_GFLOP_PER_CLIP * 1e9 / (peaks["tensor"] * 1e12),
        }
        feat_ms = acc["feat"][0]
        feat_bytes = FEAT_MB_PER_CLIP * 1e6 * batch
        feat_gbs = feat_bytes / (feat_ms * 1e-3) / 1e9 if feat_ms > 0 else 0.0
        roofline_feat = {"bound": "hbm", "kernel": "feat_kernel (STFT+mel+log)", "achieved": feat_gbs,
                         "peak": peaks["hbm"], "unit": "GB/s", "frac": feat_gbs / peaks["hbm"], "traffic": None,
                         "ms_per_launch": feat_ms, "peak_source": peaks["source"]}

    # ---- CPU baseline (oracle port on this box's host cores), rank 0 at N = 1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = 8
        step = cpu_port_step_fn(cb)
        t0 = time.perf_counter()
        step()
        first = time.perf_counter() - t0
        reps = max(1, min(4, int(20.0 / max(first, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": cb * reps / dt, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": "%d steps x %d of the 64 clips (10 s each), torch CPU fp32 oracle, %d threads, "
                                  "after 1 warm-up step" % (reps, cb, os.cpu_count())}

    if rank == 0:
        line = {
            "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPES[args.precision],
            "data": "synthetic", "config": workload_config(batch, world, args.precision),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_feat": roofline_feat, "phases_ms": phases,
            "phases_note": "per-family device time of one step with the wgrad / weight-pack side-stream overlap OFF "
                           "(sum > ms_per_step, which is measured with the overlap ON)",
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default=os.environ.get("FSB200_PRECISION", "mixed"),
                    choices=["fp32", "fp16x3", "fp16", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
