#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): clips/s through fused STFT -> mel -> log -> CNN forward + LSEP + backward +
Adam-amsgrad on synthetic 44.1 kHz clips.

    python bench.py --gpus 1 --steps 10 --warmup 3                # configs[1]: 2D CNN, batch 64 x 10 s per GPU (default)
    python bench.py --config 1d                                   # configs[2]: 1D CNN on raw STFT win 256 / hop 128
    python bench.py --config mixup_dp --gpus 8                    # configs[3]: 32 clips per GPU, device MixUp p = 0.5
    python bench.py --config infer_sweep --clips 20000 --gpus 8   # configs[4]: length-bucketed inference, 1-30 s clips
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
    python bench.py --impl reference ...     # the reference's CPU path on this box's host cores (same metric / config)
    python bench.py --impl torch_gpu ...     # informational: the oracle's stock-PyTorch ops on the B200 (fp32, then TF32)

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = pinned HOST inputs copied every step + result
read back, through the public model API; `roofline` = conv GEMM family (tcgen05) against the measured bf16 tensor peak;
`roofline_feat` = feature kernel against measured HBM bandwidth; `cpu_baseline` = the reference algorithm timed on host
cores (bounded sample).  Weak scaling: per-GPU work is fixed, one NCCL gradient all-reduce per step for N > 1.
"""
import argparse
import csv
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
N_CLASSES = 80
# algorithmic conv FLOPs per 10 s clip (fwd + dgrad + wgrad, unpadded channels) and feature-kernel bytes
GFLOP_PER_CLIP = {"2d": 42.81, "1d": 6.34}
FEAT_MB_PER_CLIP = {"2d": 1.985, "1d": 3.542}
FEATURES = {"2d": "mel_2048_1024_128", "1d": "stft_256_128"}
DTYPES = {
    "fp32": "f32",
    "fp16x3": "fp16x3 (split-half operands hi+lo, three tcgen05 products, f32 accumulate: fp32-grade, fwd and bwd)",
    "fp16": "fp16 (single tcgen05 pass, f32 accumulate)",
    "mixed": "fp16x3 forward (split-half operands, three tcgen05 products, f32 accumulate: fp32-grade) + fp16 single-pass "
             "dgrad/wgrad with per-tensor power-of-two gradient scaling, f32 accumulate",
}
WORKLOADS = {
    "2d": "2D-CNN (5 resnet blocks, base 100, growth 1.5, 128 mel, 80 classes) STFT+mel+fwd+LSEP+bwd+Adam-amsgrad, "
          "batch %d x 10 s @ 44.1 kHz per GPU",
    "1d": "1D-CNN on raw STFT (win 256, hop 128; 5 resnet blocks, base 100, growth 1.5, 80 classes) "
          "STFT+fwd+LSEP+bwd+Adam-amsgrad, batch %d x 10 s @ 44.1 kHz per GPU",
    "mixup_dp": "2D-CNN training step with on-device MixUp (p = 0.5, OR labels) batch assembly over a resident PCM pool, "
                "LSEP, Adam-amsgrad, batch %d x 10 s per GPU, one NCCL gradient all-reduce",
    "infer_sweep": "2D-CNN eval forward + sigmoid over %d length-bucketed clips U(1 s, 30 s), zero-padded per batch "
                   "(BucketingSampler semantics), batches sharded over the ranks",
}


def model_config(kind, dropout):
    from fsb200.experiment import make_config
    return make_config(features=FEATURES["1d" if kind == "1d" else "2d"], num_conv_blocks=5, conv_base_depth=100,
                       growth_rate=1.5, start_deep_supervision_on=1, output_dropout=dropout, n_classes=N_CLASSES)


def synth_batch(batch, seed, seconds=10):
    """white noise + one chirp per clip (fast to generate; amplitude like real audio)."""
    rng = np.random.RandomState(seed)
    t_len = int(SR * seconds)
    x = 0.05 * rng.randn(batch, t_len).astype(np.float32)
    t = np.arange(t_len, dtype=np.float32) / SR
    for i in range(batch):
        f0, f1 = rng.uniform(80, 4000), rng.uniform(80, 8000)
        x[i] += (0.3 * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / seconds))).astype(np.float32)
    return x


def synth_labels(batch, seed):
    """float32 multi-hot (batch, 80), 1-3 positives per row (lwlrap needs >= 1)."""
    rng = np.random.RandomState(seed + 1)
    labels = np.zeros((batch, N_CLASSES), dtype=np.float32)
    for i in range(batch):
        labels[i, rng.choice(N_CLASSES, size=rng.randint(1, 4), replace=False)] = 1.0
    return labels


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
        return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel_substr, pick="max"):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` export
    profiles/r02_ncu_raw.csv (tools/profile_round.sh writes it); None when the capture is not in the tree."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_raw.csv")
    if not os.path.isfile(path):
        return None
    try:
        rows = list(csv.reader(open(path)))
        header = next(r for r in rows if "Kernel Name" in r)
        ki = header.index("Kernel Name")
        ri, wi = header.index("dram__bytes_read.sum"), header.index("dram__bytes_write.sum")
        units = rows[rows.index(header) + 1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

        def val(r, i):
            return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        names = kernel_substr if isinstance(kernel_substr, (tuple, list)) else (kernel_substr,)
        vals = [val(r, ri) + val(r, wi) for r in rows[rows.index(header) + 2:]
                if len(r) == len(header) and any(n in r[ki] for n in names)]
        if not vals:
            return None
        return max(vals) if pick == "max" else sum(vals) / len(vals)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's algorithm on host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(kind, batch, device="cpu", seed=42):
    """One training step of the reference algorithm (forward incl. feature extraction, LSEP, backward, Adam-amsgrad) on
    `device`.  Uses the reference's OWN modules through the import shim when a checkout is reachable
    (/root/reference in the build container, baseline/_ref if one travelled), else the oracle's restatement ("port")."""
    from oracle import reference_shim, restate
    two_d = kind != "1d"
    config = model_config(kind, 0.0)
    signal = torch.from_numpy(synth_batch(batch, 0))[..., None].to(device)
    labels = torch.from_numpy(synth_labels(batch, 0)).to(device)
    root = reference_shim.find_reference_root()
    if root is not None and device == "cpu":
        try:
            ref = reference_shim.ReferenceModules(root)
            cls = ref.classifiers.TwoDimensionalCNNClassificationModel if two_d else ref.classifiers.HierarchicalCNNClassificationModel
            torch.manual_seed(seed)
            model = cls(reference_shim.FakeExperiment(config), device="cpu")
            model.make_optimizer(max_steps=1000)
            model.train()

            def step():
                model.optimizer.zero_grad()
                out = model(signal)["class_logits"]
                loss = ref.losses.lsep_loss(out.reshape(batch, -1), labels, average=not two_d)
                loss = loss.mean()
                loss.backward()
                model.optimizer.step()
                return loss.item()
            step()
            return step, "reference"
        except Exception as exc:       # noqa: BLE001
            print("reference modules not usable here (%r): timing the oracle port" % (exc,), file=sys.stderr)
    sd = restate.init_state_dict(config, two_d=two_d, seed=seed)
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k]
    params = {k: (v.clone().to(device).requires_grad_() if k in names else v.to(device)) for k, v in sd.items()}
    opt = torch.optim.Adam([params[k] for k in names], lr=1e-3, amsgrad=True)
    fwd = restate.net2d_forward if two_d else restate.net1d_forward

    def step():
        opt.zero_grad()
        out = fwd(params, config, signal, training=True)
        loss = restate.lsep_loss(out, labels, average=False).mean()
        loss.backward()
        opt.step()
        return loss.item()
    return step, "port"


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on this box's host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.set_num_threads(os.cpu_count())
    kind = "1d" if args.config == "1d" else "2d"
    batch = 8
    step, how = cpu_step_fn(kind, batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    cores = os.cpu_count()
    sample = ("each step = %d clips of 10 s (a bounded sample of the %d-clip batch; same model, same step definition: "
              "features + forward + LSEP + backward + Adam), %d timed steps, torch CPU fp32, %d threads, %s"
              % (batch, args.batch or 64, args.steps, cores,
                 "the reference's own modules through oracle/reference_shim.py" if how == "reference"
                 else "oracle/restate.py port (no reference checkout on this box)"))
    cfg = workload_config(args, kind, args.batch or 64, args.gpus)
    cfg["reference_arm_batch"] = batch
    line = {
        "impl": "reference", "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": how, "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_torch_gpu(args):
    """Informational: the oracle's stock-PyTorch ops (cuFFT / cuDNN / cuBLAS) on one B200, strict
    fp32 first, then with TF32 allowed -- the number the hand-written kernels should beat."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    kind = "1d" if args.config == "1d" else "2d"
    batch = args.batch or 64
    out = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        step, _ = cpu_step_fn(kind, batch, device="cuda")
        for _ in range(max(2, args.warmup)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out["tf32" if tf32 else "fp32"] = {"clips_per_s": batch * 1e3 / ms, "ms_per_step": ms}
    line = {"impl": "torch_gpu", "metric": "clips_per_sec", "value": out["fp32"]["clips_per_s"], "unit": "clips/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": out["fp32"]["ms_per_step"],
            "higher_is_better": True, "dtype": "f32 (cudnn/cublas allow_tf32 = False); `tf32` key: allow_tf32 = True",
            "data": "synthetic", "config": workload_config(args, kind, batch, 1), "tf32": out["tf32"],
            "note": "oracle/restate.py torch ops on cuda:0 (torch.stft, F.conv1d/conv2d, batch_norm, prelu, max_pool, "
                    "torch.optim.Adam amsgrad), same step definition as the product arm"}
    print(json.dumps(line), flush=True)


def workload_config(args, kind, batch, n_gpus, extra=None):
    name = args.config
    cfg = {
        "workload": WORKLOADS[name] % (args.clips if name == "infer_sweep" else batch),
        "global_batch": batch * n_gpus, "per_gpu_batch": batch, "clip_seconds": 10,
        "features": FEATURES[kind], "parallelism": "dp%d" % n_gpus, "precision": args.precision,
        "l2": "inputs rotate over 4 distinct batches (> 126 MB L2 together); each step also streams >10 GB of "
              "activations, so nothing survives in L2 between timed steps",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
class Harness:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = "cuda:%d" % self.local_rank
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.device))
        os.environ["FSB200_PRECISION"] = args.precision

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, *a):
        """CUDA-event time of fn(*a) on the launch stream, barrier + synchronize on both sides, max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*a)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.device)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def phase_profile(h, plan, run_step, nprof=3):
    """Per-kernel-family device time (CUDA events on the launch stream) with the side-stream overlap of weight-gradient
    GEMMs / weight packing switched off, so that each family's time is its own device time.  Every rank runs the steps
    (they contain the gradient all-reduce); rank 0 records."""
    acc = {}
    plan.set_overlap(False)
    if h.rank == 0:
        plan.set_profiling(True)
    for i in range(nprof):
        run_step(i)
        torch.cuda.synchronize()
        if h.rank == 0:
            for name, (ms, fl) in plan.timings().items():
                a = acc.setdefault(name, [0.0, 0.0])
                a[0] += ms / nprof
                a[1] += fl / nprof
    plan.set_overlap(True)
    if h.rank == 0:
        plan.set_profiling(False)
    return acc


def rooflines(args, acc, kind, batch, value_per_gpu, peaks, training=True):
    fams = ("gemm_fwd", "gemm_dgrad", "gemm_wgrad") if training else ("gemm_fwd",)
    gemm_ms = sum(acc[k][0] for k in fams)
    gemm_fl = sum(acc[k][1] for k in fams)
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    n_launches = 19 * len(fams)
    products = {"fp16x3": "3 MMAs per algorithmic one in every GEMM (ceiling = peak / 3)",
                "mixed": "3 MMAs per algorithmic one in the forward GEMMs, 1 in dgrad / wgrad (ceiling = peak x 3/5)",
                "fp16": "1 MMA per algorithmic one", "fp32": "CUDA-core GEMMs"}[args.precision]
    roofline = {
        "bound": "tensor",
        "kernel": "conv_tc_kernel + wgrad_tc_kernel (tcgen05 row-shifted conv GEMMs: %s of the 19 convs, %s)"
                  % (" + ".join(f.replace("gemm_", "") for f in fams), args.precision),
        "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor"],
        # DRAM bytes of one launch of the dominant kernel (largest conv_tc_kernel launch) from the committed
        # `ncu --set full` export of this round, when present
        "traffic": ncu_traffic("conv_tc_kernel"),
        "peak_source": "%s bf16 sustained (kernel timed inside a long step); %s" % (peaks["source"], products),
        "algorithmic_gflop_per_step": gemm_fl / 1e9, "algorithmic_gflop_per_launch": gemm_fl / 1e9 / n_launches,
        "ms_per_step": gemm_ms, "ms_per_launch": gemm_ms / n_launches,
    }
    if training:
        roofline["whole_step_frac"] = value_per_gpu * GFLOP_PER_CLIP[kind] * 1e9 / (peaks["tensor"] * 1e12)
    feat_ms = acc["feat"][0]
    feat_bytes = FEAT_MB_PER_CLIP[kind] * 1e6 * batch
    feat_gbs = feat_bytes / (feat_ms * 1e-3) / 1e9 if feat_ms > 0 else 0.0
    roofline_feat = {"bound": "hbm", "kernel": "%s (framing + Hann + rFFT%s + log)" % (
                         "feat2048_mel_kernel" if kind == "2d" else "feat_kernel", " + mel" if kind == "2d" else ""),
                     "achieved": feat_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": feat_gbs / peaks["hbm"],
                     "traffic": ncu_traffic(("feat2048_mel_kernel", "feat_kernel"), "mean"), "ms_per_launch": feat_ms,
                     "algorithmic_mb_per_launch": feat_bytes / 1e6, "peak_source": peaks["source"]}
    return roofline, roofline_feat


def cpu_baseline(kind):
    """The reference algorithm on this box's host cores: bounded sample (about 10-30 s of CPU work)."""
    cb = 8
    torch.set_num_threads(os.cpu_count())
    step, how = cpu_step_fn(kind, cb)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    reps = max(1, min(4, int(20.0 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = time.perf_counter() - t0
    return {"value": cb * reps / dt, "unit": "clips/s", "cores": os.cpu_count(), "kind": how,
            "sample": "%d steps x %d clips of 10 s (bounded sample of the batch; features + forward + LSEP + backward + "
                      "Adam), torch CPU fp32, %d threads, after 1 warm-up step" % (reps, cb, os.cpu_count())}


def run_training(args):
    """configs[1] (2d), configs[2] (1d) and configs[3] (mixup_dp)."""
    import fsb200
    from fsb200.experiment import StandaloneExperiment
    from networks.classifiers import HierarchicalCNNClassificationModel, TwoDimensionalCNNClassificationModel
    from networks.losses import lsep_loss
    from ops.training import make_step

    h = Harness(args)
    kind = "1d" if args.config == "1d" else "2d"
    mixup = args.config == "mixup_dp"
    batch = args.batch or (32 if mixup else 64)
    device, rank, world = h.device, h.rank, h.world

    torch.manual_seed(42)
    cls = HierarchicalCNNClassificationModel if kind == "1d" else TwoDimensionalCNNClassificationModel
    model = cls(StandaloneExperiment(model_config(kind, 0.5)), device=device)
    model.make_optimizer(max_steps=4 * (args.warmup + args.steps) + 16)
    model.train()

    n_rot = 4
    host = [torch.from_numpy(synth_batch(batch, 1000 * rank + i)).pin_memory() for i in range(n_rot)]
    host_labels = [torch.from_numpy(synth_labels(batch, 1000 * rank + i)).pin_memory() for i in range(n_rot)]
    step_counter = [0]

    def train_step(signal, labels):
        step_counter[0] += 1
        make_step(model.scheduler, step=step_counter[0])
        out = model(signal)["class_logits"]
        if kind == "1d":
            loss = lsep_loss(out.reshape(labels.shape), labels, average=True)      # reference :347-350 (1D loop)
        else:
            loss = lsep_loss(out, labels, average=False).mean()                     # reference :668-677
        loss.backward()
        model._sync_gradients()
        model.optimizer.step()
        model.optimizer.zero_grad()
        return loss

    if mixup:
        # resident PCM pool (all rotation batches of this rank) + per-step draws in the reference's RNG order
        from fsb200.assemble import DeviceBatchAssembler, DevicePcmPool
        clips = [row for hb in host for row in hb.numpy()]
        pool = DevicePcmPool(clips, np.concatenate([hl.numpy() for hl in host_labels]), device=device)
        assembler = DeviceBatchAssembler(pool, p_mixup=0.5, np_rng=np.random.RandomState(42 + rank),
                                         py_rng=__import__("random").Random(42 + rank))
        order = np.random.RandomState(7 + rank)

        def resident_step(i):
            indices = order.choice(len(pool), size=batch, replace=False)
            signal, labels = assembler.assemble(indices)
            return train_step(signal, labels)
    else:
        dev = [hb.to(device) for hb in host]
        dev_labels = [hl.to(device) for hl in host_labels]

        def resident_step(i):
            return train_step(dev[i % n_rot][..., None], dev_labels[i % n_rot])

    def resident(steps):
        for i in range(steps):
            resident_step(i)

    resident(args.warmup)
    sampler = ClockSampler(h.local_rank)
    if rank == 0:
        sampler.start()
    fsb200.lib().fsb_launch_count(1)
    ms_total = h.timed(resident, args.steps)
    launches = fsb200.lib().fsb_launch_count(0)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- end to end: pinned host -> device every step (prefetched on a copy stream), loss read back
    copy_stream = torch.cuda.Stream()
    stage = [torch.empty((batch, host[0].shape[1]), dtype=torch.float32, device=device) for _ in range(2)]
    stage_labels = [torch.empty((batch, N_CLASSES), dtype=torch.float32, device=device) for _ in range(2)]
    loss_host = torch.empty(args.steps + args.warmup + 1, dtype=torch.float32).pin_memory()
    if mixup:
        from fsb200.assemble import DevicePcmPool as _Pool

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
            if mixup:
                # the step's clips arrive from pinned host memory; MixUp partners are drawn among them on the device
                staged = _Pool.__new__(_Pool)
                staged.device = torch.device(device)
                staged.lengths = np.full(batch, stage[s].shape[1], dtype=np.int64)
                staged.offsets = np.arange(batch, dtype=np.int64) * stage[s].shape[1]
                staged.pcm, staged.labels = stage[s].reshape(-1), stage_labels[s]
                assembler.pool = staged
                signal, labels = assembler.assemble(np.arange(batch))
                assembler.pool = pool
                loss = train_step(signal, labels)
            else:
                loss = train_step(stage[s][..., None], stage_labels[s])
            consumed[s].record()
            loss_host[i].copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_host[steps - 1])

    e2e(2)
    ms_e2e = h.timed(e2e, args.steps)

    value = batch * world * args.steps / (ms_total / 1e3)
    e2e_value = batch * world * args.steps / (ms_e2e / 1e3)
    h2d = host[0].numel() * 4 + host_labels[0].numel() * 4
    peaks = measured_peaks()

    acc = phase_profile(h, model._plan, resident_step)
    line = None
    if rank == 0:
        phases = {k: {"ms": round(v[0], 4), "gflop": round(v[1] / 1e9, 2)} for k, v in acc.items()}
        roofline, roofline_feat = rooflines(args, acc, kind, batch, value / world, peaks)
        cpu = cpu_baseline(kind) if (world == 1 and not args.no_cpu_baseline) else None
        extra = {"mixup_p": 0.5, "pcm_pool_clips": len(pool)} if mixup else None
        line = {
            "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPES[args.precision], "data": "synthetic",
            "config": workload_config(args, kind, batch, world, extra), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_feat": roofline_feat, "phases_ms": phases,
            "phases_note": "per-family device time of one step with the wgrad / weight-pack side-stream overlap OFF "
                           "(sum > ms_per_step, which is measured with the overlap ON)",
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    h.finish()


def run_infer_sweep(args):
    """configs[4]: `args.clips` clips with lengths ~ U(1 s, 30 s), binned by length (BucketingSampler semantics:
    np.digitize + max_batch_elems), zero-padded per batch, eval forward + sigmoid, batches sharded over the ranks.
    A 'step' is one padded batch; the timed region is this rank's whole share of the sweep."""
    import fsb200
    from fsb200.experiment import StandaloneExperiment
    from fsb200.inference import pack_batches, predict_bucketed
    from networks.classifiers import TwoDimensionalCNNClassificationModel

    h = Harness(args)
    device, rank, world = h.device, h.rank, h.world
    torch.manual_seed(42)
    model = TwoDimensionalCNNClassificationModel(StandaloneExperiment(model_config("2d", 0.5)), device=device)
    model.eval()

    # clips are prefixes of a small bank of distinct 30 s waveforms (generating 20k independent clips = 55 GB of noise
    # would dominate the run); lengths are drawn once, seed 42
    bank_n = 32
    bank = torch.from_numpy(synth_batch(bank_n, 4242, seconds=30)).pin_memory()
    rng = np.random.RandomState(42)
    lengths = rng.randint(1 * SR, 30 * SR + 1, size=args.clips)
    source = rng.randint(0, bank_n, size=args.clips)
    clips = [bank[source[i], :lengths[i]] for i in range(args.clips)]           # views of pinned host memory
    buckets = [int(s * SR) for s in (1, 2, 3, 4, 5, 6, 8, 10, 12, 15, 18, 22, 26, 30)] + [30 * SR + 1]
    max_batch_elems = 64 * 10 * SR                                              # one canonical batch worth of samples
    batches, _ = pack_batches(lengths, buckets, max_batch_elems)
    n_batches = len(batches)

    dev_bank = bank.to(device)
    dev_clips = [dev_bank[source[i], :lengths[i]] for i in range(args.clips)]

    predict_bucketed(model, dev_clips[:256], buckets, max_batch_elems)          # warm-up (plans, tensor maps, tables)
    sampler = ClockSampler(h.local_rank)
    if rank == 0:
        sampler.start()
    fsb200.lib().fsb_launch_count(1)
    stats = {}
    ms_total = h.timed(lambda: stats.update(predict_bucketed(model, dev_clips, buckets, max_batch_elems, return_stats=True)[1]))
    launches = fsb200.lib().fsb_launch_count(0)
    clocks = sampler.stop() if rank == 0 else {}
    ms_e2e = h.timed(lambda: predict_bucketed(model, clips, buckets, max_batch_elems))

    value = args.clips / (ms_total / 1e3)
    e2e_value = args.clips / (ms_e2e / 1e3)
    peaks = measured_peaks()

    # roofline of the forward GEMMs on one canonical batch (64 x 10 s) in eval mode
    probe = dev_bank[:, :10 * SR].repeat(2, 1)[:64, :, None].contiguous()

    def probe_step(i):
        with torch.no_grad():
            model(probe)
    acc = phase_profile(h, model._plan, probe_step)
    if rank == 0:
        phases = {k: {"ms": round(v[0], 4), "gflop": round(v[1] / 1e9, 2)} for k, v in acc.items()}
        roofline, roofline_feat = rooflines(args, acc, "2d", 64, value / world, peaks, training=False)
        real = int(lengths.sum())
        line = {
            "metric": "clips_per_sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": n_batches,
            "warmup": args.warmup, "ms_per_step": ms_total / max(1, n_batches // world), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": DTYPES[args.precision], "data": "synthetic",
            "config": workload_config(args, "2d", 64, world, {
                "clips": args.clips, "clip_seconds": "U(1, 30)", "buckets_s": [b / SR for b in buckets],
                "max_batch_elems": max_batch_elems, "batches": n_batches,
                "padding_overhead": stats.get("padding_overhead"), "mean_clip_seconds": real / args.clips / SR,
                "audio_seconds_per_s": real / SR / (ms_total / 1e3)}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(4 * real / max(1, n_batches)),
                    "d2h_bytes_per_step": int(args.clips * N_CLASSES * 4 / max(1, n_batches)),
                    "ms_per_step": ms_e2e / max(1, n_batches // world)},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_feat": roofline_feat, "phases_ms": phases,
            "phases_note": "roofline / phases: eval forward of one canonical 64 x 10 s batch; value / e2e: the whole sweep",
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    h.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--config", default="2d", choices=["2d", "1d", "mixup_dp", "infer_sweep"])
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (default 64; 32 for mixup_dp)")
    ap.add_argument("--clips", type=int, default=20000, help="infer_sweep: number of clips")
    ap.add_argument("--precision", default=os.environ.get("FSB200_PRECISION", "mixed"),
                    choices=["fp32", "fp16x3", "fp16", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args)
    elif args.config == "infer_sweep":
        run_infer_sweep(args)
    else:
        run_training(args)


if __name__ == "__main__":
    main()
