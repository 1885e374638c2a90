"""Drop-in replacements for the reference's `networks/classifiers.py` model classes.

`TwoDimensionalCNNClassificationModel` and `HierarchicalCNNClassificationModel` keep the reference's
constructor, `forward(signal) -> {"class_logits"}`, training / evaluation loops, optimizer wiring and
-- because the reference's exact `nn.Module` tree is kept as the PARAMETER CONTAINER -- its 200-key
`state_dict`, default-init RNG consumption and `parameters()` order (reference
networks/classifiers.py:497-549, :120-173).  Only the arithmetic moved: `forward` hands raw pointers of
the parameters / BN buffers to the sm_100a plan in libfsb200.so (fused STFT->mel->log kernel, row-shifted
GEMM convolutions, fused BN/PReLU/residual/pool passes, global-max heads, FC head) and autograd sees the
whole network as ONE function whose backward is a second library call.  There is no CPU path.
"""
import os
from collections import deque

import numpy as np
import torch
import torch.nn as nn
from tqdm import tqdm

from fsb200 import runtime
from networks.losses import binary_cross_entropy, focal_loss, lsep_loss  # noqa: F401  (reference imports)
from ops.training import OPTIMIZERS, make_scheduler, make_step
from ops.utils import is_mel, is_stft, lwlrap, make_mel_filterbanks


def _summary_writer(log_dir):
    try:
        from tensorboardX import SummaryWriter
        return SummaryWriter(log_dir=log_dir)
    except ImportError:
        pass
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir=log_dir)
    except Exception:
        class _Null:
            def __getattr__(self, name):
                return lambda *a, **k: None
        return _Null()


class ConvLockedDropout(nn.Module):
    """Kept for API parity (reference :21-34; unused by the two CNN models)."""

    def __init__(self, dropout_rate=0.0):
        super().__init__()
        self.dropout_rate = dropout_rate

    def forward(self, x):
        if not self.training or not self.dropout_rate:
            return x
        n, s, t = x.size()
        m = torch.zeros(n, s, 1, device=x.device).bernoulli_(1 - self.dropout_rate)
        return m.expand_as(x) * x


class _ResContainer(nn.Module):
    """Parameter container with the reference's registration order conv1, bn1, conv2, bn2, conv3, bn3,
    prelu1..3 (reference :37-104).  The arithmetic lives in the CUDA plan."""

    def __init__(self, depth, conv, bn):
        super().__init__()
        self.conv1 = conv(depth, depth, kernel_size=1)
        self.bn1 = bn(depth)
        self.conv2 = conv(depth, depth, kernel_size=3, padding=1)
        self.bn2 = bn(depth)
        self.conv3 = conv(depth, depth, kernel_size=1)
        self.bn3 = bn(depth)
        self.prelu1 = nn.PReLU(depth)
        self.prelu2 = nn.PReLU(depth)
        self.prelu3 = nn.PReLU(depth)

    def forward(self, x):
        raise RuntimeError("parameter container only: the block runs inside the fsb200 plan")


class ResnetBlock(_ResContainer):
    def __init__(self, depth):
        super().__init__(depth, nn.Conv1d, nn.BatchNorm1d)


class ResnetBlock2d(_ResContainer):
    def __init__(self, depth):
        super().__init__(depth, nn.Conv2d, nn.BatchNorm2d)


class _AcceleratedCNN(nn.Module):
    """Shared implementation of the two CNN classifiers (module tree, plan hand-off, loops)."""

    two_d = True

    def __init__(self, experiment, device="cuda"):
        super().__init__()
        self.device = device
        self.experiment = experiment
        self.config = experiment.config
        net, data = self.config.network, self.config.data
        if not (is_mel(data.features) or is_stft(data.features)):
            raise NotImplementedError("features %r: only mel_* / stft_* descriptors are accelerated" % data.features)
        if net.aggregation_type not in ("max", "rnn") or (net.aggregation_type == "rnn" and not self.two_d):
            raise NotImplementedError("aggregation_type=%r is not implemented for %s (max: both models; rnn: the 2D "
                                      "model, as in the reference)" % (net.aggregation_type, type(self).__name__))
        if self.two_d and not is_mel(data.features) and not is_stft(data.features):
            raise NotImplementedError

        self._filterbank_np = make_mel_filterbanks(data.features) if is_mel(data.features) else None
        if self._filterbank_np is not None:
            # plain tensor attribute like the reference (:493-495): not a buffer, not in the state_dict
            self.filterbanks = torch.from_numpy(self._filterbank_np).to(self.device)

        conv, bn = (nn.Conv2d, nn.BatchNorm2d) if self.two_d else (nn.Conv1d, nn.BatchNorm1d)
        pool = nn.MaxPool2d if self.two_d else nn.MaxPool1d
        res = ResnetBlock2d if self.two_d else ResnetBlock

        self.conv_modules = torch.nn.ModuleList()
        self.rnns = torch.nn.ModuleList()
        total_depth = 0
        self._depths = []
        depth = None
        for k in range(net.num_conv_blocks):
            if self.two_d:
                input_size = 2 if not k else depth
            else:
                input_size = data._input_dim if not k else depth
            depth = int(net.growth_rate ** k * net.conv_base_depth)
            self._depths.append(depth)
            if k >= net.start_deep_supervision_on:
                if net.aggregation_type == "rnn":
                    # registered (and default-initialised) BEFORE the block's own modules, like the reference :512-522
                    rnn_size = 128
                    total_depth += rnn_size * 2
                    self.rnns.append(nn.Sequential(
                        nn.LayerNorm((depth,)),
                        nn.GRU(depth, rnn_size, batch_first=True, bidirectional=True)))
                else:
                    total_depth += depth
            self.conv_modules.append(nn.Sequential(
                bn(input_size),
                conv(input_size, depth, kernel_size=3, padding=1),
                pool(kernel_size=2, stride=2),
                bn(depth),
                nn.PReLU(depth),
                res(depth)))

        self.global_maxpool = nn.AdaptiveMaxPool2d(1) if self.two_d else nn.AdaptiveMaxPool1d(1)
        self.output_transform = nn.Sequential(
            nn.BatchNorm1d(total_depth),
            nn.Linear(total_depth, total_depth),
            nn.BatchNorm1d(total_depth),
            nn.PReLU(total_depth),
            nn.Dropout(p=net.output_dropout),
            nn.Linear(total_depth, data._n_classes))

        self.to(self.device)
        self._plan = None
        self._param_list = None
        self._dropout_calls = 0
        self.global_step = 0

    # ------------------------------------------------------------------------------------------
    def _get_plan(self):
        if self._plan is None:
            dev = torch.device(self.device)
            if dev.type != "cuda":
                raise RuntimeError("%s runs on CUDA (sm_100a) only; device=%r has no implementation"
                                   % (type(self).__name__, self.device))
            net, data = self.config.network, self.config.data
            if not self.two_d:
                n_bins = int(data.features.split("_")[3]) if is_mel(data.features) else \
                    int(data.features.split("_")[1]) // 2 + 1
                if data._input_dim != n_bins:
                    raise ValueError("_input_dim=%d does not match features %r (%d)" % (data._input_dim, data.features, n_bins))
            self._plan = runtime.NetPlan(self.two_d, data.features, self._depths, net.start_deep_supervision_on,
                                         data._n_classes, net.output_dropout, filterbank=self._filterbank_np,
                                         device=dev, aggregation=net.aggregation_type)
            named = dict(self.named_parameters())
            bufs = dict(self.named_buffers())
            nb = net.num_conv_blocks
            self._param_list = [named[n] for n in runtime.canonical_param_names(nb, len(self.rnns))]
            if [id(p) for p in self._param_list] != [id(p) for p in self.parameters()]:
                raise RuntimeError("canonical parameter order differs from named_parameters() order")
            prefixes = runtime.canonical_bn_prefixes(nb)
            self._bn_lists = ([bufs[p + ".running_mean"] for p in prefixes],
                              [bufs[p + ".running_var"] for p in prefixes],
                              [bufs[p + ".num_batches_tracked"] for p in prefixes])
        return self._plan

    def forward(self, signal):
        plan = self._get_plan()
        if signal.dim() == 3:
            signal = signal.squeeze(-1)
        if not signal.is_cuda:
            raise RuntimeError("forward: `signal` must be on the model's CUDA device (got CPU tensor)")
        params = self._param_list
        plan.set_pointers([p.data for p in params], *self._bn_lists)
        if self.training:
            self._dropout_calls += 1
            seed = (torch.initial_seed() * 1000003 + self._dropout_calls) & 0x7FFFFFFFFFFFFFFF
            if torch.is_grad_enabled():
                class_logits = runtime.net_apply(plan, signal, seed, params)
            else:
                class_logits = plan.forward(signal, True, seed)
        else:
            class_logits = plan.forward(signal, False, 0)
        return dict(class_logits=class_logits)

    def extract_features(self, signal):
        """The model's input features `(N, n_features, frames)` (log-mel / log-STFT) for `signal (N, T[, 1])`: the fused
        feature kernel alone (reference networks/classifiers.py:565-579).  They depend on the feature descriptor only,
        not on the weights, so fold models with the same descriptor can share them (`forward_features`)."""
        from fsb200.runtime import FeatureExtractor
        if signal.dim() == 3:
            signal = signal.squeeze(-1)
        data = self.config.data
        kind, n_fft, hop = data.features.split("_")[0], int(data.features.split("_")[1]), int(data.features.split("_")[2])
        if getattr(self, "_feature_extractor", None) is None:
            self._feature_extractor = FeatureExtractor(n_fft, hop, self._filterbank_np, device=self.device)
        return self._feature_extractor(signal, 2 if kind == "mel" else 1)

    def forward_features(self, features, n_samples):
        """Eval-mode logits from precomputed features of clips with `n_samples` samples (no autograd)."""
        plan = self._get_plan()
        plan.set_pointers([p.data for p in self._param_list], *self._bn_lists)
        with torch.no_grad():
            return dict(class_logits=plan.forward_features(features, int(n_samples), training=False))

    # ------------------------------------------------------------------------------------------
    def add_scalar_summaries(self, loss, metric, writer, global_step):
        writer.add_scalar("loss", loss, global_step)
        writer.add_scalar("metric", metric, global_step)

    def add_histogram_summaries(self, losses, writer, global_step):
        writer.add_histogram("losses", np.array(losses), global_step=global_step)

    def add_image_summaries(self, signal, global_step, writer, to_plot=8):
        """Reference :621-631 (image grid of the first raw signals).  A summary that the installed
        tensorboard/PIL cannot encode must not stop training, so failures are reported once and skipped."""
        import torchvision.utils
        if len(signal) > to_plot:
            signal = signal[:to_plot]
        try:
            image_grid = torchvision.utils.make_grid(signal.data.cpu().unsqueeze(1), normalize=True, scale_each=True)
            writer.add_image("signal", image_grid, global_step)
        except Exception as exc:       # noqa: BLE001
            if not getattr(self, "_image_summary_warned", False):
                print("image summary skipped: %s" % exc)
                self._image_summary_warned = True

    def _loss(self, class_logits, labels, average):
        return lsep_loss(class_logits, labels, average=average)

    def _sync_replicas(self):
        """Data parallel: every replica starts from rank 0's parameters AND BatchNorm buffers.  Only gradients are
        exchanged afterwards, so replicas that were built under different RNG states (or loaded different
        checkpoints) would otherwise diverge silently."""
        from fsb200 import dist as fdist
        if fdist.world()[1] == 1:
            return
        import torch.distributed as dist
        with torch.no_grad():
            for t in list(self.parameters()) + list(self.buffers()):
                dist.broadcast(t.data, src=0)

    def _sync_gradients(self):
        """Data parallel: ONE all-reduce (SUM) of the flat gradient, averaged inside the Adam kernel."""
        from fsb200 import dist as fdist
        if fdist.world()[1] == 1:
            return
        scale = fdist.allreduce_gradients([p.grad for p in self._param_list],
                                          flat=getattr(self._plan, "last_flat_grad", None))
        if hasattr(self.optimizer, "grad_scale"):
            self.optimizer.grad_scale = scale
        else:
            for p in self._param_list:
                p.grad.mul_(scale)

    def train_epoch(self, train_loader, epoch, log_interval, write_summary=True):
        """Reference :633-707.  Same per-batch order of operations (LR step, forward, LSEP/accum, backward,
        optimiser step on `batch_idx % accumulation_steps == 0`, sigmoid + lwlrap); the three device->host
        reads per step are issued asynchronously into pinned memory and consumed one step later so the GPU
        never waits on sklearn."""
        self.train()
        print("\n" + " " * 10 + "****** Epoch {epoch} ******\n".format(epoch=epoch))
        training_losses = []
        history = deque(maxlen=30)
        self.optimizer.zero_grad()
        pending = None

        def resolve(item, pb):
            ev, losses_h, loss_h, metric_h, batch_idx, step, first_signal = item
            ev.synchronize()
            training_losses.extend(losses_h.numpy().copy())
            metric = float(metric_h)
            history.append(metric)
            pb.update()
            pb.set_description("Loss: {:.4f}, Metric: {:.4f}".format(float(loss_h), np.mean(history)))
            if batch_idx % log_interval == 0:
                self.add_scalar_summaries(float(loss_h), metric, self.train_writer, step)
            if first_signal is not None:
                self.add_image_summaries(first_signal, step, self.train_writer)

        with tqdm(total=len(train_loader), ncols=80) as pb:
            for batch_idx, sample in enumerate(train_loader):
                self.global_step += 1
                make_step(self.scheduler, step=self.global_step)
                signal = sample["signal"].to(self.device, non_blocking=True)
                labels = sample["labels"].to(self.device, non_blocking=True).float()

                outputs = self(signal)
                class_logits = outputs["class_logits"]
                if not self.two_d:
                    class_logits = class_logits.squeeze()
                if self.two_d:
                    loss_vec = self._loss(class_logits, labels, average=False) / self.config.train.accumulation_steps
                    loss = loss_vec.mean()
                else:
                    loss_vec = None
                    loss = self._loss(class_logits, labels, average=True) / self.config.train.accumulation_steps
                loss.backward()

                if batch_idx % self.config.train.accumulation_steps == 0:
                    self._sync_gradients()
                    self.optimizer.step()
                    self.optimizer.zero_grad()

                with torch.no_grad():
                    # per-batch lwlrap of sigmoid(logits) (reference :687-690) computed on the device: only the
                    # per-sample losses, the loss and the metric (8 bytes) travel to the host, asynchronously
                    probs = torch.sigmoid(class_logits.detach()).reshape(labels.shape)
                    metric_d = self._lwlrap_meter().batch(labels, probs)
                    losses_d = loss_vec.detach() if loss_vec is not None else loss.detach().reshape(1)
                    losses_h = torch.empty(losses_d.shape, dtype=torch.float32, pin_memory=True)
                    loss_h = torch.empty((), dtype=torch.float32, pin_memory=True)
                    metric_h = torch.empty((), dtype=torch.float64, pin_memory=True)
                    losses_h.copy_(losses_d, non_blocking=True)
                    loss_h.copy_(loss.detach(), non_blocking=True)
                    metric_h.copy_(metric_d, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                item = (ev, losses_h, loss_h, metric_h, batch_idx, self.global_step,
                        signal if batch_idx == 0 else None)
                if pending is not None:
                    resolve(pending, pb)
                pending = item
            if pending is not None:
                resolve(pending, pb)
        if self.two_d:
            self.add_histogram_summaries(training_losses, self.train_writer, self.global_step)

    def _lwlrap_meter(self):
        if getattr(self, "_lwlrap_dev", None) is None:
            self._lwlrap_dev = runtime.DeviceLwlrap(self.device)
        return self._lwlrap_dev

    def evaluate(self, loader, verbose=False, write_summary=False, epoch=None):
        """Reference :709-763: eval-mode forward, `loss * len(batch) / len(dataset)` accumulated over the batches,
        whole-set lwlrap of sigmoid(logits).  Loss terms and the lwlrap numerator / weight are accumulated on the
        device; the host reads them once at the end (the reference syncs per batch and runs sklearn on the host)."""
        self.eval()
        meter = runtime.DeviceLwlrap(self.device)
        valid_loss_d = torch.zeros((), dtype=torch.float64, device=self.device)
        with torch.no_grad():
            for batch_idx, sample in enumerate(loader):
                signal = sample["signal"].to(self.device)
                labels = sample["labels"].to(self.device).float()
                class_logits = self(signal)["class_logits"]
                if not self.two_d:
                    class_logits = class_logits.squeeze()
                loss = self._loss(class_logits, labels, average=True)
                valid_loss_d += loss.double() * (len(labels) / len(loader.dataset))
                meter.update(labels, torch.sigmoid(class_logits).reshape(labels.shape))
            valid_loss = float(valid_loss_d)
            metric = meter.compute()
            if write_summary:
                self.add_scalar_summaries(valid_loss, metric, writer=self.valid_writer, global_step=self.global_step)
            if verbose:
                print("\nValidation loss: {:.4f}".format(valid_loss))
                print("Validation metric: {:.4f}".format(metric))
            return metric

    def validation(self, valid_loader, epoch):
        return self.evaluate(valid_loader, verbose=True, write_summary=True, epoch=epoch)

    def predict(self, loader, n_tta=1):
        """Reference :770-797: sigmoid probabilities, mean over `n_tta` passes, numpy (n, C)."""
        self.eval()
        all_class_probs = []
        for k in range(n_tta):
            tta_probs = []
            with torch.no_grad():
                for sample in loader:
                    signal = sample["signal"].to(self.device)
                    class_logits = self(signal)["class_logits"]
                    tta_probs.append(torch.sigmoid(class_logits))
            all_class_probs.append(torch.cat(tta_probs).cpu().numpy())
        return np.mean(all_class_probs, 0)

    def predict_clips(self, clips, buckets, max_batch_elems, padding_value=0.0, return_stats=False):
        """Length-bucketed prediction over a list of 1-D waveforms (numpy arrays, pinned CPU tensors or CUDA tensors):
        the reference's unused `BucketingSampler` rule (ops/padding.py:36-81) wired into the `predict` loop of
        predict_2d_cnn.py:89-118 -- clips are binned by length, packed into batches of at most `max_batch_elems`
        samples, zero-padded per batch like `make_collate_fn`, and the sigmoid probabilities come back in the
        original clip order.  Under `torchrun` the batches are sharded over the ranks (fsb200.inference)."""
        from fsb200.inference import predict_bucketed
        return predict_bucketed(self, clips, buckets, max_batch_elems, padding_value=padding_value,
                                return_stats=return_stats)

    def fit_validate(self, train_loader, valid_loader, epochs, fold, log_interval=25):
        """Reference :799-868."""
        self.experiment.register_directory("summaries")
        self.train_writer = _summary_writer(os.path.join(self.experiment.summaries, "fold_{}".format(fold), "train"))
        self.valid_writer = _summary_writer(os.path.join(self.experiment.summaries, "fold_{}".format(fold), "valid"))
        os.makedirs(os.path.join(self.experiment.checkpoints, "fold_{}".format(fold)), exist_ok=True)

        from fsb200 import dist as fdist
        self.global_step = 0
        self.make_optimizer(max_steps=len(train_loader) * epochs)
        scores = []
        best_score = 0
        for epoch in range(epochs):
            make_step(self.scheduler, epoch=epoch)
            if epoch == self.config.train.switch_off_augmentations_on:
                train_loader.dataset.transform.switch_off_augmentations()
            self.train_epoch(train_loader, epoch, log_interval, write_summary=True)
            validation_score = self.validation(valid_loader, epoch)
            scores.append(validation_score)
            # data parallel: rank 0 writes the checkpoints (its BatchNorm running statistics; every rank would
            # otherwise race on the same file with its own per-rank statistics)
            writer_rank = fdist.world()[0] == 0
            if epoch % self.config.train._save_every == 0 and writer_rank:
                print("\nSaving model on epoch", epoch)
                torch.save(self.state_dict(), os.path.join(
                    self.experiment.checkpoints, "fold_{}".format(fold), "model_on_epoch_{}.pth".format(epoch)))
            if validation_score > best_score:
                if writer_rank:
                    torch.save(self.state_dict(), os.path.join(
                        self.experiment.checkpoints, "fold_{}".format(fold), "best_model.pth"))
                best_score = validation_score
        return scores

    def make_optimizer(self, max_steps):
        """Reference :870-880."""
        self._sync_replicas()
        optimizer = OPTIMIZERS[self.config.train.optimizer]
        optimizer = optimizer(self.parameters(), self.config.train.learning_rate,
                              weight_decay=self.config.train.weight_decay)
        self.optimizer = optimizer
        self.scheduler = make_scheduler(self.config.train.scheduler, max_steps=max_steps)(optimizer)

    def load_best_model(self, fold):
        """Reference :882-892."""
        self.load_state_dict(torch.load(
            os.path.join(self.experiment.checkpoints, "fold_{}".format(fold), "best_model.pth"),
            map_location=self.device))


class HierarchicalCNNClassificationModel(_AcceleratedCNN):
    """1D CNN over raw STFT bins (reference :107-480)."""
    two_d = False


class TwoDimensionalCNNClassificationModel(_AcceleratedCNN):
    """Frequency-encoded 2D CNN over log-mel features (reference :483-892)."""
    two_d = True
