"""Length-bucketed inference over variable-length clips (BASELINE.json configs[4]).

The reference predicts in dataframe order with every batch zero-padded to its longest clip
(predict_2d_cnn.py:89-118, ops/padding.py:8-32) and ships an unused `BucketingSampler` (ops/padding.py:36-81); its
README says the final submission grouped clips by length.  This module wires that sampler's semantics into a
prediction loop: clips are binned by length, packed into batches of at most `max_batch_elems` samples, zero-padded
per batch exactly like `make_collate_fn`, run through the model in eval mode, and the sigmoid probabilities are
scattered back to the original clip order.  Batches are independent, so with `torch.distributed` initialised every rank
takes a contiguous share of the batches and the results are summed with one all-reduce.
"""
import ctypes

import numpy as np
import torch

from . import dist as fdist


def pack_batches(lengths, buckets, max_batch_elems):
    """Deterministic variant of the reference sampler's packing (no shuffling: prediction order does not matter).
    Returns (batches, dropped): lists of clip indices; clips shorter than buckets[0] or at/after buckets[-1] fall in no
    bin -- the reference silently drops them (ops/padding.py:53), here they are reported."""
    lengths = np.asarray(lengths)
    bin_of = np.digitize(lengths, buckets)
    batches = []
    for b in range(1, len(buckets)):
        members = np.nonzero(bin_of == b)[0]
        members = members[np.argsort(lengths[members], kind="stable")]     # neighbours in length share a batch
        current, load = [], 0
        for idx in members:
            if load >= max_batch_elems:
                batches.append(current)
                current, load = [], 0
            current.append(int(idx))
            load += int(lengths[idx])
        if current:
            batches.append(current)
    dropped = [int(i) for i in np.nonzero((bin_of == 0) | (bin_of == len(buckets)))[0]]
    return batches, dropped


def pad_batch(clips, indices, padding_value=0.0):
    """(len(indices), T_max, 1) float32, right-padded with `padding_value` like `make_collate_fn`."""
    t_max = max(len(clips[i]) for i in indices)
    out = np.full((len(indices), t_max, 1), padding_value, dtype=np.float32)
    for row, i in enumerate(indices):
        out[row, :len(clips[i]), 0] = np.asarray(clips[i], dtype=np.float32).reshape(-1)
    return out


def device_batch(clips, indices, device, padding_value=0.0):
    """The zero-padded batch `(len(indices), T_max, 1)` of `make_collate_fn`, built directly in device memory.

    * clips that already live on the device as views of ONE buffer (a resident PCM pool) are gathered and padded by
      the batch-assembly kernel (`fsb_assemble_batch`, one launch);
    * host clips (numpy arrays or -- preferably pinned -- CPU tensors) are copied row by row with asynchronous
      host-to-device copies into a pre-filled batch: no padded copy is ever materialised on the host."""
    device = torch.device(device)
    srcs = [clips[i] if isinstance(clips[i], torch.Tensor)
            else torch.from_numpy(np.ascontiguousarray(clips[i], dtype=np.float32)) for i in indices]
    srcs = [c.reshape(-1) for c in srcs]
    n, t_max = len(srcs), max(int(c.numel()) for c in srcs)
    pooled = all(c.is_cuda and c.dtype == torch.float32 and c.is_contiguous() for c in srcs)
    if pooled:
        storages = {c.untyped_storage().data_ptr() for c in srcs}
        pooled = len(storages) == 1
    if pooled:
        from ._lib import check, lib
        from .assemble import ROW_DTYPE
        from .runtime import _ptr, _stream
        base = srcs[0].untyped_storage().data_ptr()
        rows = np.zeros(n, dtype=ROW_DTYPE)
        rows["a_off"] = [(c.data_ptr() - base) // 4 for c in srcs]
        rows["a_len"] = [int(c.numel()) for c in srcs]
        rows["b_len"] = -1
        rows_dev = torch.from_numpy(rows.view(np.uint8).reshape(n, ROW_DTYPE.itemsize)).to(device, non_blocking=True)
        out = torch.empty((n, t_max, 1), dtype=torch.float32, device=device)
        dummy = torch.zeros((n + 1, 1), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            check(lib().fsb_assemble_batch(ctypes.c_void_p(base), _ptr(dummy), _ptr(rows_dev), n, 1, t_max,
                                           float(padding_value), _ptr(out), _ptr(dummy[1:]), _stream()), "assemble_batch")
        return out
    out = torch.full((n, t_max, 1), float(padding_value), dtype=torch.float32, device=device)
    for row, c in enumerate(srcs):
        out[row, :c.numel(), 0].copy_(c, non_blocking=True)
    return out


def predict_bucketed(model, clips, buckets, max_batch_elems, padding_value=0.0, return_stats=False):
    """Sigmoid class probabilities (n_clips, n_classes) float32 for a list of 1-D waveforms.

    model: an `_AcceleratedCNN` (eval mode is set here); buckets / max_batch_elems: as in `BucketingSampler`.
    clips: numpy arrays, CPU tensors (pinned memory makes the copies asynchronous) or CUDA tensors.
    Clips that fall outside the buckets get NaN rows (the reference would silently skip them).
    With an initialised process group the batches are sharded over the ranks and the result is all-reduced."""
    lengths = [int(c.numel()) if isinstance(c, torch.Tensor) else len(c) for c in clips]
    batches, dropped = pack_batches(lengths, buckets, max_batch_elems)
    rank, world_size = fdist.world()
    begin, end = fdist.shard_range(len(batches), rank, world_size)
    n_classes = model.config.data._n_classes
    device = torch.device(model.device)
    probs = torch.zeros((len(clips), n_classes), dtype=torch.float32, device=device)
    # padded-sample overhead of the WHOLE sweep (all ranks' batches)
    padded = sum(len(b) * max(lengths[i] for i in b) for b in batches)
    real = sum(lengths[i] for b in batches for i in b)
    model.eval()
    with torch.no_grad():
        for indices in batches[begin:end]:
            batch = device_batch(clips, indices, device, padding_value)
            logits = model(batch)["class_logits"]
            probs[torch.as_tensor(indices, device=device)] = torch.sigmoid(logits)
    if world_size > 1:
        torch.distributed.all_reduce(probs)
    out = probs.cpu().numpy()
    if dropped:
        out[dropped] = np.nan
    if return_stats:
        return out, dict(batches=len(batches), dropped=len(dropped), padded_samples=padded, real_samples=real,
                         padding_overhead=(padded / real - 1.0) if real else 0.0)
    return out


def predict_folds(models, clips, buckets, max_batch_elems, padding_value=0.0):
    """Fold ensemble (predict_2d_cnn.py:111-118 loads one model per fold and averages their probabilities): every padded
    batch is assembled ONCE, its features are extracted ONCE (they do not depend on the weights) and each fold model runs
    on the shared features.  Returns the mean sigmoid probability `(n_clips, n_classes)`; clips outside the buckets get
    NaN rows.  All models must use the same feature descriptor."""
    descriptors = {m.config.data.features for m in models}
    if len(descriptors) != 1:
        raise ValueError("fold models must share one feature descriptor, got %r" % (sorted(descriptors),))
    lengths = [int(c.numel()) if isinstance(c, torch.Tensor) else len(c) for c in clips]
    batches, dropped = pack_batches(lengths, buckets, max_batch_elems)
    rank, world_size = fdist.world()
    begin, end = fdist.shard_range(len(batches), rank, world_size)
    first = models[0]
    device = torch.device(first.device)
    probs = torch.zeros((len(clips), first.config.data._n_classes), dtype=torch.float32, device=device)
    for m in models:
        m.eval()
    with torch.no_grad():
        for indices in batches[begin:end]:
            batch = device_batch(clips, indices, device, padding_value)
            feats = first.extract_features(batch)
            acc = None
            for m in models:
                p = torch.sigmoid(m.forward_features(feats, batch.shape[1])["class_logits"])
                acc = p if acc is None else acc + p
            probs[torch.as_tensor(indices, device=device)] = acc / len(models)
    if world_size > 1:
        torch.distributed.all_reduce(probs)
    out = probs.cpu().numpy()
    if dropped:
        out[dropped] = np.nan
    return out
