"""Length-bucketed inference over variable-length clips (BASELINE.json configs[4]).

The reference predicts in dataframe order with every batch zero-padded to its longest clip
(predict_2d_cnn.py:89-118, ops/padding.py:8-32) and ships an unused `BucketingSampler` (ops/padding.py:36-81); its
README says the final submission grouped clips by length.  This module wires that sampler's semantics into a
prediction loop: clips are binned by length, packed into batches of at most `max_batch_elems` samples, zero-padded
per batch exactly like `make_collate_fn`, run through the model in eval mode, and the sigmoid probabilities are
scattered back to the original clip order.  Batches are independent, so with `torch.distributed` initialised every rank
takes a contiguous share of the batches and the results are summed with one all-reduce.
"""
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


def predict_bucketed(model, clips, buckets, max_batch_elems, padding_value=0.0, return_stats=False):
    """Sigmoid class probabilities (n_clips, n_classes) float32 for a list of 1-D waveforms.

    model: an `_AcceleratedCNN` (eval mode is set here); buckets / max_batch_elems: as in `BucketingSampler`.
    Clips that fall outside the buckets get NaN rows (the reference would silently skip them).
    With an initialised process group the batches are sharded over the ranks and the result is all-reduced."""
    lengths = [len(c) for c in clips]
    batches, dropped = pack_batches(lengths, buckets, max_batch_elems)
    rank, world_size = fdist.world()
    begin, end = fdist.shard_range(len(batches), rank, world_size)
    n_classes = model.config.data._n_classes
    device = torch.device(model.device)
    probs = torch.zeros((len(clips), n_classes), dtype=torch.float32, device=device)
    padded = real = 0
    model.eval()
    with torch.no_grad():
        for indices in batches[begin:end]:
            batch = pad_batch(clips, indices, padding_value)
            padded += batch.shape[0] * batch.shape[1]
            real += sum(lengths[i] for i in indices)
            logits = model(torch.from_numpy(batch).to(device, non_blocking=True))["class_logits"]
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
