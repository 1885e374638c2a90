"""Batch collation + length bucketing (host side; mirrors the semantics of the reference's ops/padding.py)."""
import random

import numpy as np
from torch.utils.data.dataloader import default_collate


def _pad_tail(array, target_len, fill):
    """Right-pad axis 0 of `array` to `target_len` with a constant, or by repeating the last element ("edge")."""
    short_by = target_len - len(array)
    if short_by == 0:
        return array
    spec = [(0, short_by)] + [(0, 0)] * (array.ndim - 1)
    if fill == "edge":
        return np.pad(array, spec, mode="edge")
    return np.pad(array, spec, mode="constant", constant_values=fill)


def make_collate_fn(padding_values):
    """collate_fn that right-pads every field listed in `padding_values` ({field: constant | "edge"}) to the longest
    sample of the batch and then applies torch's `default_collate` (reference :8-32).  For the signal this produces the
    zero-padded (N, T_max, 1) float32 tensor the feature kernel ingests."""

    def collate(batch):
        for field, fill in padding_values.items():
            target = max(len(item[field]) for item in batch)
            for item in batch:
                item[field] = _pad_tail(item[field], target, fill)
        return default_collate(batch)

    return collate


class BucketingSampler:
    """Length-bucketed batch sampler with the reference's semantics (:36-81): samples are binned with
    `np.digitize(dataset.lengths, buckets)`; inside each bin the shuffled ids are packed greedily -- a batch is
    closed by the first sample that arrives after its summed length reached `max_batch_elems` -- and the list of
    batches is shuffled.  Bin 0 (shorter than buckets[0]) and bin len(buckets) (at or beyond buckets[-1]) are dropped,
    as in the reference."""

    def __init__(self, dataset, max_batch_elems, buckets):
        self.dataset, self.max_batch_elems, self.buckets = dataset, max_batch_elems, buckets
        self.n_bins = len(buckets)
        self.batches = self._pack()
        self.n_batches = len(self.batches)

    def _pack(self):
        lengths = self.dataset.lengths
        bin_of = np.digitize(lengths, self.buckets)
        out = []
        for b in range(1, self.n_bins):
            members = np.nonzero(bin_of == b)[0]
            random.shuffle(members)
            current, load = [], 0
            for idx in members:
                if load >= self.max_batch_elems:
                    out.append(current)
                    current, load = [], 0
                current.append(idx)
                load += lengths[idx]
            if current:
                out.append(current)
        random.shuffle(out)
        return out

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return self.n_batches
