"""Batch collation + length bucketing (host side; reference ops/padding.py)."""
import random

import numpy as np
from torch.utils.data.dataloader import default_collate


def make_collate_fn(padding_values):
    """Right-pad every field named in `padding_values` to the batch maximum (constant or "edge"),
    then `default_collate` (reference :8-32).  Defines the (N, T_max, 1) layout K-feat ingests."""

    def _collate_fn(batch):
        for name, padding_value in padding_values.items():
            longest = max(len(sample[name]) for sample in batch)
            for sample in batch:
                missing = longest - len(sample[name])
                if not missing:
                    continue
                widths = [(0, missing)] + [(0, 0)] * (sample[name].ndim - 1)
                if padding_value == "edge":
                    sample[name] = np.pad(sample[name], widths, mode="edge")
                else:
                    sample[name] = np.pad(sample[name], widths, mode="constant", constant_values=padding_value)
        return default_collate(batch)

    return _collate_fn


class BucketingSampler:
    """Length-bucketed batch sampler (reference :36-81): `np.digitize(lengths, buckets)`, per bin a
    shuffled greedy fill until the summed length reaches `max_batch_elems`, then shuffled batches.
    Bins 0 (below buckets[0]) and len(buckets) (at/above buckets[-1]) are dropped, as in the reference."""

    def __init__(self, dataset, max_batch_elems, buckets):
        self.buckets = buckets
        self.dataset = dataset
        self.max_batch_elems = max_batch_elems
        self._create_batches()

    def _create_batches(self):
        self.n_bins = len(self.buckets)
        lengths = self.dataset.lengths
        binned = np.digitize(lengths, self.buckets)
        batches = []
        for bin_idx in range(1, self.n_bins):
            ids = np.nonzero(binned == bin_idx)[0]
            random.shuffle(ids)
            filled, batch = 0, []
            for i in ids:
                if filled < self.max_batch_elems:
                    batch.append(i)
                    filled += lengths[i]
                else:
                    batches.append(batch)
                    filled, batch = lengths[i], [i]
            if batch:
                batches.append(batch)
        random.shuffle(batches)
        self.n_batches = len(batches)
        self.batches = batches

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return self.n_batches
