"""On-device batch assembly over a resident PCM pool (BASELINE.json configs[3]).

The reference assembles a training batch on the CPU, per sample in DataLoader workers:
`SampleLongAudio` crop (ops/transforms.py:292-309) -> `MixUp` with a random clean partner that went through the same crop
(ops/transforms.py:44-65, datasets/sound_dataset.py:41-56, ops/audio.py:32-52) -> zero-pad collate to the longest clip
(ops/padding.py:8-32), then ships ~113 MB per 64-clip batch over PCIe.  Here the decoded clips live in HBM once; per
batch the host only draws the random decisions -- in the reference's order and from the same RNG streams (numpy's global
generator for coin / alpha / crop start, `random` for partner index / mix offset), so a seeded run picks the same crops,
partners and offsets as the reference pipeline -- and one kernel (`fsb_assemble_batch`) does all PCM traffic.
"""
import random as _py_random

import numpy as np
import torch

from ._lib import check, lib
from .runtime import _ptr, _stream

ROW_DTYPE = np.dtype({"names": ["a_off", "b_off", "alpha", "one_minus", "a_len", "b_len", "a_label", "b_label",
                                "mix_offset", "pad"],
                      "formats": ["<i8", "<i8", "<f8", "<f8", "<i4", "<i4", "<i4", "<i4", "<i4", "<i4"],
                      "offsets": [0, 8, 16, 24, 32, 36, 40, 44, 48, 52], "itemsize": 56})


class DevicePcmPool:
    """Decoded clips (1-D float32 arrays of any lengths) and their multi-hot labels, resident on the device."""

    def __init__(self, clips, labels, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePcmPool lives in GPU memory; device=%r has no implementation" % (device,))
        self.lengths = np.asarray([len(c) for c in clips], dtype=np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)
        flat = np.concatenate([np.asarray(c, dtype=np.float32).reshape(-1) for c in clips])
        self.pcm = torch.from_numpy(flat).to(self.device)
        self.labels = torch.from_numpy(np.asarray(labels, dtype=np.float32)).to(self.device).contiguous()
        if self.labels.shape[0] != len(clips):
            raise ValueError("one label row per clip expected")

    def __len__(self):
        return len(self.lengths)


class DeviceBatchAssembler:
    """`assemble(indices)` -> (`signal` (N, T_max, 1), `labels` (N, C)) CUDA tensors, equal (bit for bit) to what the
    reference's transform chain + collate produce for the same clips under the same RNG state."""

    def __init__(self, pool, p_mixup=0.0, max_length=None, sr=44100, padding_value=0.0, np_rng=None, py_rng=None):
        self.pool = pool
        self.p_mixup = float(p_mixup)
        self.max_length = max_length            # seconds, like SampleLongAudio
        self.sr = int(sr)
        self.padding_value = float(padding_value)
        self.np_rng = np_rng if np_rng is not None else np.random      # np.random.RandomState works too
        self.py_rng = py_rng if py_rng is not None else _py_random     # random.Random works too

    def switch_off_augmentations(self):
        """What `Compose.switch_off_augmentations` does to MixUp (reference ops/transforms.py:343-358)."""
        self.p_mixup = 0.0

    def _crop(self, index):
        """SampleLongAudio: (first pool sample, length) of clip `index` after the crop."""
        off, size = int(self.pool.offsets[index]), int(self.pool.lengths[index])
        if self.max_length is not None and size / self.sr > self.max_length:
            window = int(self.max_length * self.sr)
            start = int(self.np_rng.randint(0, size - window))
            return off + start, window
        return off, size

    def draw(self, indices):
        """Host-side random decisions for one batch, as the 56-byte records the kernel consumes."""
        rows = np.zeros(len(indices), dtype=ROW_DTYPE)
        for r, index in enumerate(indices):
            index = int(index)
            a_off, a_len = self._crop(index)
            rows[r]["a_off"], rows[r]["a_len"], rows[r]["a_label"] = a_off, a_len, index
            rows[r]["b_len"] = -1
            if self.np_rng.uniform() < self.p_mixup:
                partner = self.py_rng.randint(0, len(self.pool) - 1)          # dataset.random_clean_sample()
                b_off, b_len = self._crop(partner)                             # its clean_transform crops too
                alpha = self.np_rng.uniform(0.4, 0.6)                          # drawn even for equal lengths
                rows[r]["b_off"], rows[r]["b_len"], rows[r]["b_label"] = b_off, b_len, partner
                if a_len != b_len:
                    n_long, n_short = max(a_len, b_len), min(a_len, b_len)
                    rows[r]["mix_offset"] = self.py_rng.randint(0, n_long - 1 - n_short)
                    rows[r]["alpha"] = alpha
                    rows[r]["one_minus"] = 1 - alpha
        return rows

    def assemble(self, indices, rows=None):
        rows = self.draw(indices) if rows is None else rows
        n = len(rows)
        mixed = rows["b_len"] >= 0
        t_out = int(np.where(mixed, np.maximum(rows["a_len"], rows["b_len"]), rows["a_len"]).max())
        dev = self.pool.device
        rows_dev = torch.from_numpy(rows.view(np.uint8).reshape(n, ROW_DTYPE.itemsize)).to(dev, non_blocking=True)
        signal = torch.empty((n, t_out, 1), dtype=torch.float32, device=dev)
        labels = torch.empty((n, self.pool.labels.shape[1]), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().fsb_assemble_batch(_ptr(self.pool.pcm), _ptr(self.pool.labels), _ptr(rows_dev), n,
                                           self.pool.labels.shape[1], t_out, self.padding_value, _ptr(signal),
                                           _ptr(labels), _stream()), "assemble_batch")
        return signal, labels
