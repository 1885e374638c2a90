"""Losses with the reference's signatures (networks/losses.py of the reference).  Only `lsep_loss`
is on the training path; it runs as a warp-per-sample CUDA kernel (forward and backward)."""
import torch


def _lsep(input, target, average, stable):
    from fsb200.runtime import lsep_per_sample
    if not input.is_cuda:
        raise RuntimeError("lsep_loss: CUDA tensors only (this package has no CPU path)")
    squeeze = input.dim() == 1
    if squeeze:                         # the 1D model squeezes a batch of one (reference :268)
        input, target = input[None], target[None]
    lsep = lsep_per_sample(input, target.to(input.device), stable)
    if average:
        return lsep.mean()
    return lsep[0] if squeeze else lsep


def lsep_loss(input, target, average=True):
    """`log(1 + sum_{i,j: t_j < t_i} exp(s_j - s_i))` per sample (reference :47-58); mean over the
    batch when `average`, else the per-sample vector."""
    return _lsep(input, target, average, False)


def lsep_loss_stable(input, target, average=True):
    """The reference's max-shifted variant (:25-44): `m + log(exp(-m) + sum exp(d - m))` with
    `m = max_{i,j}(s_j - s_i)` -- equal to `lsep_loss` where that does not overflow, finite beyond."""
    return _lsep(input, target, average, True)


def binary_cross_entropy(input, target, raw=True):
    if raw:
        input = torch.sigmoid(input)
    return torch.nn.functional.binary_cross_entropy(input, target)


def focal_loss(input, target, focus=2.0, raw=True):
    if raw:
        input = torch.sigmoid(input)
    eps = 1e-7
    prob_true = torch.clamp(input * target + (1 - input) * (1 - target), eps, 1 - eps)
    return (-(1.0 - prob_true).pow(focus) * prob_true.log()).mean()
