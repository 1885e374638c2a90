"""Losses with the reference's signatures (networks/losses.py of the reference).  Only `lsep_loss`
is on the training path; it runs as a warp-per-sample CUDA kernel (forward and backward)."""
import torch


def lsep_loss(input, target, average=True):
    """`log(1 + sum_{i,j: t_j < t_i} exp(s_j - s_i))` per sample (reference :47-58); mean over the
    batch when `average`, else the per-sample vector."""
    from fsb200.runtime import lsep_per_sample
    if not input.is_cuda:
        raise RuntimeError("lsep_loss: CUDA tensors only (this package has no CPU path)")
    squeeze = input.dim() == 1
    if squeeze:                         # the 1D model squeezes a batch of one (reference :268)
        input, target = input[None], target[None]
    lsep = lsep_per_sample(input, target.to(input.device))
    if average:
        return lsep.mean()
    return lsep[0] if squeeze else lsep


def lsep_loss_stable(input, target, average=True):
    """Max-shifted variant kept for API parity (reference :25-44); mathematically identical to
    `lsep_loss` wherever the latter does not overflow, so it shares the kernel."""
    return lsep_loss(input, target, average)


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
