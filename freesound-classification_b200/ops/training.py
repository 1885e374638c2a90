"""Optimiser / LR-schedule helpers with the reference's names (ops/training.py of the reference).

`OPTIMIZERS["adam"]` is Adam(amsgrad=True) executed as ONE multi-tensor CUDA kernel
(`fsb_adam_amsgrad_step`) instead of torch's per-tensor loop; state tensors, `param_groups`,
`step()`, `zero_grad()` and `state_dict()` behave like `torch.optim.Adam` so schedulers and
checkpoints keep working.
"""
import ctypes
from functools import partial

import numpy as np
import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import StepLR


class FusedAdam(Optimizer):
    """Adam with amsgrad, torch 2.x update rule, coupled L2 `weight_decay`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=True):
        if not amsgrad:
            raise ValueError("FusedAdam implements the amsgrad variant only (reference ops/training.py:10)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=True))
        self._table = None
        self._table_key = None
        self.grad_scale = 1.0          # set to 1/world_size after a SUM all-reduce

    def _build_table(self, group, plist):
        from fsb200._lib import lib
        chunk = lib().fsb_adam_chunk()
        # merge tensors that are adjacent in memory (flat parameter / gradient buffers) into one record
        recs = []
        for p in plist:
            st = self.state[p]
            rec = [p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                   st["max_exp_avg_sq"].data_ptr(), p.numel()]
            if recs:
                last = recs[-1]
                nbytes = last[5] * 4
                if all(last[i] + nbytes == rec[i] for i in range(5)):
                    last[5] += rec[5]
                    continue
            recs.append(rec)
        table = np.zeros((len(recs), 6), dtype=np.int64)
        blocks = []
        for i, r in enumerate(recs):
            table[i] = r
            for c in range((r[5] + chunk - 1) // chunk):
                blocks.append((i, c))
        dev = plist[0].device
        return (torch.from_numpy(table).to(dev), torch.tensor(blocks, dtype=torch.int32, device=dev).contiguous(),
                len(blocks))

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._table = self._table_key = None       # the moments were replaced: rebuild the device pointer table

    @torch.no_grad()
    def step(self, closure=None):
        from fsb200._lib import check, lib
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam: parameters must live on a CUDA device (no CPU path)")
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam: contiguous float32 parameters and gradients required")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    st["max_exp_avg_sq"] = torch.zeros_like(p)
                else:
                    for name in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):     # loaded state: kernel layout
                        t = st[name]
                        if t.device != p.device or t.dtype != torch.float32 or not t.is_contiguous():
                            st[name] = t.to(device=p.device, dtype=torch.float32).contiguous()
            # the device pointer table bakes in parameter, gradient AND state addresses: all of them key the cache
            # (load_state_dict / state resets replace the moment tensors without touching the parameters)
            key = (gi, tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                              self.state[p]["exp_avg_sq"].data_ptr(), self.state[p]["max_exp_avg_sq"].data_ptr())
                             for p in plist))
            if key != self._table_key:
                self._table = self._build_table(group, plist)
                self._table_key = key
            table, block_map, n_blocks = self._table
            step = int(self.state[plist[0]]["step"]) + 1       # torch.optim.Adam checkpoints carry a tensor `step`
            for p in plist:
                self.state[p]["step"] = step
            beta1, beta2 = group["betas"]
            with torch.cuda.device(plist[0].device):
                stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                check(lib().fsb_adam_amsgrad_step(
                    ctypes.c_void_p(table.data_ptr()), ctypes.c_void_p(block_map.data_ptr()), n_blocks, step,
                    float(group["lr"]), beta1, beta2, group["eps"], group["weight_decay"], float(self.grad_scale),
                    stream), "adam_amsgrad_step")
        return loss


OPTIMIZERS = {
    "adam": partial(FusedAdam, amsgrad=True),
    "momentum": partial(torch.optim.SGD, momentum=0.9, nesterov=True)
}


def annealing_linear(start, end, r):
    return start + r * (end - start)


def annealing_cos(start, end, r):
    cos_out = np.cos(np.pi * r) + 1
    return end + (start - end) / 2 * cos_out


class OneCycleScheduler:
    """Linear warm-up `min_lr -> max_lr` over the first round(0.3 * max_steps) steps, then linear decay
    to `min_lr / 1000`; overwrites `param_group["lr"]` on every call (reference :208-234)."""

    def __init__(self, optimizer, min_lr, max_lr, max_steps, annealing=annealing_linear):
        self.optimizer = optimizer
        self.min_lr = min_lr
        self.max_lr = max_lr
        self.max_steps = max_steps
        self.annealing = annealing
        self.epoch = -1

    def step(self):
        self.epoch += 1
        mid = int(round(self.max_steps * 0.3))
        if self.epoch < mid:
            lr = self.annealing(self.min_lr, self.max_lr, self.epoch / mid)
        else:
            lr = self.annealing(self.max_lr, self.min_lr / 1e3, (self.epoch - mid) / (self.max_steps - mid))
        for param_group in self.optimizer.param_groups:
            param_group["lr"] = lr


def make_scheduler(params, max_steps):
    """`"steplr_<step>_<gamma>"` or `"1cycle_<min_lr>_<max_lr>"` -> scheduler factory (reference :15-34)."""
    name, *args = params.split("_")
    if name == "steplr":
        step_size, gamma = args
        return partial(StepLR, step_size=int(step_size), gamma=float(gamma))
    elif name == "1cycle":
        min_lr, max_lr = args
        return partial(OneCycleScheduler, min_lr=float(min_lr), max_lr=float(max_lr), max_steps=max_steps)


def make_step(scheduler, epoch=None, step=None, val_score=None):
    if isinstance(scheduler, StepLR) and epoch is not None:
        scheduler.step(epoch)
    elif isinstance(scheduler, OneCycleScheduler) and step is not None:
        scheduler.step()
