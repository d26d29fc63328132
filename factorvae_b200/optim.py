"""Fused optimizer over the flat parameter buffer.

The reference steps `torch.optim.Adam(factorVAE.parameters(), lr)` and a per-batch `CosineAnnealingLR(T_max)`
(main.py:60-61, train_model.py:30-32): 28 + 5K small tensors -> as many tiny kernels per step.  Here parameters,
gradients and both moments are single flat fp32 buffers (the C ABI's layout) and a step is ONE kernel
(`fvae_adam_step`), issued right behind the gradient all-reduce."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _cabi
from .engine import _on, _stream


def cosine_annealing_lr(base_lr: float, t: int, T_max: int, eta_min: float = 0.0) -> float:
    """Closed form of torch.optim.lr_scheduler.CosineAnnealingLR after t scheduler steps (main.py:61)."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * t / T_max)) / 2.0


class FlatAdam:
    """torch.optim.Adam semantics (betas, eps, weight_decay as in torch; amsgrad off) on one flat buffer."""

    def __init__(self, flat_params: torch.Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, T_max: Optional[int] = None, eta_min: float = 0.0):
        if not flat_params.is_cuda or flat_params.dtype != torch.float32 or not flat_params.is_contiguous():
            raise RuntimeError("FlatAdam needs the contiguous fp32 CUDA parameter buffer (factorvae_b200 has no CPU path)")
        self.params = flat_params
        self._ptr = flat_params.data_ptr()
        self.exp_avg = torch.zeros_like(flat_params)
        self.exp_avg_sq = torch.zeros_like(flat_params)
        self.base_lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.T_max, self.eta_min = T_max, float(eta_min)
        self.step_count = 0          # optimizer steps taken
        self.sched_count = 0         # scheduler steps taken (CosineAnnealingLR.step(), train_model.py:31-32)

    @property
    def lr(self) -> float:
        if self.T_max is None:
            return self.base_lr
        return cosine_annealing_lr(self.base_lr, self.sched_count, self.T_max, self.eta_min)

    def step(self, flat_grad: torch.Tensor, grad_scale: float = 1.0) -> None:
        """optimizer.step() followed by scheduler.step() (the reference's order, train_model.py:30-32)."""
        if flat_grad.shape != self.params.shape or flat_grad.dtype != torch.float32 or not flat_grad.is_cuda:
            raise ValueError("flat_grad must match the flat fp32 CUDA parameter buffer")
        if self.params.data_ptr() != self._ptr:
            raise RuntimeError("the flat parameter buffer was reallocated (model.to() / a dtype change after the optimizer was "
                               "built): rebuild FlatAdam on model.flat_parameters() or call rebind()")
        self.step_count += 1
        b1, b2 = self.betas
        dev = self.params.device
        with _on(dev):
            rc = _cabi.lib().fvae_adam_step(self.params.data_ptr(), flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), self.params.numel(), self.lr, b1, b2, self.eps,
                                            self.weight_decay, self.step_count, float(grad_scale), _stream(dev))
        _cabi.check(rc, "fvae_adam_step")
        if self.T_max is not None:
            self.sched_count += 1

    def state_dict(self):
        return dict(exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, step=self.step_count, sched=self.sched_count,
                    lr=self.base_lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, T_max=self.T_max,
                    eta_min=self.eta_min)

    def rebind(self, flat_params: torch.Tensor) -> None:
        """Follow a reallocated parameter buffer (same length): the moments are kept."""
        if flat_params.shape != self.params.shape or flat_params.dtype != torch.float32 or not flat_params.is_cuda:
            raise ValueError("rebind needs a flat fp32 CUDA buffer of the same length")
        self.params, self._ptr = flat_params, flat_params.data_ptr()
        if self.exp_avg.device != flat_params.device:
            self.exp_avg, self.exp_avg_sq = self.exp_avg.to(flat_params.device), self.exp_avg_sq.to(flat_params.device)

    def load_state_dict(self, sd) -> None:
        """Restores the moments, both counters AND every hyper-parameter state_dict() saved, so a resumed run continues on
        the saved schedule (torch.optim.Optimizer.load_state_dict restores param_groups the same way)."""
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count, self.sched_count = int(sd["step"]), int(sd["sched"])
        if "lr" in sd:
            self.base_lr = float(sd["lr"])
        if "betas" in sd:
            self.betas = (float(sd["betas"][0]), float(sd["betas"][1]))
        if "eps" in sd:
            self.eps = float(sd["eps"])
        if "weight_decay" in sd:
            self.weight_decay = float(sd["weight_decay"])
        if "T_max" in sd:
            self.T_max = None if sd["T_max"] is None else int(sd["T_max"])
        if "eta_min" in sd:
            self.eta_min = float(sd["eta_min"])
