"""Evaluation metric of the reference on the device: RankIC / RankIC_IR (utils.py:113-129)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _cabi
from .engine import _on, _stream


def rank_ic(pred: torch.Tensor, label: torch.Tensor, date_ptr: torch.Tensor) -> Tuple[torch.Tensor, float, float]:
    """Per-date Spearman correlation (average ranks for ties) of pred and label, both (S,), grouped by the CSR date_ptr
    (B+1,).  Returns (ric per date (B,), RankIC = mean, RankIC_IR = mean / std) exactly as utils.RankIC aggregates them
    (np.mean / np.std with ddof 0 over the dates; IR is NaN when std == 0)."""
    if not pred.is_cuda or not label.is_cuda:
        raise RuntimeError("rank_ic runs on a CUDA device: factorvae_b200 has no CPU path")
    pred = pred.reshape(-1).to(torch.float32).contiguous()
    label = label.reshape(-1).to(device=pred.device, dtype=torch.float32).contiguous()
    date_ptr = date_ptr.to(device=pred.device, dtype=torch.int32).contiguous()
    B = date_ptr.numel() - 1
    counts = (date_ptr[1:] - date_ptr[:-1])
    nmax = int(counts.max().item()) if B > 0 else 0
    ric = torch.empty(B, dtype=torch.float32, device=pred.device)
    with _on(pred.device):
        rc = _cabi.lib().fvae_rank_ic(pred.data_ptr(), label.data_ptr(), date_ptr.data_ptr(), B, max(nmax, 1), ric.data_ptr(),
                                      _stream(pred.device))
    _cabi.check(rc, "fvae_rank_ic")
    r64 = ric.double()
    mean = float(r64.mean().item()) if B else float("nan")
    std = float(r64.std(unbiased=False).item()) if B else float("nan")
    return ric, mean, (mean / std if std != 0 else float("nan"))
