"""Fused Adam + per-batch cosine schedule against torch.optim.Adam / CosineAnnealingLR, the
optimizer and scheduler the reference constructs (main.py:60-61) and steps per batch (train_model.py:30-32)."""
import math

import pytest
import torch


def test_cosine_schedule_matches_torch_scheduler():
    from factorvae_b200.optim import cosine_annealing_lr
    p = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.Adam([p], lr=3e-4)
    T_max = 37
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=T_max)
    for t in range(2 * T_max + 3):
        assert math.isclose(opt.param_groups[0]["lr"], cosine_annealing_lr(3e-4, t, T_max), rel_tol=1e-6, abs_tol=1e-12), t
        opt.step()
        sch.step()


@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(wd, cuda_device):
    from factorvae_b200.optim import FlatAdam
    torch.manual_seed(0)
    n = 62630 + 3                                   # not a multiple of 4: exercises the tail
    p0 = torch.randn(n)
    ref_p = torch.nn.Parameter(p0.clone().double())  # fp64 torch Adam as the yardstick for both
    ref32 = torch.nn.Parameter(p0.clone())
    opt64 = torch.optim.Adam([ref_p], lr=1e-3, weight_decay=wd)
    opt32 = torch.optim.Adam([ref32], lr=1e-3, weight_decay=wd)
    T_max = 50
    s64 = torch.optim.lr_scheduler.CosineAnnealingLR(opt64, T_max=T_max)
    s32 = torch.optim.lr_scheduler.CosineAnnealingLR(opt32, T_max=T_max)
    flat = torch.zeros(n + 1, device=cuda_device)[:n]
    flat.copy_(p0)
    mine = FlatAdam(flat, lr=1e-3, weight_decay=wd, T_max=T_max)
    for step in range(25):
        g = torch.randn(n) * (0.1 + 0.05 * step)
        ref_p.grad = g.double(); ref32.grad = g.clone()
        opt64.step(); s64.step(); opt32.step(); s32.step()
        mine.step(g.to(cuda_device))
    got = flat.cpu().double()
    err_mine = float((got - ref_p.detach()).abs().max())
    err_torch32 = float((ref32.detach().double() - ref_p.detach()).abs().max())
    assert err_mine <= max(2.0 * err_torch32, 1e-6), (err_mine, err_torch32)      # as close to exact Adam as torch's own fp32
    assert float((got - ref32.detach().double()).abs().max()) <= 2e-6
    assert float((mine.exp_avg.cpu() - opt32.state[ref32]["exp_avg"]).abs().max()) <= 1e-6
    assert float((mine.exp_avg_sq.cpu() - opt32.state[ref32]["exp_avg_sq"]).abs().max()) <= 1e-6
