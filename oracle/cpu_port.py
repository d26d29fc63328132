"""CPU port of the reference's per-date training step -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.

Purpose: the `cpu_baseline` and `--impl reference` legs of bench.py.  /root/reference is Python and
does not travel to the GPU box, so the reference's CPU path is re-expressed here with the SAME
library calls in the SAME structure (fp32 ATen on the host cores), so that its timing is
representative of the reference:
  * F.layer_norm -> F.linear -> F.leaky_relu -> fused ATen GRU           (module.py:26-31)
  * softmax over stocks, mm, two linears + softplus                       (module.py:56-64, :48-49)
  * a PYTHON LOOP over the K attention heads, each with its own key/value projection, dropout on
    the scores, relu, softmax, the `.any()` guard (a host sync), and torch.cat  (module.py:134-153, :172-177)
  * alpha/beta heads, in-place sigma clamp, sqrt, randn_like               (module.py:80-84, :93, :109-123)
  * F.mse_loss + KL, the `torch.any` guard                                 (module.py:242-248, :261-268)
  * one date per step: zero_grad -> forward -> loss.item() -> backward     (train_model.py:26-29)
It is pinned against tests/golden (tests/test_oracle_golden.py::test_cpu_port_*), with eps and
dropout masks injected.  Only tests/, __graft_entry__.smoke() and bench.py may import this file.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


class CpuPort:
    def __init__(self, params: Dict[str, torch.Tensor], dtype=torch.float32):
        self.p = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in params.items()}
        self.H = self.p["feature_extractor.gru.weight_hh_l0"].shape[1]
        self.K = self.p["factor_encoder.linear_mu.weight"].shape[0]
        self.C = self.p["feature_extractor.normalize.weight"].numel()
        self._gru_w = [self.p["feature_extractor.gru." + n] for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]

    def zero_grad(self):
        for v in self.p.values():
            v.grad = None

    def feature_extractor(self, x):
        p = self.p
        xn = F.layer_norm(x, (self.C,), p["feature_extractor.normalize.weight"], p["feature_extractor.normalize.bias"])
        u = F.leaky_relu(F.linear(xn, p["feature_extractor.linear.weight"], p["feature_extractor.linear.bias"]))
        h0 = x.new_zeros(1, x.shape[0], self.H)
        out, _ = torch._VF.gru(u, h0, self._gru_w, True, 1, 0.0, False, False, True)
        return out[:, -1, :]

    def attention_head(self, k, e, train, keep):
        p = self.p
        a = f"factor_predictor.attention_layers.{k}."
        key = F.linear(e, p[a + "key_layer.weight"], p[a + "key_layer.bias"])
        value = F.linear(e, p[a + "value_layer.weight"], p[a + "value_layer.bias"])
        s = torch.matmul(p[a + "query"], key.transpose(1, 0))
        s = s / torch.sqrt(torch.tensor(key.shape[1]) + 1e-6)
        if keep is not None:
            s = s * keep / 0.9
        elif train:
            s = F.dropout(s, 0.1, True)
        w = F.softmax(F.relu(s), dim=0)
        if torch.isnan(w).any() or torch.isinf(w).any():
            return torch.zeros_like(value[0])
        return torch.matmul(w, value)

    def step_forward(self, x, y, train=True, eps=None, keep_mask=None):
        p = self.p
        e = self.feature_extractor(x)
        w = F.softmax(F.linear(e, p["factor_encoder.linear.weight"], p["factor_encoder.linear.bias"]), dim=0)
        yp = torch.mm(w.transpose(1, 0), y.reshape(-1, 1)).squeeze(1)
        mu_post = F.linear(yp, p["factor_encoder.linear_mu.weight"], p["factor_encoder.linear_mu.bias"])
        sg_post = F.softplus(F.linear(yp, p["factor_encoder.linear_sigma.weight"], p["factor_encoder.linear_sigma.bias"]))
        # decoder with the posterior
        ha = F.leaky_relu(F.linear(e, p["factor_decoder.alpha_layer.linear1.weight"], p["factor_decoder.alpha_layer.linear1.bias"]))
        a_mu = F.linear(ha, p["factor_decoder.alpha_layer.mu_layer.weight"], p["factor_decoder.alpha_layer.mu_layer.bias"])
        a_sg = F.softplus(F.linear(ha, p["factor_decoder.alpha_layer.sigma_layer.weight"], p["factor_decoder.alpha_layer.sigma_layer.bias"]))
        beta = F.linear(e, p["factor_decoder.beta_layer.linear1.weight"], p["factor_decoder.beta_layer.linear1.bias"])
        fmu, fsg = mu_post.view(-1, 1), sg_post.view(-1, 1)
        fsg[fsg == 0] = 1e-6
        mu_y = a_mu + torch.matmul(beta, fmu)
        sg_y = torch.sqrt(a_sg ** 2 + torch.matmul(beta ** 2, fsg ** 2) + 1e-6)
        noise = torch.randn_like(sg_y) if eps is None else eps.reshape(-1, 1)
        yhat = mu_y + noise * sg_y
        # prior: K heads, one after the other
        heads = None
        for k in range(self.K):
            c = self.attention_head(k, e, train, None if keep_mask is None else keep_mask[k])
            heads = c if heads is None else torch.cat((heads, c), dim=0)
        hm = F.leaky_relu(F.linear(heads.view(self.K, -1), p["factor_predictor.linear.weight"], p["factor_predictor.linear.bias"]))
        mu_prior = F.linear(hm, p["factor_predictor.mu_layer.weight"], p["factor_predictor.mu_layer.bias"]).view(-1)
        sg_prior = F.softplus(F.linear(hm, p["factor_predictor.sigma_layer.weight"], p["factor_predictor.sigma_layer.bias"])).view(-1)
        rec = F.mse_loss(yhat, y.reshape(-1, 1))
        if torch.any(sg_prior == 0):
            sg_prior[sg_prior == 0] = 1e-6
        kl = (torch.log(sg_prior / sg_post) + (sg_post ** 2 + (mu_post - mu_prior) ** 2) / (2 * sg_prior ** 2) - 0.5).sum()
        return rec + kl, yhat, mu_y, sg_y

    def train_step(self, x, y, **kw) -> float:
        """zero_grad -> forward -> loss.item() -> backward, as train_model.py:26-29 (no optimizer)."""
        self.zero_grad()
        loss, *_ = self.step_forward(x, y, train=True, **kw)
        v = loss.item()
        loss.backward()
        return v


def pick_threads(port: "CpuPort", x, y) -> int:
    """Intra-op thread count at which the reference-style step is fastest on this host.

    "All the host threads it can use": these are small ATen ops, and on a many-core host the intra-op
    pool at full width is far SLOWER than a few threads (6.4 s/date at 128 threads vs ~25 ms at 8 on the
    first B200 box).  Probing a few widths keeps the baseline the reference at its best, not a strawman."""
    import os
    ncpu = os.cpu_count() or 1
    best = (float("inf"), 1)
    for cand in sorted({1, 4, 8, 16, 32, min(64, ncpu), ncpu}):
        if cand > ncpu:
            continue
        torch.set_num_threads(cand)
        port.train_step(x, y)
        t0 = time.perf_counter()
        port.train_step(x, y)
        dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, cand)
        if dt > 2.0:          # already hopeless at this width; wider will not help
            break
    torch.set_num_threads(best[1])
    return best[1]


def time_cpu_steps(params: Dict[str, torch.Tensor], N: int, T: int, C: int, *, budget_s: float = 10.0, warmup: int = 2,
                   min_steps: int = 3, threads: Optional[int] = None, seed: int = 0):
    """Per-date reference-style steps on the host cores; returns (units_per_s, ms_per_date, steps, threads)."""
    import os
    port = CpuPort(params)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, T, C, generator=g).clamp_(-3, 3)
    y = torch.randn(N, 1, generator=g)
    if threads is None:
        threads = pick_threads(port, x, y)
    nthreads = threads
    torch.set_num_threads(nthreads)
    for _ in range(warmup):
        port.train_step(x, y)
    times: List[float] = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_steps or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        port.train_step(x, y)
        times.append(time.perf_counter() - t0)
        if len(times) >= 2000:
            break
    times.sort()
    med = times[len(times) // 2]
    return N / med, med * 1e3, len(times), nthreads
