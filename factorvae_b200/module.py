"""Drop-in replacement for the reference's `module.py` (reference module.py:10-278).

Same class names, constructor signatures, attribute names, `state_dict` keys, `forward` /
`prediction` return shapes -- so the reference's `main.py`, `train_model.py` and `utils.py` run
unchanged (`from module import FactorVAE, FeatureExtractor, ...`, main.py:12).  The arithmetic of
one ELBO step -- `FactorVAE.forward` (module.py:250-270) and its backward (train_model.py:29) --
runs in the sm_100a CUDA library behind include/fvae_b200.h; PyTorch only owns memory, streams and
the autograd edge.  There is no CPU or eager fallback: CPU tensors raise.

Parameters are ordinary fp32 `nn.Parameter`s created by the same torch layers in the same order as
the reference, so `torch.manual_seed(s)` gives bit-identical initial weights, and reference
checkpoints load with `load_state_dict`.  At the first forward (and after any `.to()`), FactorVAE
re-homes every parameter as a view of one flat buffer in the library's layout; optimizers keep
working because they hold the Parameter objects.

The six sub-modules (FactorEncoder, AlphaLayer, BetaLayer, FactorDecoder, AttentionLayer, FactorPredictor) answer when called on
their own as well -- one launch of the heads kernel on the caller's stock latents (fvae_heads_parts) -- forward only: the reference
never differentiates them outside FactorVAE.forward, whose backward is the fused one.

Noise: the reference draws eps with `randn_like` (module.py:104, in eval too) and dropout masks
with `nn.Dropout(0.1)` (module.py:132,144).  Here both come from an in-kernel Philox stream keyed
by (torch.initial_seed(), step counter, stock, head); `inject_noise` supplies explicit tensors for
parity tests.
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch
import torch.nn as nn

from . import engine

__all__ = ["FeatureExtractor", "FactorEncoder", "AlphaLayer", "BetaLayer", "FactorDecoder", "AttentionLayer",
           "FactorPredictor", "FactorVAE", "inject_noise", "set_default_precision"]

# Default 'fp32': an unmodified reference main.py then trains with reference-grade numerics (1e-5 parity with the CPU
# path).  The tcgen05 bf16 mode (stated tolerances: ELBO 2e-2, gradients rel-L2 3e-2) is opt-in: FVAE_PRECISION=bf16|auto in
# the environment or set_default_precision(); the date-batched engine (batched.DateShardedStep, bench.py) asks for it itself.
_DEFAULT_PRECISION = os.environ.get("FVAE_PRECISION", "fp32")
_INJECTED = None


def set_default_precision(p: str) -> None:
    """'fp32' (CUDA-core, 1e-5 parity; the default), 'bf16' (tcgen05 tensor cores) or 'auto' (bf16 when supported)."""
    global _DEFAULT_PRECISION
    if p not in ("fp32", "bf16", "auto"):
        raise ValueError(p)
    _DEFAULT_PRECISION = p


@contextlib.contextmanager
def inject_noise(eps: torch.Tensor, keep_mask: Optional[torch.Tensor] = None):
    """Use explicit eps (N,) and dropout keep-mask (N, K) instead of the Philox stream (tests)."""
    global _INJECTED
    prev, _INJECTED = _INJECTED, (eps, keep_mask)
    try:
        yield
    finally:
        _INJECTED = prev


def _cuda_only(t: torch.Tensor, who: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{who}: input is on {t.device}; factorvae_b200 runs on CUDA (sm_100a) only and has no "
                           "CPU fallback. Move the model and the batch to a CUDA device.")


class _FeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        layout = engine.ParamLayout(mod.num_latent, mod.hidden_size, 1, 1)
        flat = torch.zeros(layout.total, dtype=torch.float32, device=x.device)
        for name, p in zip(_FE_NAMES, params):
            layout.view(flat, "feature_extractor." + name).copy_(p.detach())
        e, saved = engine.fe_forward(layout, flat, x, mod._precision())
        ctx.layout, ctx.saved = layout, saved
        return e

    @staticmethod
    def backward(ctx, de):
        grad = engine.fe_backward(ctx.layout, ctx.saved, de)
        return (None, None) + tuple(ctx.layout.view(grad, "feature_extractor." + n) for n in _FE_NAMES)


_FE_NAMES = ["normalize.weight", "normalize.bias", "linear.weight", "linear.bias", "gru.weight_ih_l0",
             "gru.weight_hh_l0", "gru.bias_ih_l0", "gru.bias_hh_l0"]


def _resolve_precision(p: str, C: int, H: int) -> str:
    if p != "auto":
        return p
    return "bf16" if engine.tc_supported(C, H) else "fp32"


class FeatureExtractor(nn.Module):
    """LayerNorm -> Linear -> LeakyReLU -> GRU -> last hidden state (reference module.py:10-31)."""

    def __init__(self, num_latent, hidden_size, num_layers=1):
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("only num_layers=1 is supported (every reference call site uses the default)")
        self.num_latent, self.hidden_size, self.num_layers = num_latent, hidden_size, num_layers
        self.normalize = nn.LayerNorm(num_latent)
        self.linear = nn.Linear(num_latent, num_latent)
        self.leakyrelu = nn.LeakyReLU()
        self.gru = nn.GRU(num_latent, hidden_size, num_layers, batch_first=True)

    def _precision(self):
        return _resolve_precision(_DEFAULT_PRECISION, self.num_latent, self.hidden_size)

    def forward(self, x):
        _cuda_only(x, "FeatureExtractor.forward")
        params = [self.normalize.weight, self.normalize.bias, self.linear.weight, self.linear.bias,
                  self.gru.weight_ih_l0, self.gru.weight_hh_l0, self.gru.bias_ih_l0, self.gru.bias_hh_l0]
        return _FeFn.apply(self, x, *params)


_PART_STEP = [0]
_WARNED = set()


def _parts(who: str, named: dict, H: int, K: int, M: int, stock_latent: torch.Tensor, *, train: bool = False, **kw):
    """One stand-alone sub-module call (include/fvae_b200.h: fvae_heads_parts).  `named` maps reference state_dict names of
    the full model to this module's parameters; the sections of the other modules stay zero.  Forward only: the outputs
    carry no autograd edge (the gradient of the heads exists inside FactorVAE.forward's backward)."""
    import warnings
    _cuda_only(stock_latent, who + ".forward")
    if torch.is_grad_enabled() and who not in _WARNED and (stock_latent.requires_grad or any(p.requires_grad for p in named.values())):
        _WARNED.add(who)
        warnings.warn(f"{who}.forward on its own is forward-only on factorvae_b200: its outputs are not attached to autograd. "
                      "Train through FactorVAE.forward (the reference's only call site); use torch.no_grad() to silence this.")
    layout = engine.ParamLayout(1, H, K, M)
    flat = torch.zeros(layout.total, dtype=torch.float32, device=stock_latent.device)
    for name, p in named.items():
        layout.view(flat, name).copy_(p.detach())
    N = stock_latent.shape[0]
    if _INJECTED is not None:
        eps, km = _INJECTED
        if eps is None:
            eps = torch.zeros(N, device=stock_latent.device)
        kw.update(eps=eps, keep_mask=km if train else None)
    else:
        _PART_STEP[0] += 1
        kw.update(philox=(torch.initial_seed() ^ 0x5DEECE66D, _PART_STEP[0], 0))
    return engine.heads_parts(layout, flat, stock_latent, train=train, **kw)


def _sub(prefix: str, mod: nn.Module) -> dict:
    return {prefix + n: p for n, p in mod.named_parameters()}


class FactorEncoder(nn.Module):
    """Portfolio layer + mapping layer (reference module.py:33-67): parameter holder of the fused step."""

    def __init__(self, num_factors, num_portfolio, hidden_size):
        super().__init__()
        self.num_factors = num_factors
        self.linear = nn.Linear(hidden_size, num_portfolio)
        self.softmax = nn.Softmax(dim=0)
        self.linear_mu = nn.Linear(num_portfolio, num_factors)
        self.linear_sigma = nn.Linear(num_portfolio, num_factors)
        self.softplus = nn.Softplus()

    def forward(self, stock_latent, returns):
        """(N, H), (N, 1) | (N,) -> (factor_mu (K,), factor_sigma (K,)), reference module.py:52-67.  A sigma that underflows to
        exactly 0 comes back as 1e-6 (the :117 clamp the decoder would apply is fused into the kernel)."""
        _cuda_only(returns, "FactorEncoder.forward")
        o = _parts("FactorEncoder", _sub("factor_encoder.", self), self.linear.in_features, self.linear_mu.out_features,
                   self.linear.out_features, stock_latent, y=returns)
        return o["mu_post"], o["sigma_post"]


class AlphaLayer(nn.Module):
    """Idiosyncratic-return head (reference module.py:69-84): parameter holder of the fused step."""

    def __init__(self, hidden_size):
        super().__init__()
        self.linear1 = nn.Linear(hidden_size, hidden_size)
        self.leakyrelu = nn.LeakyReLU()
        self.mu_layer = nn.Linear(hidden_size, 1)
        self.sigma_layer = nn.Linear(hidden_size, 1)
        self.softplus = nn.Softplus()

    def forward(self, stock_latent):
        """(N, H) -> (alpha_mu (N, 1), alpha_sigma (N, 1)), reference module.py:78-84."""
        o = _parts("AlphaLayer", _sub("factor_decoder.alpha_layer.", self), self.linear1.in_features, 1, 1, stock_latent,
                   want=("alpha",))
        return o["alpha_mu"], o["alpha_sigma"]


class BetaLayer(nn.Module):
    """Factor exposure beta (N, K) (reference module.py:86-94): parameter holder of the fused step."""

    def __init__(self, hidden_size, num_factors):
        super().__init__()
        self.linear1 = nn.Linear(hidden_size, num_factors)

    def forward(self, stock_latent):
        """(N, H) -> beta (N, K), reference module.py:92-94."""
        o = _parts("BetaLayer", _sub("factor_decoder.beta_layer.", self), self.linear1.in_features, self.linear1.out_features, 1,
                   stock_latent, want=("beta",))
        return o["beta"]


class FactorDecoder(nn.Module):
    """alpha + beta . z with reparameterisation (reference module.py:96-123)."""

    def __init__(self, alpha_layer, beta_layer):
        super().__init__()
        self.alpha_layer = alpha_layer
        self.beta_layer = beta_layer

    def reparameterize(self, mu, sigma):
        """mu + eps * sigma (reference module.py:103-105); inside forward the draw happens in the kernel."""
        return mu + torch.randn_like(sigma) * sigma

    def forward(self, stock_latent, factor_mu, factor_sigma):
        """(N, H), (K,), (K,) -> sampled returns (N, 1), reference module.py:107-123; zeros of factor_sigma are replaced by
        1e-6 in the caller's tensor too, as :117 does through its view."""
        beta = self.beta_layer.linear1
        with torch.no_grad():
            factor_sigma.masked_fill_(factor_sigma == 0, 1e-6)
        o = _parts("FactorDecoder", _sub("factor_decoder.", self), beta.in_features, beta.out_features, 1, stock_latent,
                   z=(factor_mu, factor_sigma))
        return o["yhat"].reshape(-1, 1)


class AttentionLayer(nn.Module):
    """One attention head of the prior (reference module.py:125-153).  The K heads are stacked and
    collapsed algebraically inside the kernels; this class keeps the per-head parameter names."""

    def __init__(self, hidden_size):
        super().__init__()
        self.query = nn.Parameter(torch.randn(hidden_size))
        self.key_layer = nn.Linear(hidden_size, hidden_size)
        self.value_layer = nn.Linear(hidden_size, hidden_size)
        self.dropout = nn.Dropout(0.1)

    def forward(self, stock_latent):
        """(N, H) -> context vector (H,), zeros if the attention weights hold NaN / Inf (reference module.py:134-153).
        Dropout on the scores follows self.training."""
        H = self.key_layer.in_features
        o = _parts("AttentionLayer", _sub("factor_predictor.attention_layers.0.", self), H, 1, 1, stock_latent,
                   train=self.training, want=("context",))
        return o["context"].reshape(H)


class FactorPredictor(nn.Module):
    """K attention heads + shared MLP head -> prior (mu, sigma) (reference module.py:155-188)."""

    def __init__(self, hidden_size, num_factor):
        super().__init__()
        self.hidden_size, self.num_factor = hidden_size, num_factor
        self.attention_layers = nn.ModuleList([AttentionLayer(hidden_size) for _ in range(num_factor)])
        self.linear = nn.Linear(hidden_size, hidden_size)
        self.leakyrelu = nn.LeakyReLU()
        self.mu_layer = nn.Linear(hidden_size, 1)
        self.sigma_layer = nn.Linear(hidden_size, 1)
        self.softplus = nn.Softplus()

    def forward(self, stock_latent):
        """(N, H) -> (pred_mu (K,), pred_sigma (K,)), reference module.py:169-188.  A sigma that underflows to exactly 0 comes
        back as 1e-6 (the :264-265 clamp of FactorVAE.forward is fused into the kernel)."""
        o = _parts("FactorPredictor", _sub("factor_predictor.", self), self.hidden_size, self.num_factor, 1, stock_latent,
                   train=self.training)
        return o["mu_prior"], o["sigma_prior"]


class _ElboFn(torch.autograd.Function):
    """FactorVAE.forward as one autograd node: forward = fvae_elbo_forward, backward = fvae_elbo_backward."""

    @staticmethod
    def forward(ctx, model, x, returns, noise_kw, *params):
        N = x.shape[0]
        date_ptr = engine.single_date_ptr(N, x.device)
        out, st = engine.elbo_forward(model._layout, model._flat, x, returns.reshape(-1), date_ptr,
                                      train=model.training, precision=model.precision, **noise_kw)
        ctx.model, ctx.st = model, st
        loss = out["loss"].reshape(())
        rest = (out["yhat"].reshape(-1, 1), out["mu_post"].reshape(-1), out["sigma_post"].reshape(-1),
                out["mu_prior"].reshape(-1), out["sigma_prior"].reshape(-1))
        ctx.mark_non_differentiable(*rest)
        return (loss,) + rest

    @staticmethod
    def backward(ctx, gloss, *unused):
        model = ctx.model
        grad = engine.elbo_backward(model._layout, ctx.st)
        grad = grad * gloss
        return (None, None, None, None) + tuple(model._layout.view(grad, n) for n in model._param_names)


class FactorVAE(nn.Module):
    """The wrapper whose forward/prediction are the drop-in boundary (reference module.py:234-278)."""

    def __init__(self, feature_extractor, factor_encoder, factor_decoder, factor_predictor):
        super().__init__()
        self.feature_extractor = feature_extractor
        self.factor_encoder = factor_encoder
        self.factor_decoder = factor_decoder
        self.factor_predictor = factor_predictor
        self._flat = None
        self._layout = None
        self._param_names = None
        self._step = 0
        self._last = None          # StepState of the latest forward (mu_y / sigma_y live there)

    # ---- dimensions ---------------------------------------------------------------------------
    def dims(self):
        fe, en = self.feature_extractor, self.factor_encoder
        return dict(C=fe.num_latent, H=fe.hidden_size, K=en.linear_mu.out_features, M=en.linear.out_features)

    @property
    def precision(self) -> str:
        d = self.dims()
        return _resolve_precision(_DEFAULT_PRECISION, d["C"], d["H"])

    @staticmethod
    def KL_Divergence(mu1, sigma1, mu2, sigma2):
        return (torch.log(sigma2 / sigma1) + (sigma1 ** 2 + (mu1 - mu2) ** 2) / (2 * sigma2 ** 2) - 0.5).sum()

    # ---- flat parameter storage ---------------------------------------------------------------
    def flat_parameters(self) -> torch.Tensor:
        """Make every parameter a view of one flat fp32 CUDA buffer (library layout) and return it."""
        d = self.dims()
        if self._layout is None or (self._layout.C, self._layout.H, self._layout.K, self._layout.M) != tuple(d.values()):
            self._layout = engine.ParamLayout(d["C"], d["H"], d["K"], d["M"])
            self._param_names = [n for n, _ in self.named_parameters()]
            missing = set(self._param_names) ^ set(self._layout.slices)
            if missing:
                raise RuntimeError(f"parameter inventory differs from the reference layout: {sorted(missing)[:4]}...")
            self._flat = None
        params = dict(self.named_parameters())
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("FactorVAE parameters are on the CPU: call .to('cuda') first; there is no CPU path.")
        flat = self._flat
        ok = flat is not None and flat.device == dev
        if ok:
            base = flat.data_ptr()
            for name, (off, _) in self._layout.slices.items():
                p = params[name]
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
        if not ok:
            flat = torch.zeros(self._layout.total, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for name in self._layout.slices:
                    v = self._layout.view(flat, name)
                    v.copy_(params[name].detach().to(torch.float32))
                    params[name].data = v
            self._flat = flat
        return flat

    def _noise_kwargs(self, N, K, device):
        if _INJECTED is not None:
            eps, km = _INJECTED
            return dict(eps=eps, keep_mask=km)
        self._step += 1
        return dict(philox=(torch.initial_seed(), self._step, 0))

    # ---- the boundary -------------------------------------------------------------------------
    def forward(self, x, returns):
        """x (N, T, C), returns (N, 1) or (N,) -> (vae_loss, reconstruction (N,1), factor_mu (K,),
        factor_sigma (K,), pred_mu (K,), pred_sigma (K,)) exactly as reference module.py:270."""
        _cuda_only(x, "FactorVAE.forward")
        self.flat_parameters()
        params = [p for _, p in self.named_parameters()]
        nk = self._noise_kwargs(x.shape[0], self.dims()["K"], x.device)
        res = _ElboFn.apply(self, x, returns, nk, *params)
        return res

    def prediction(self, x):
        """Prior factors through the decoder -> sampled y_pred (N, 1) (reference module.py:273-278)."""
        _cuda_only(x, "FactorVAE.prediction")
        flat = self.flat_parameters()
        N = x.shape[0]
        nk = self._noise_kwargs(N, self.dims()["K"], x.device)
        nk.pop("keep_mask", None)
        date_ptr = engine.single_date_ptr(N, x.device)
        with torch.no_grad():
            out, st = engine.elbo_forward(self._layout, flat, x, None, date_ptr, train=False, precision=self.precision,
                                          predict=True, **nk)
        self._last = st
        return out["yhat"].reshape(-1, 1)

    predict = prediction     # alias named by the task statement; the reference method is `prediction`
