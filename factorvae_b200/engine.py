"""Host side of the C ABI: parameter layout, buffer ownership (torch) and the ELBO-step calls.

PyTorch is plumbing here: it owns device memory and streams; every number is produced by
libfvae_b200.so.  The reference calls being replaced are `FactorVAE.forward` + `loss.backward()`
(reference module.py:250-270, train_model.py:27-29) and `FactorVAE.prediction` (module.py:273-278).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _cabi

PRECISIONS = {"fp32": _cabi.PREC_FP32, "bf16": _cabi.PREC_BF16_TC}


class ParamLayout:
    """The flat fp32 parameter buffer of include/fvae_b200.h, keyed by reference state_dict names."""

    def __init__(self, C_: int, H: int, K: int, M: int):
        self.C, self.H, self.K, self.M = C_, H, K, M
        offs = _cabi.param_offsets(C_, H, K, M)
        self.offsets = dict(zip(_cabi.SECTIONS, offs[:-1]))
        self.total = offs[-1]
        o = self.offsets
        s: "OrderedDict[str, Tuple[int, Tuple[int, ...]]]" = OrderedDict()
        fe, en, de, pr = "feature_extractor.", "factor_encoder.", "factor_decoder.", "factor_predictor."
        s[fe + "normalize.weight"] = (o["LN_W"], (C_,))
        s[fe + "normalize.bias"] = (o["LN_B"], (C_,))
        s[fe + "linear.weight"] = (o["W1"], (C_, C_))
        s[fe + "linear.bias"] = (o["B1"], (C_,))
        s[fe + "gru.weight_ih_l0"] = (o["WIH"], (3 * H, C_))
        s[fe + "gru.weight_hh_l0"] = (o["WHH"], (3 * H, H))
        s[fe + "gru.bias_ih_l0"] = (o["BIH"], (3 * H,))
        s[fe + "gru.bias_hh_l0"] = (o["BHH"], (3 * H,))
        s[en + "linear.weight"] = (o["ENC_W"], (M, H))
        s[en + "linear.bias"] = (o["ENC_B"], (M,))
        s[en + "linear_mu.weight"] = (o["ENC_MU_W"], (K, M))
        s[en + "linear_mu.bias"] = (o["ENC_MU_B"], (K,))
        s[en + "linear_sigma.weight"] = (o["ENC_SG_W"], (K, M))
        s[en + "linear_sigma.bias"] = (o["ENC_SG_B"], (K,))
        s[de + "alpha_layer.linear1.weight"] = (o["AL_W"], (H, H))
        s[de + "alpha_layer.linear1.bias"] = (o["AL_B"], (H,))
        s[de + "alpha_layer.mu_layer.weight"] = (o["AL_MU_W"], (1, H))
        s[de + "alpha_layer.mu_layer.bias"] = (o["AL_MU_B"], (1,))
        s[de + "alpha_layer.sigma_layer.weight"] = (o["AL_SG_W"], (1, H))
        s[de + "alpha_layer.sigma_layer.bias"] = (o["AL_SG_B"], (1,))
        s[de + "beta_layer.linear1.weight"] = (o["BETA_W"], (K, H))
        s[de + "beta_layer.linear1.bias"] = (o["BETA_B"], (K,))
        for k in range(K):
            a = f"{pr}attention_layers.{k}."
            s[a + "query"] = (o["ATT_Q"] + k * H, (H,))
            s[a + "key_layer.weight"] = (o["ATT_KW"] + k * H * H, (H, H))
            s[a + "key_layer.bias"] = (o["ATT_KB"] + k * H, (H,))
            s[a + "value_layer.weight"] = (o["ATT_VW"] + k * H * H, (H, H))
            s[a + "value_layer.bias"] = (o["ATT_VB"] + k * H, (H,))
        s[pr + "linear.weight"] = (o["PR_W"], (H, H))
        s[pr + "linear.bias"] = (o["PR_B"], (H,))
        s[pr + "mu_layer.weight"] = (o["PR_MU_W"], (1, H))
        s[pr + "mu_layer.bias"] = (o["PR_MU_B"], (1,))
        s[pr + "sigma_layer.weight"] = (o["PR_SG_W"], (1, H))
        s[pr + "sigma_layer.bias"] = (o["PR_SG_B"], (1,))
        self.slices = s

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self.slices[name]
        n = 1
        for d in shape:
            n *= d
        return flat[off:off + n].view(shape)

    def pack(self, state: Dict[str, torch.Tensor], device) -> torch.Tensor:
        flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        for name in self.slices:
            self.view(flat, name).copy_(state[name].to(device=device, dtype=torch.float32))
        return flat

    def unpack(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {name: self.view(flat, name) for name in self.slices}


@dataclass
class StepState:
    """Everything one forward leaves behind for backward (all device buffers owned by torch)."""
    shape: "_cabi.Shape"
    panel: "_cabi.Panel"
    noise: "_cabi.Noise"
    outs: "_cabi.Outputs"
    flags: int
    precision: int
    tensors: Dict[str, torch.Tensor]   # keeps every buffer alive
    workspace: torch.Tensor
    flat_version: Optional[int] = None   # flat._version at forward: backward refuses parameters modified in between


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: factorvae_b200 has no CPU path "
                           f"(got device {t.device}).")


class IndexedWindows:
    """The look-back windows in resident-panel form (include/fvae_b200.h, fvae_panel.row_index): `table` is the
    (date, instrument) row table (R, pitch >= C) fp32|bf16 on the device, `row_index` (S, T) int32 names the table row of
    every (sequence, time step).  Accepted wherever the engine takes the dense x (S, T, C); the kernels read the rows in
    place, the T-fold duplicated window tensor of the reference's DataLoader is never built."""

    def __init__(self, table: torch.Tensor, row_index: torch.Tensor, C_: int):
        if table.dim() != 2 or table.stride(1) != 1 or table.dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("table must be a (rows, pitch) fp32 / bf16 matrix with unit inner stride")
        if row_index.dim() != 2 or row_index.dtype != torch.int32 or not row_index.is_contiguous():
            raise ValueError("row_index must be a contiguous (S, T) int32 tensor")
        if C_ > table.shape[1]:
            raise ValueError("C exceeds the table width")
        self.table, self.row_index, self.C = table, row_index, int(C_)

    @property
    def shape(self):
        return (self.row_index.shape[0], self.row_index.shape[1], self.C)

    @property
    def is_cuda(self):
        return self.table.is_cuda and self.row_index.is_cuda

    @property
    def device(self):
        return self.table.device

    @property
    def dtype(self):
        return self.table.dtype


def _panel(x) -> "_cabi.Panel":
    if isinstance(x, IndexedWindows):
        return x, _cabi.Panel(x.table.data_ptr(), _cabi.F32 if x.table.dtype == torch.float32 else _cabi.BF16, 0,
                              x.table.stride(0), x.row_index.data_ptr(), x.table.shape[0])
    if x.dim() != 3:
        raise ValueError("x must be (S, T, C)")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    if x.stride(2) != 1:
        x = x.contiguous()
    return x, _cabi.Panel(x.data_ptr(), _cabi.F32 if x.dtype == torch.float32 else _cabi.BF16, x.stride(0), x.stride(1), None, 0)


def tc_supported(C_: int, H: int) -> bool:
    """True when the tcgen05 (bf16) FeatureExtractor path covers this (C, H)."""
    shape = _cabi.Shape(128, 1, 1, C_, H, 1, 1)
    return _cabi.lib().fvae_workspace_bytes(C.byref(shape), _cabi.PREC_BF16_TC) > 0


_DATE_PTR_CACHE: Dict[Tuple[int, str], torch.Tensor] = {}


def single_date_ptr(N: int, device) -> torch.Tensor:
    key = (N, str(device))
    t = _DATE_PTR_CACHE.get(key)
    if t is None:
        if len(_DATE_PTR_CACHE) > 4096:
            _DATE_PTR_CACHE.clear()
        t = torch.tensor([0, N], dtype=torch.int32, device=device)
        _DATE_PTR_CACHE[key] = t
    return t


def _stream(device) -> C.c_void_p:
    """The current stream OF THE DEVICE the buffers live on (not of the current device)."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _on(device):
    """Every ABI call runs with the buffers' device current: kernel attributes, launches and the stream all belong to it
    (a model on cuda:1 while cuda:0 is current must not launch on device 0 with device-1 pointers)."""
    return torch.cuda.device(device)


def uniform_date_ptr(B: int, N: int, device) -> torch.Tensor:
    return torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=device)


def _philox_noise(philox, dev, keep: Dict) -> "_cabi.Noise":
    """philox = (seed, step, unit_base).  `step` is a Python int, or a one-element int64 CUDA tensor: the kernels then read the
    counter from device memory at run time (fvae_noise.step_dev) -- what a step captured in a CUDA graph needs."""
    seed, step, base = philox
    if isinstance(step, torch.Tensor):
        if step.dtype != torch.int64 or step.numel() != 1 or step.device != dev:
            raise ValueError("a device step counter must be a one-element int64 tensor on the batch's device")
        keep["step_dev"] = step
        return _cabi.Noise(None, None, int(seed) & (2 ** 64 - 1), 0, int(base), step.data_ptr())
    return _cabi.Noise(None, None, int(seed) & (2 ** 64 - 1), int(step), int(base), None)


def elbo_forward(layout: ParamLayout, flat: torch.Tensor, x: torch.Tensor, y: Optional[torch.Tensor],
                 date_ptr: torch.Tensor, *, eps: Optional[torch.Tensor] = None,
                 keep_mask: Optional[torch.Tensor] = None, train: bool = True, precision: str = "fp32",
                 philox: Optional[Tuple[int, int, int]] = None, predict: bool = False,
                 workspace: Optional[torch.Tensor] = None,
                 loss_out: Optional[torch.Tensor] = None,
                 out_cache: Optional[Dict] = None) -> Tuple[Dict[str, torch.Tensor], StepState]:
    """One forward over B dates.  x (S,T,C) fp32|bf16, y (S,), date_ptr int32 (B+1,) CSR over dates.

    Noise: pass eps (S,) [and keep_mask (S,K) uint8 when train] for injected-noise parity runs, or
    philox=(seed, step, unit_base) for the in-kernel counter RNG (shard invariant).
    out_cache: a dict owned by the caller; the nine output tensors are then allocated once per (S, B) and REUSED by later calls
    (a training loop that consumes the outputs before the next step: saves nine allocations per step on the host)."""
    L = _cabi.lib()
    _require_cuda(flat, "parameters")
    _require_cuda(x, "x")
    dev = x.device
    x, panel = _panel(x)
    S, T, Cf = x.shape
    if Cf != layout.C:
        raise ValueError(f"x has {Cf} features, the model was built for {layout.C}")
    B = date_ptr.numel() - 1
    K, H, M = layout.K, layout.H, layout.M
    if date_ptr.dtype != torch.int32 or not date_ptr.is_cuda:
        date_ptr = date_ptr.to(device=dev, dtype=torch.int32)
    shape = _cabi.Shape(S, B, T, Cf, H, K, M)
    prec = PRECISIONS[precision]
    flags = 0
    if train and not predict:
        flags |= _cabi.FLAG_TRAIN
    keep: Dict[str, torch.Tensor] = dict(x=x, date_ptr=date_ptr, flat=flat)
    if philox is not None:
        flags |= _cabi.FLAG_PHILOX
        noise = _philox_noise(philox, dev, keep)
    else:
        if eps is None:
            raise ValueError("either eps (and keep_mask in train mode) or philox=(seed, step, unit_base) is required")
        eps = eps.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        keep["eps"] = eps
        km_ptr = None
        if flags & _cabi.FLAG_TRAIN:
            if keep_mask is None:
                raise ValueError("train mode needs keep_mask (S,K) or philox")
            keep_mask = keep_mask.to(device=dev, dtype=torch.uint8).contiguous()
            if keep_mask.shape != (S, K):
                raise ValueError("keep_mask must be (S, K)")
            keep["keep_mask"] = keep_mask
            km_ptr = keep_mask.data_ptr()
        noise = _cabi.Noise(eps.data_ptr(), km_ptr, 0, 0, 0)
    f32 = dict(dtype=torch.float32, device=dev)
    cached = out_cache.get((S, B, K, str(dev))) if out_cache is not None else None
    if cached is not None:
        out = dict(cached)
        if loss_out is not None:
            out["loss"] = loss_out
    else:
        out = dict(loss=loss_out if loss_out is not None else torch.empty(1, **f32), date_loss=torch.empty(B, **f32), yhat=torch.empty(S, **f32),
                   mu_y=torch.empty(S, **f32), sigma_y=torch.empty(S, **f32), mu_post=torch.empty(B, K, **f32),
                   sigma_post=torch.empty(B, K, **f32), mu_prior=torch.empty(B, K, **f32),
                   sigma_prior=torch.empty(B, K, **f32))
        if out_cache is not None:
            if len(out_cache) > 8:
                out_cache.clear()
            out_cache[(S, B, K, str(dev))] = dict(out)
    outs = _cabi.Outputs(*[out[n].data_ptr() for n, _ in _cabi.Outputs._fields_])
    need = L.fvae_workspace_bytes(C.byref(shape), prec)
    if need < 0:
        _cabi.check(int(need), "fvae_workspace_bytes")
    if workspace is None or workspace.numel() < need or workspace.device != dev:
        workspace = torch.empty(int(need), dtype=torch.uint8, device=dev)
    if flat.device != dev:
        raise RuntimeError(f"parameters live on {flat.device}, the batch on {dev}")
    if predict:
        with _on(dev):
            rc = L.fvae_predict(C.byref(shape), C.byref(panel), date_ptr.data_ptr(), flat.data_ptr(), C.byref(noise), flags,
                                prec, C.byref(outs), workspace.data_ptr(), workspace.numel(), _stream(dev))
        _cabi.check(rc, "fvae_predict")
    else:
        _require_cuda(y, "returns")
        y = y.to(dtype=torch.float32).reshape(-1).contiguous()
        if y.numel() != S:
            raise ValueError("returns must have one entry per stock")
        keep["y"] = y
        with _on(dev):
            rc = L.fvae_elbo_forward(C.byref(shape), C.byref(panel), y.data_ptr(), date_ptr.data_ptr(), flat.data_ptr(),
                                     C.byref(noise), flags, prec, C.byref(outs), workspace.data_ptr(), workspace.numel(),
                                     _stream(dev))
        _cabi.check(rc, "fvae_elbo_forward")
    keep.update(out)
    st = StepState(shape, panel, noise, outs, flags, prec, keep, workspace, flat._version)
    return out, st


def elbo_backward(layout: ParamLayout, st: StepState, grad: Optional[torch.Tensor] = None) -> torch.Tensor:
    """d loss / d params for the forward that produced `st`, as one flat fp32 buffer (overwritten)."""
    L = _cabi.lib()
    t = st.tensors
    if grad is None:
        grad = torch.empty(layout.total, dtype=torch.float32, device=t["flat"].device)
    dev = t["flat"].device
    if st.flat_version is not None and t["flat"]._version != st.flat_version:
        raise RuntimeError("the parameters were modified in place between forward and backward (optimizer.step() / "
                           "load_state_dict before loss.backward()): the saved activations no longer match them")
    with _on(dev):
        rc = L.fvae_elbo_backward(C.byref(st.shape), C.byref(st.panel), t["y"].data_ptr(), t["date_ptr"].data_ptr(),
                                  t["flat"].data_ptr(), C.byref(st.noise), st.flags, st.precision, C.byref(st.outs),
                                  grad.data_ptr(), st.workspace.data_ptr(), st.workspace.numel(), _stream(dev))
    _cabi.check(rc, "fvae_elbo_backward")
    return grad


def rerun_front_forward(st: StepState) -> None:
    """Diagnostics: launch only the dominant tensor-core kernel again on the state of a bf16 forward (bench.py)."""
    dev = st.workspace.device
    with _on(dev):
        rc = _cabi.lib().fvae_debug_front_forward(C.byref(st.shape), C.byref(st.panel), st.workspace.data_ptr(),
                                                  st.workspace.numel(), _stream(dev))
    _cabi.check(rc, "fvae_debug_front_forward")


def philox_noise(seed: int, step: int, unit_base: int, S: int, K: int, device):
    """Diagnostics: (eps (S,), keep_mask (S, K) uint8) a philox=(seed, step, unit_base) step draws (fvae_debug_noise)."""
    eps = torch.empty(S, dtype=torch.float32, device=device)
    keep = torch.empty(S, K, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        rc = _cabi.lib().fvae_debug_noise(int(seed) & (2 ** 64 - 1), int(step), int(unit_base), S, K, eps.data_ptr(),
                                          keep.data_ptr(), C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _cabi.check(rc, "fvae_debug_noise")
    return eps, keep


def latent(st: StepState) -> torch.Tensor:
    """e = h_T (S, H) of the forward that produced `st` (a copy)."""
    L = _cabi.lib()
    ptr = L.fvae_workspace_latent(C.byref(st.shape), st.precision, st.workspace.data_ptr())
    off = ptr - st.workspace.data_ptr()
    S, H = st.shape.S, st.shape.H
    return st.workspace[off:off + S * H * 4].view(torch.float32).view(S, H).clone()


def heads_parts(layout: ParamLayout, flat: torch.Tensor, e: torch.Tensor, *, y: Optional[torch.Tensor] = None,
                z: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, train: bool = False,
                eps: Optional[torch.Tensor] = None, keep_mask: Optional[torch.Tensor] = None,
                philox: Optional[Tuple[int, int, int]] = None, want=()) -> Dict[str, torch.Tensor]:
    """The per-date sub-modules on caller-supplied stock latents e (N, H) of ONE date (fvae_heads_parts): FactorEncoder
    (needs y), FactorPredictor / AttentionLayer, AlphaLayer, BetaLayer, FactorDecoder (z = (mu, sigma) or the prior).
    `want` names the optional outputs to produce: "alpha", "beta", "context".  Forward only."""
    L = _cabi.lib()
    _require_cuda(flat, "parameters")
    _require_cuda(e, "stock_latent")
    dev = e.device
    if flat.device != dev:
        raise RuntimeError(f"parameters live on {flat.device}, stock_latent on {dev}")
    if e.dim() != 2 or e.shape[1] != layout.H:
        raise ValueError(f"stock_latent must be (N, {layout.H}), got {tuple(e.shape)}")
    e = e.detach().to(torch.float32).contiguous()
    N, H, K, M = e.shape[0], layout.H, layout.K, layout.M
    shape = _cabi.Shape(N, 1, 1, layout.C, H, K, M)
    f32 = dict(dtype=torch.float32, device=dev)
    keep: Dict[str, torch.Tensor] = dict(e=e)
    flags = _cabi.FLAG_TRAIN if train else 0
    if philox is not None:
        flags |= _cabi.FLAG_PHILOX
        noise = _philox_noise(philox, dev, keep)
    else:
        if eps is None:
            raise ValueError("either eps (and keep_mask in train mode) or philox=(seed, step, unit_base) is required")
        keep["eps"] = eps = eps.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if eps.numel() != N:
            raise ValueError("eps must have one entry per stock")
        km_ptr = None
        if train:
            if keep_mask is None or tuple(keep_mask.shape) != (N, K):
                raise ValueError("train mode needs keep_mask (N, K) or philox")
            keep["keep_mask"] = keep_mask = keep_mask.to(device=dev, dtype=torch.uint8).contiguous()
            km_ptr = keep_mask.data_ptr()
        noise = _cabi.Noise(eps.data_ptr(), km_ptr, 0, 0, 0)
    out = dict(loss=torch.empty(1, **f32), date_loss=torch.empty(1, **f32), yhat=torch.empty(N, **f32), mu_y=torch.empty(N, **f32),
               sigma_y=torch.empty(N, **f32), mu_post=torch.empty(K, **f32), sigma_post=torch.empty(K, **f32),
               mu_prior=torch.empty(K, **f32), sigma_prior=torch.empty(K, **f32))
    outs = _cabi.Outputs(*[out[n].data_ptr() for n, _ in _cabi.Outputs._fields_])
    parts = _cabi.Parts(None, None, None, None, None, None)
    if z is not None:
        keep["z_mu"] = zm = z[0].detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        keep["z_sigma"] = zs = z[1].detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if zm.numel() != K or zs.numel() != K:
            raise ValueError(f"factor_mu / factor_sigma must have {K} entries")
        parts.z_mu, parts.z_sigma = zm.data_ptr(), zs.data_ptr()
    if "alpha" in want:
        out["alpha_mu"], out["alpha_sigma"] = torch.empty(N, 1, **f32), torch.empty(N, 1, **f32)
        parts.alpha_mu, parts.alpha_sigma = out["alpha_mu"].data_ptr(), out["alpha_sigma"].data_ptr()
    if "beta" in want:
        out["beta"] = torch.empty(N, K, **f32)
        parts.beta = out["beta"].data_ptr()
    if "context" in want:
        out["context"] = torch.empty(K, H, **f32)
        parts.context = out["context"].data_ptr()
    y_ptr = None
    if y is not None:
        _require_cuda(y, "returns")
        keep["y"] = y = y.detach().to(dtype=torch.float32).reshape(-1).contiguous()
        if y.numel() != N:
            raise ValueError("returns must have one entry per stock")
        y_ptr = y.data_ptr()
    need = L.fvae_workspace_bytes(C.byref(shape), _cabi.PREC_FP32)
    if need < 0:
        _cabi.check(int(need), "fvae_workspace_bytes")
    ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
    date_ptr = single_date_ptr(N, dev)
    with _on(dev):
        rc = L.fvae_heads_parts(C.byref(shape), e.data_ptr(), y_ptr, date_ptr.data_ptr(), flat.data_ptr(), C.byref(noise), flags,
                                C.byref(parts), C.byref(outs), ws.data_ptr(), ws.numel(), _stream(dev))
    _cabi.check(rc, "fvae_heads_parts")
    return out


def fe_forward(layout: ParamLayout, flat: torch.Tensor, x: torch.Tensor, precision: str = "fp32"):
    L = _cabi.lib()
    _require_cuda(flat, "parameters")
    _require_cuda(x, "x")
    x, panel = _panel(x)
    S, T, Cf = x.shape
    shape = _cabi.Shape(S, 1, T, Cf, layout.H, layout.K, layout.M)
    prec = PRECISIONS[precision]
    need = L.fvae_workspace_bytes(C.byref(shape), prec)
    if need < 0:
        _cabi.check(int(need), "fvae_workspace_bytes")
    ws = torch.empty(int(need), dtype=torch.uint8, device=x.device)
    e = torch.empty(S, layout.H, dtype=torch.float32, device=x.device)
    with _on(x.device):
        rc = L.fvae_fe_forward(C.byref(shape), C.byref(panel), flat.data_ptr(), prec, e.data_ptr(), ws.data_ptr(), ws.numel(),
                               _stream(x.device))
    _cabi.check(rc, "fvae_fe_forward")
    return e, (shape, panel, prec, ws, x, flat)


def fe_backward(layout: ParamLayout, saved, de: torch.Tensor) -> torch.Tensor:
    L = _cabi.lib()
    shape, panel, prec, ws, x, flat = saved
    de = de.to(dtype=torch.float32).contiguous()
    grad = torch.zeros(layout.total, dtype=torch.float32, device=x.device)
    with _on(x.device):
        rc = L.fvae_fe_backward(C.byref(shape), C.byref(panel), flat.data_ptr(), prec, de.data_ptr(), grad.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream(x.device))
    _cabi.check(rc, "fvae_fe_backward")
    return grad
