"""Generate tests/golden/*.npz from the LIVE reference (/root/reference/module.py).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference is not on the GPU
box); the fixtures it writes are committed and travel.  Usage:

    python oracle/gen_golden.py            # rewrites every tests/golden/case_*.npz

How the reference is driven (nothing in /root/reference is modified or copied):
  * modules are built with the reference constructors exactly as main.py:27-33 does,
    under torch.manual_seed(seed);
  * `FactorDecoder.reparameterize` (module.py:103-105) is replaced at run time by a
    function that uses an injected eps and records (mu_y, sigma_y) -- these are
    temporaries in the reference (module.py:120-121);
  * each `AttentionLayer.dropout` (module.py:132) is replaced by a module that
    multiplies by an injected keep-mask / 0.9 (train mode) or is the identity (eval);
  * one reference step per date: zero_grad -> forward -> backward (train_model.py:26-29);
    the batched loss/gradients are the means over dates.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class _InjectedDropout(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.keep = None            # (N,) float 0/1 or None

    def forward(self, scores):
        if self.keep is None:
            return scores
        return scores * self.keep / 0.9


def build_reference(ref, C, H, K, M, seed):
    torch.manual_seed(seed)
    fe = ref.FeatureExtractor(num_latent=C, hidden_size=H)
    enc = ref.FactorEncoder(num_factors=K, num_portfolio=M, hidden_size=H)
    dec = ref.FactorDecoder(ref.AlphaLayer(H), ref.BetaLayer(H, K))
    pred = ref.FactorPredictor(H, K)
    model = ref.FactorVAE(fe, enc, dec, pred)
    for layer in model.factor_predictor.attention_layers:
        layer.dropout = _InjectedDropout()
    return model


def run_reference(ref, model, xs, ys, epss, masks, train):
    """Loop the unmodified reference over dates; returns outputs and mean gradients."""
    rec = {}

    def reparam(self, mu, sigma):
        rec["mu_y"], rec["sigma_y"] = mu.detach().clone(), sigma.detach().clone()
        return mu + rec["eps"] * sigma

    ref.FactorDecoder.reparameterize = reparam
    model.train(train)
    B = len(xs)
    names = [n for n, _ in model.named_parameters()]
    gsum = {n: torch.zeros_like(p) for n, p in model.named_parameters()}
    out = {k: [] for k in ("date_loss", "yhat", "mu_y", "sigma_y", "mu_post", "sigma_post", "mu_prior", "sigma_prior", "e")}
    for d in range(B):
        rec["eps"] = epss[d].reshape(-1, 1)
        for k, layer in enumerate(model.factor_predictor.attention_layers):
            layer.dropout.keep = None if masks is None else masks[d][k]
        model.zero_grad(set_to_none=True)
        loss, yhat, mu_post, sigma_post, mu_prior, sigma_prior = model(xs[d], ys[d])
        loss.backward()
        with torch.no_grad():
            e = model.feature_extractor(xs[d])
        for n, p in model.named_parameters():
            if p.grad is not None:
                gsum[n] += p.grad / B
        out["date_loss"].append(loss.detach().reshape(1))
        out["yhat"].append(yhat.detach().reshape(-1))
        out["mu_y"].append(rec["mu_y"].reshape(-1))
        out["sigma_y"].append(rec["sigma_y"].reshape(-1))
        out["e"].append(e)
        for k2, v in (("mu_post", mu_post), ("sigma_post", sigma_post), ("mu_prior", mu_prior), ("sigma_prior", sigma_prior)):
            out[k2].append(v.detach().reshape(1, -1))
    res = {k: torch.cat(v) for k, v in out.items()}
    res["loss"] = res["date_loss"].mean()
    return res, gsum, names


def run_reference_prediction(ref, model, xs, epss):
    rec = {}

    def reparam(self, mu, sigma):
        rec["mu_y"], rec["sigma_y"] = mu.detach().clone(), sigma.detach().clone()
        return mu + rec["eps"] * sigma

    ref.FactorDecoder.reparameterize = reparam
    model.eval()
    for layer in model.factor_predictor.attention_layers:
        layer.dropout.keep = None
    ys, mus, sgs = [], [], []
    with torch.no_grad():
        for d in range(len(xs)):
            rec["eps"] = epss[d].reshape(-1, 1)
            ys.append(model.prediction(xs[d]).reshape(-1))
            mus.append(rec["mu_y"].reshape(-1))
            sgs.append(rec["sigma_y"].reshape(-1))
    return torch.cat(ys), torch.cat(mus), torch.cat(sgs)


CASES = [
    # name, C, H, K, M, T, stocks per date, train, seed, tweak
    dict(name="train_ragged", C=158, H=12, K=6, M=16, T=5, ns=[9, 16, 5], train=True, seed=1),
    dict(name="eval_k20", C=158, H=20, K=20, M=128, T=4, ns=[7, 7], train=False, seed=2),
    dict(name="single_stock", C=158, H=8, K=4, M=8, T=3, ns=[1, 3], train=True, seed=3),
    dict(name="k_ne_h_t1", C=158, H=16, K=5, M=24, T=1, ns=[11], train=True, seed=4),
    dict(name="guard_inf_query", C=158, H=12, K=6, M=16, T=3, ns=[6, 8], train=False, seed=5, tweak="inf_query"),
    dict(name="sigma_zero_clamp", C=158, H=12, K=6, M=16, T=3, ns=[6, 8], train=True, seed=6, tweak="sigma_zero"),
    dict(name="small_c20", C=20, H=20, K=8, M=128, T=6, ns=[10, 13], train=True, seed=7),
    dict(name="cfg1_shape", C=158, H=20, K=20, M=128, T=20, ns=[64], train=True, seed=42),
]


def make_case(ref, cs):
    C, H, K, M, T = cs["C"], cs["H"], cs["K"], cs["M"], cs["T"]
    model = build_reference(ref, C, H, K, M, cs["seed"])
    tw = cs.get("tweak")
    with torch.no_grad():
        if tw == "inf_query":          # head 2 trips the NaN/Inf guard of module.py:149-150
            model.factor_predictor.attention_layers[2].query[3] = float("inf")
        if tw == "sigma_zero":         # softplus underflows to exactly 0 -> module.py:117 and :264-265 clamps fire
            model.factor_encoder.linear_sigma.weight[1].zero_()
            model.factor_encoder.linear_sigma.bias[1] = -200.0
            model.factor_predictor.sigma_layer.weight.zero_()
            model.factor_predictor.sigma_layer.bias.fill_(-200.0)
    g = torch.Generator().manual_seed(1000 + cs["seed"])
    xs = [torch.randn(n, T, C, generator=g).clamp_(-3, 3) for n in cs["ns"]]
    ys = [torch.randn(n, 1, generator=g) for n in cs["ns"]]
    epss = [torch.randn(n, generator=g) for n in cs["ns"]]
    masks = None
    if cs["train"]:
        masks = [(torch.rand(K, n, generator=g) >= 0.1).float() for n in cs["ns"]]
    res, grads, names = run_reference(ref, model, xs, ys, epss, masks, cs["train"])
    py, pmu, psg = run_reference_prediction(ref, model, xs, epss)
    blob = {}
    for n, v in model.state_dict().items():
        blob["param:" + n] = v.detach().numpy().copy()
    for n in names:
        blob["grad:" + n] = grads[n].numpy().copy()
    for k, v in res.items():
        blob["out:" + k] = v.numpy().copy()
    blob["pred:yhat"], blob["pred:mu_y"], blob["pred:sigma_y"] = py.numpy(), pmu.numpy(), psg.numpy()
    blob["in:x"] = torch.cat(xs).numpy()
    blob["in:y"] = torch.cat(ys).reshape(-1).numpy()
    blob["in:eps"] = torch.cat(epss).numpy()
    blob["in:date_ptr"] = np.cumsum([0] + cs["ns"]).astype(np.int32)
    if masks is not None:
        blob["in:keep_mask"] = torch.cat(masks, dim=1).numpy().astype(np.uint8)     # (K, S)
    blob["meta:dims"] = np.array([C, H, K, M, T, int(cs["train"])], dtype=np.int32)
    return blob


def main():
    if not os.path.isdir(REF):
        raise SystemExit("the reference is not mounted here; fixtures are generated in the build container only")
    sys.path.insert(0, REF)
    import module as ref   # noqa: the unmodified reference

    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)
    for cs in CASES:
        blob = make_case(ref, cs)
        path = os.path.join(OUT, f"case_{cs['name']}.npz")
        np.savez_compressed(path, **blob)
        print(f"{path}: loss={float(blob['out:loss']):.6f} bytes={os.path.getsize(path)}")


if __name__ == "__main__":
    main()
