"""The oracle (oracle/restatement.py) against the golden vectors produced by the live reference.

CPU only.  This is what pins the oracle: every fixture in tests/golden/ was written by
oracle/gen_golden.py running the unmodified /root/reference/module.py.
"""
import pytest
import torch

from conftest import golden_cases, load_golden, split_by_date
from oracle import restatement as R


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("name", golden_cases())
def test_restatement_matches_reference_forward_and_grads(name):
    g = load_golden(name)
    xs, ys, epss, masks = split_by_date(g)
    # The clamp fixture depends on softplus underflowing to exactly 0, which only happens in the
    # reference's own fp32 arithmetic -> that case is restated in fp32, everything else in fp64.
    dtype = torch.float32 if name == "sigma_zero_clamp" else torch.float64
    out, grads = R.elbo_step(g["params"], xs, ys, epss, masks, need_grad=True, dtype=dtype)
    ref = g["out"]
    # the fixture is fp32 arithmetic; the restatement runs in fp64 -> agreement to fp32 round-off
    assert abs(float(out["loss"]) - float(ref["loss"])) <= 2e-5 * abs(float(ref["loss"]))
    for k in ("yhat", "mu_y", "sigma_y", "mu_post", "sigma_post", "mu_prior", "sigma_prior", "e"):
        assert _rel(out[k], ref[k]) < 2e-5, k
    # per-tensor: ||ours - ref|| <= rtol*||ref|| + atol, atol scaled by the largest gradient entry
    # (several gradients are mathematically zero -- softmax over stocks is shift invariant -- and
    #  are pure round-off on both sides; heads behind a tripped guard are exactly zero)
    gmax = max(float(v.abs().max()) for v in g["grads"].values())
    for k, gr in g["grads"].items():
        ours = grads[k].double()
        err = float((ours - gr.double()).norm())
        tol = 5e-4 * float(gr.double().norm()) + 2e-6 * gmax * (gr.numel() ** 0.5)
        assert err <= tol, (k, err, tol)


@pytest.mark.parametrize("name", golden_cases())
def test_restatement_prediction(name):
    g = load_golden(name)
    xs, ys, epss, masks = split_by_date(g)
    p = {k: v.double() for k, v in g["params"].items()}
    ys_, mus, sgs = [], [], []
    for x, eps in zip(xs, epss):
        o = R.prediction_one_date(p, x.double(), eps.double())
        ys_.append(o["yhat"]); mus.append(o["mu_y"]); sgs.append(o["sigma_y"])
    assert _rel(torch.cat(ys_), g["pred"]["yhat"]) < 2e-5
    assert _rel(torch.cat(mus), g["pred"]["mu_y"]) < 2e-5
    assert _rel(torch.cat(sgs), g["pred"]["sigma_y"]) < 2e-5


def test_guard_case_really_trips():
    g = load_golden("guard_inf_query")
    for k in ("query", "key_layer.weight", "key_layer.bias", "value_layer.weight", "value_layer.bias"):
        assert float(g["grads"][f"factor_predictor.attention_layers.2.{k}"].abs().max()) == 0.0


def test_clamp_case_really_trips():
    g = load_golden("sigma_zero_clamp")
    assert float(g["out"]["sigma_post"][:, 1].max()) == pytest.approx(1e-6)
    assert float(g["out"]["sigma_prior"].max()) == pytest.approx(1e-6)


@pytest.mark.parametrize("name", [n for n in golden_cases() if n not in ("guard_inf_query",)])
def test_cpu_port_matches_reference(name):
    """oracle/cpu_port.py (the timed CPU baseline) reproduces the reference's loss and gradients."""
    from oracle.cpu_port import CpuPort
    g = load_golden(name)
    xs, ys, epss, masks = split_by_date(g)
    port = CpuPort(g["params"])
    B = len(xs)
    total, gsum = 0.0, {k: torch.zeros_like(v) for k, v in port.p.items()}
    for d in range(B):
        port.zero_grad()
        loss, yhat, mu_y, sg_y = port.step_forward(xs[d], ys[d], train=g["dims"]["train"], eps=epss[d],
                                                    keep_mask=None if masks is None else masks[d].float())
        loss.backward()
        total += float(loss) / B
        for k, v in port.p.items():
            if v.grad is not None:
                gsum[k] += v.grad / B
    assert abs(total - float(g["out"]["loss"])) <= 1e-5 * abs(float(g["out"]["loss"]))
    gmax = max(float(v.abs().max()) for v in g["grads"].values())
    for k, gr in g["grads"].items():
        err = float((gsum[k] - gr).norm())
        assert err <= 1e-4 * float(gr.norm()) + 2e-6 * gmax * gr.numel() ** 0.5, k
