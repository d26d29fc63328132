"""GPU parity of the configuration bench.py actually times: bf16 tensor-core mode (tcgen05) + in-kernel Philox noise on a
bf16 panel, at the per-date shapes of BASELINE.json's configs, several 128-stock tiles per date -- against the CPU oracle
(oracle/restatement.py, fp64), which replays the step with the kernel's OWN noise (fvae_debug_noise evaluates the same device
functions the kernels call: eps of reference module.py:104, dropout keep decisions of module.py:132,144).

Stated tolerances (BASELINE.md section 4 / SURVEY 8c), bf16 tensor-core mode vs the fp32/fp64 reference arithmetic:
    ELBO rel <= 2e-2, mu_y abs <= 1e-2, sigma_y rel <= 2e-2, gradient cosine >= 0.999 and rel-L2 <= 3e-2.
"""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

C = 158
# per-date shapes of BASELINE.json configs[1..4] (H = K, M = 128); B = dates replayed through the oracle
BENCH_SHAPES = {
    "cfg2": dict(B=8, N=300, T=20, H=20, K=20),      # 3 tiles per date (the last one ragged: 300 = 2*128 + 44)
    "cfg3": dict(B=2, N=500, T=60, H=60, K=60),      # 4 tiles per date
    "cfg4": dict(B=2, N=1000, T=20, H=48, K=48),     # 8 tiles per date
    "cfg5": dict(B=2, N=3000, T=60, H=60, K=60),     # 24 tiles per date
}


def _build(H, K, M=128, seed=42):
    """Same construction as bench.py build_params: reference initialisers under torch.manual_seed(42) (main.py:109)."""
    import factorvae_b200 as fb
    torch.manual_seed(seed)
    m = fb.FactorVAE(fb.FeatureExtractor(C, H), fb.FactorEncoder(K, M, H), fb.FactorDecoder(fb.AlphaLayer(H), fb.BetaLayer(H, K)),
                     fb.FactorPredictor(H, K))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _panel(B, N, T, dev, dtype=torch.bfloat16):
    """bench.py's synthetic panel: per global date id, N(0,1) clipped to +-3, rounded to the panel dtype; rows padded to a
    16-byte pitch (160 elements, 158 used) exactly as bench.py lays it out -- the layout the TMA kernels take."""
    store = torch.full((B * N, T, 160), 7.0, dtype=dtype, device=dev)           # padding holds junk: it must never be read
    x = store[:, :, :C]
    y = torch.empty(B * N, dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev)
    for d in range(B):
        gen.manual_seed(1234 + d)
        x[d * N:(d + 1) * N] = torch.randn(N, T, C, generator=gen, device=dev).clamp_(-3, 3).to(dtype)
        y[d * N:(d + 1) * N] = torch.randn(N, generator=gen, device=dev)
    return x, y


def _cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _oracle(params, x, y, eps, keep, B, N):
    from oracle import restatement as R
    xs = [x[d * N:(d + 1) * N].float().cpu() for d in range(B)]        # the bf16 panel values, exactly
    ys = [y[d * N:(d + 1) * N].cpu() for d in range(B)]
    es = [eps[d * N:(d + 1) * N].cpu() for d in range(B)]
    ms = [keep[d * N:(d + 1) * N].t().contiguous().float().cpu() for d in range(B)]      # (K, N) per date
    return R.elbo_step(params, xs, ys, es, ms, need_grad=True, dtype=torch.float64)


def _run(shape, precision, dev, philox=(42, 1, 0)):
    from factorvae_b200 import engine
    B, N, T, H, K = (shape[k] for k in "BNTHK")
    params = _build(H, K)
    L = engine.ParamLayout(C, H, K, 128)
    flat = L.pack(params, dev)
    x, y = _panel(B, N, T, dev)
    ptr = engine.uniform_date_ptr(B, N, dev)
    out, st = engine.elbo_forward(L, flat, x, y, ptr, train=True, precision=precision, philox=philox)
    grad = engine.elbo_backward(L, st).clone()
    eps, keep = engine.philox_noise(philox[0], philox[1], philox[2], B * N, K, dev)
    torch.cuda.synchronize()
    ref, rgrads = _oracle(params, x, y, eps, keep, B, N)
    return L, out, grad, ref, rgrads, eps


def _grad_vectors(L, grad, rgrads):
    allg = torch.cat([L.view(grad, k).reshape(-1).double().cpu() for k in rgrads])
    allr = torch.cat([v.reshape(-1).double() for v in rgrads.values()])
    return allg, allr


@pytest.mark.parametrize("cfg", sorted(BENCH_SHAPES))
def test_bf16_tc_philox_vs_oracle_at_bench_shapes(cfg, cuda_device):
    """The timed mode (bf16 tcgen05 + Philox, bf16 panel) against the oracle at the stated tolerances."""
    from factorvae_b200 import engine
    shape = BENCH_SHAPES[cfg]
    if not engine.tc_supported(C, shape["H"]):
        pytest.skip("tensor-core path does not cover this shape")
    L, out, grad, ref, rgrads, eps = _run(shape, "bf16", cuda_device)
    loss, rl = float(out["loss"]), float(ref["loss"])
    assert abs(loss - rl) <= 2e-2 * abs(rl), (loss, rl)
    dl = (out["date_loss"].double().cpu() - ref["date_loss"]).abs() / ref["date_loss"].abs()
    assert float(dl.max()) <= 2e-2, float(dl.max())
    mu_err = float((out["mu_y"].double().cpu() - ref["mu_y"]).abs().max())
    assert mu_err <= 1e-2, mu_err
    sg_err = float(((out["sigma_y"].double().cpu() - ref["sigma_y"]).abs() / ref["sigma_y"]).max())
    assert sg_err <= 2e-2, sg_err
    for k in ("mu_post", "mu_prior"):
        assert float((out[k].double().cpu() - ref[k]).abs().max()) <= 1e-2, k
    for k in ("sigma_post", "sigma_prior"):
        assert float(((out[k].double().cpu() - ref[k]).abs() / ref[k]).max()) <= 2e-2, k
    allg, allr = _grad_vectors(L, grad, rgrads)
    cos, rel = _cos(allg, allr), float((allg - allr).norm() / allr.norm())
    print(f"{cfg}: loss {loss:.6f} vs oracle {rl:.6f} (rel {abs(loss - rl) / abs(rl):.2e}); mu_y abs {mu_err:.2e}; sigma_y rel "
          f"{sg_err:.2e}; grad cos {cos:.6f} rel-L2 {rel:.2e}")
    assert cos >= 0.999, cos
    assert rel <= 3e-2, rel


def test_fp32_kernels_philox_replay_is_exact_at_cfg2_shape(cuda_device):
    """fp32 mode with in-kernel noise vs the oracle fed with fvae_debug_noise's tensors at the fp32 tolerance (1e-5): proves that
    the replayed eps and keep decisions ARE the ones the step's kernels drew (a single differing keep bit moves the loss by
    far more than 1e-5)."""
    L, out, grad, ref, rgrads, eps = _run(BENCH_SHAPES["cfg2"], "fp32", cuda_device, philox=(7, 3, 1000))
    assert abs(float(out["loss"]) - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    assert float((out["mu_y"].double().cpu() - ref["mu_y"]).abs().max()) <= 1e-5 * max(1.0, float(ref["mu_y"].abs().max()))
    assert float(((out["sigma_y"].double().cpu() - ref["sigma_y"]).abs() / ref["sigma_y"]).max()) <= 1e-5
    z = (out["yhat"] - out["mu_y"]) / out["sigma_y"]                     # eps recovered from the step's own outputs
    assert float((z - eps).abs().max()) <= 1e-3
    allg, allr = _grad_vectors(L, grad, rgrads)
    assert float((allg - allr).norm() / allr.norm()) <= 1e-4


def test_philox_keep_rate_and_eps_moments(cuda_device):
    """nn.Dropout(0.1) on the attention scores (module.py:132,144): every head keeps 0.9 +- 0.005 of the stocks, for several
    (seed, step, unit_base) keys; eps is standard normal (mean, variance, fourth moment); distinct keys give distinct streams."""
    from factorvae_b200 import engine
    S, K = 200_000, 60
    seen = []
    for seed, step, base in [(42, 1, 0), (42, 2, 0), (7, 1, 76_800), (2 ** 40 + 5, 123_456, 2 ** 33)]:
        eps, keep = engine.philox_noise(seed, step, base, S, K, cuda_device)
        rate = keep.float().mean(dim=0)
        assert float((rate - 0.9).abs().max()) <= 0.005, (seed, step, base, rate.tolist())
        assert abs(float(keep.float().mean()) - 0.9) <= 0.001
        # neighbouring heads / neighbouring stocks are uncorrelated
        kf = keep.float() - 0.9
        assert abs(float((kf[:, :-1] * kf[:, 1:]).mean())) <= 1e-3
        assert abs(float((kf[:-1] * kf[1:]).mean())) <= 1e-3
        assert abs(float(eps.mean())) <= 0.01 and abs(float(eps.var()) - 1.0) <= 0.02
        assert abs(float((eps ** 4).mean()) - 3.0) <= 0.15
        seen.append((eps[:4096].clone(), keep[:4096].clone()))
    for i in range(len(seen)):
        for j in range(i + 1, len(seen)):
            assert not torch.equal(seen[i][0], seen[j][0]) and not torch.equal(seen[i][1], seen[j][1])
    # shard invariance of the stream itself: units [a, b) drawn alone == the slice of the whole
    e_all, k_all = engine.philox_noise(42, 9, 0, 5000, 20, cuda_device)
    e_part, k_part = engine.philox_noise(42, 9, 1234, 1000, 20, cuda_device)
    assert torch.equal(e_all[1234:2234], e_part) and torch.equal(k_all[1234:2234], k_part)


def test_bf16_tc_sigma_zero_clamp_semantics(cuda_device):
    """module.py:117 / :264-265 in the bf16 tensor-core mode: weights that drive softplus to exactly 0 (the fixture written by
    the live reference) -- the clamped sigmas are exactly 1e-6, the gradients the in-place clamp blocks are exactly zero, the
    rest of the step stays within the bf16 tolerances.  (The loss of this fixture is dominated by the 1/sigma_prior^2 = 1e12
    amplification of (mu_post - mu_prior)^2, so the ELBO is compared through its un-amplified parts.)"""
    from factorvae_b200 import engine
    g = load_golden("sigma_zero_clamp")
    d = g["dims"]
    L = engine.ParamLayout(d["C"], d["H"], d["K"], d["M"])
    flat = L.pack(g["params"], cuda_device)
    out, st = engine.elbo_forward(L, flat, g["inp"]["x"].to(cuda_device), g["inp"]["y"].to(cuda_device),
                                  g["inp"]["date_ptr"].to(cuda_device), eps=g["inp"]["eps"].to(cuda_device),
                                  keep_mask=g["inp"]["keep_mask"].t().contiguous().to(cuda_device), train=True, precision="bf16")
    grad = engine.elbo_backward(L, st)
    ref = g["out"]
    # the clamps fired exactly where the reference's did
    assert torch.equal(out["sigma_post"].cpu() == 1e-6, ref["sigma_post"] == 1e-6)
    assert bool((ref["sigma_post"][:, 1] == 1e-6).all()) and bool((out["sigma_post"][:, 1] == 1e-6).all())
    assert bool((out["sigma_prior"] == 1e-6).all()) and bool((ref["sigma_prior"] == 1e-6).all())
    # gradients blocked by the in-place clamp are exact zeros, as in the reference
    for name in ("factor_predictor.sigma_layer.weight", "factor_predictor.sigma_layer.bias"):
        assert float(g["grads"][name].abs().max()) == 0.0
        assert float(L.view(grad, name).abs().max()) == 0.0, name
    assert float(g["grads"]["factor_encoder.linear_sigma.weight"][1].abs().max()) == 0.0
    assert float(L.view(grad, "factor_encoder.linear_sigma.weight")[1].abs().max()) == 0.0
    assert float(L.view(grad, "factor_encoder.linear_sigma.bias")[1].abs()) == 0.0
    # un-amplified outputs within the stated bf16 tolerances
    assert float((out["mu_y"].cpu() - ref["mu_y"]).abs().max()) <= 1e-2 * max(1.0, float(ref["mu_y"].abs().max()))
    assert float(((out["sigma_y"].cpu() - ref["sigma_y"]).abs() / ref["sigma_y"]).max()) <= 2e-2
    assert float((out["mu_post"].cpu() - ref["mu_post"]).abs().max()) <= 1e-2
    assert float((out["mu_prior"].cpu() - ref["mu_prior"]).abs().max()) <= 1e-2
    assert torch.isfinite(grad).all() and torch.isfinite(out["loss"]).all()
    # the amplified ELBO: exact given the kernel's own mu/sigma (fp64 KL of the OUTPUTS + the mse of the OUTPUTS)
    yv, ptr = g["inp"]["y"].double(), g["inp"]["date_ptr"].tolist()
    o = {k: v.double().cpu() for k, v in out.items()}
    tot = 0.0
    for dd in range(len(ptr) - 1):
        a, b = ptr[dd], ptr[dd + 1]
        mse = ((o["yhat"][a:b] - yv[a:b]) ** 2).mean()
        m1, s1, m2, s2 = o["mu_post"][dd], o["sigma_post"][dd], o["mu_prior"][dd], o["sigma_prior"][dd]
        kl = (torch.log(s2 / s1) + (s1 ** 2 + (m1 - m2) ** 2) / (2 * s2 ** 2) - 0.5).sum()
        tot += float(mse + kl) / (len(ptr) - 1)
    assert abs(float(out["loss"]) - tot) <= 1e-4 * abs(tot), (float(out["loss"]), tot)


def test_bf16_tc_prediction_vs_golden(cuda_device):
    """FactorVAE.prediction (module.py:273-278) in bf16 tensor-core mode against the golden `pred:*` arrays written by the
    live reference (every fixture the tensor-core path covers)."""
    from conftest import golden_cases
    from factorvae_b200 import engine
    ran = 0
    for name in golden_cases():
        g = load_golden(name)
        d = g["dims"]
        if not engine.tc_supported(d["C"], d["H"]) or name == "sigma_zero_clamp":
            continue
        L = engine.ParamLayout(d["C"], d["H"], d["K"], d["M"])
        flat = L.pack(g["params"], cuda_device)
        out, _ = engine.elbo_forward(L, flat, g["inp"]["x"].to(cuda_device), None, g["inp"]["date_ptr"].to(cuda_device),
                                     eps=g["inp"]["eps"].to(cuda_device), train=False, precision="bf16", predict=True)
        pm, ps, py = g["pred"]["mu_y"], g["pred"]["sigma_y"], g["pred"]["yhat"]
        assert float((out["mu_y"].cpu() - pm).abs().max()) <= 1e-2 * max(1.0, float(pm.abs().max())), name
        assert float(((out["sigma_y"].cpu() - ps).abs() / ps).max()) <= 2e-2, name
        assert float((out["yhat"].cpu() - py).abs().max()) <= 3e-2 * max(1.0, float(py.abs().max())), name
        ran += 1
    assert ran >= 5
