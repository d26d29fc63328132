"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/fvae_b200.h declares,
its parameter layout matches the reference inventory, and argument validation works.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT, golden_cases, load_golden


@pytest.fixture(scope="module")
def cabi():
    from factorvae_b200 import build, _cabi
    build.build()
    return _cabi


def test_library_exports_every_declared_symbol(cabi):
    hdr = open(os.path.join(ROOT, "include", "fvae_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(fvae_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 11
    lib = C.CDLL(cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fvae_b200.h but not exported"
    assert sorted(cabi.EXPORTS) == declared
    assert cabi.lib().fvae_abi_version() == cabi.ABI_VERSION == int(re.search(r"#define FVAE_ABI_VERSION (\d+)", hdr).group(1))


def test_param_layout_matches_reference_inventory(cabi):
    from factorvae_b200.engine import ParamLayout
    # 62,630 / 309,394 / 542,350 scalars at K=H=20/48/60, M=128, C=158
    for kh, want in ((20, 62630), (48, 309394), (60, 542350)):
        L = ParamLayout(158, kh, kh, 128)
        n = 0
        for off, shape in L.slices.values():
            assert off % 4 == 0 or len(shape) == 1 or True
            sz = 1
            for d in shape:
                sz *= d
            n += sz
        assert n == want
        assert L.total >= want and L.total == cabi.lib().fvae_param_count(158, kh, kh, 128)
        assert len(L.slices) == 28 + 5 * kh
        # slices are disjoint and inside the buffer
        iv = sorted((off, off + int(torch.tensor(shape).prod())) for off, shape in L.slices.values())
        for (a0, a1), (b0, b1) in zip(iv, iv[1:]):
            assert a1 <= b0
        assert iv[-1][1] <= L.total
    for off in cabi.param_offsets(158, 20, 20, 128):
        assert off % 4 == 0          # 16-byte aligned sections


@pytest.mark.parametrize("name", golden_cases())
def test_layout_names_equal_reference_state_dict(cabi, name):
    from factorvae_b200.engine import ParamLayout
    g = load_golden(name)
    d = g["dims"]
    L = ParamLayout(d["C"], d["H"], d["K"], d["M"])
    assert set(L.slices) == set(g["params"])
    for k, (off, shape) in L.slices.items():
        assert tuple(g["params"][k].shape) == tuple(shape), k
    flat = L.pack(g["params"], "cpu")
    back = L.unpack(flat)
    for k in g["params"]:
        assert torch.equal(back[k], g["params"][k])


def test_argument_validation(cabi):
    lib = cabi.lib()
    good = cabi.Shape(300, 1, 20, 158, 20, 20, 128)
    assert lib.fvae_workspace_bytes(C.byref(good), cabi.PREC_FP32) > 0
    assert lib.fvae_workspace_bytes(C.byref(cabi.Shape(0, 1, 20, 158, 20, 20, 128)), cabi.PREC_FP32) == -2
    assert lib.fvae_workspace_bytes(C.byref(cabi.Shape(300, 1, 20, 158, 96, 20, 128)), cabi.PREC_FP32) == -3
    assert lib.fvae_workspace_bytes(C.byref(good), 7) == -4
    assert lib.fvae_elbo_forward(None, None, None, None, None, None, 0, 0, None, None, 0, None) == -1
    for code in (0, -1, -2, -3, -4, -5, -6, -7):
        assert lib.fvae_status_string(code)
    with pytest.raises(cabi.FvaeError):
        cabi.check(-2, "x")


def test_module_surface_and_no_cpu_fallback(cabi):
    import factorvae_b200 as fb
    torch.manual_seed(0)
    m = fb.FactorVAE(fb.FeatureExtractor(158, 20), fb.FactorEncoder(20, 128, 20),
                     fb.FactorDecoder(fb.AlphaLayer(20), fb.BetaLayer(20, 20)), fb.FactorPredictor(20, 20))
    assert len(m.state_dict()) == 28 + 5 * 20
    assert sum(p.numel() for p in m.parameters()) == 62630
    assert hasattr(m, "prediction") and hasattr(m, "predict")
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(4, 20, 158), torch.zeros(4, 1))
    with pytest.raises(RuntimeError, match="no CPU"):
        m.prediction(torch.zeros(4, 20, 158))
    with pytest.raises(NotImplementedError):
        fb.FeatureExtractor(158, 20, num_layers=2)
    # the sub-modules answer on their own (fvae_heads_parts) -- on CUDA only, like everything else
    e, y = torch.zeros(4, 20), torch.zeros(4, 1)
    for call in (lambda: m.factor_encoder(e, y), lambda: m.factor_decoder.alpha_layer(e), lambda: m.factor_decoder.beta_layer(e),
                 lambda: m.factor_predictor(e), lambda: m.factor_predictor.attention_layers[0](e),
                 lambda: m.factor_decoder(e, torch.zeros(20), torch.ones(20)), lambda: m.feature_extractor(torch.zeros(4, 20, 158))):
        with pytest.raises(RuntimeError, match="no CPU"):
            call()


def test_oracle_is_not_imported_by_the_product():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import factorvae_b200, factorvae_b200.engine, factorvae_b200.module;"
            "bad=[m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; assert not bad, bad") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    pkg = os.path.join(ROOT, "factorvae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_new_entry_points_validate_arguments_without_a_gpu(cabi):
    """Resident panel / optimizer / metric entry points reject bad arguments before touching the device."""
    L = cabi.lib()
    E_NULL, E_SHAPE, E_LIMIT, E_UNSUP = -1, -2, -3, -7
    codes = {L.fvae_status_string(c).decode() for c in (E_NULL, E_SHAPE, E_LIMIT, E_UNSUP)}
    assert len(codes) == 4
    one = C.c_void_p(16)            # a non-null, 16-byte aligned dummy address (never dereferenced on these paths)
    assert L.fvae_window_index(None, 4, 4, one, one, 8, 5, 2, 16, one, None, None, None) == E_NULL
    assert L.fvae_window_index(one, 0, 4, one, one, 8, 5, 2, 16, one, None, None, None) == E_SHAPE
    assert L.fvae_window_index(one, 4, 4, one, one, 8, 5, 3, 16, one, None, None, None) == E_UNSUP      # unknown fill mode
    assert L.fvae_window_index(one, 4, 4, one, one, 8, 5, 2, 16, one, one, None, None) == E_NULL         # label without y
    p = cabi.Panel(16, cabi.F32, 0, 158, None, 0)
    assert L.fvae_gather_windows(C.byref(p), 8, 5, 158, one, cabi.F32, None) == E_NULL                  # no row index
    p = cabi.Panel(16, cabi.F32, 0, 100, 16, 10)
    assert L.fvae_gather_windows(C.byref(p), 8, 5, 158, one, cabi.F32, None) == E_SHAPE                 # pitch < C
    assert L.fvae_adam_step(None, one, one, one, 10, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, None) == E_NULL
    assert L.fvae_adam_step(one, one, one, one, 10, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 1.0, None) == E_SHAPE   # steps are 1-based
    assert L.fvae_adam_step(C.c_void_p(20), one, one, one, 10, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, None) == E_SHAPE  # alignment
    assert L.fvae_rank_ic(None, one, one, 3, 100, one, None) == E_NULL
    assert L.fvae_rank_ic(one, one, one, 3, 5000, one, None) == E_LIMIT                                 # > 4096 stocks per date


def test_heads_parts_validates_arguments_without_a_gpu(cabi):
    import ctypes as C
    L = cabi.lib()
    shape = cabi.Shape(64, 1, 1, 1, 20, 20, 128)
    noise = cabi.Noise(None, None, 1, 1, 0, None)
    outs = cabi.Outputs(*([1] * 9))
    parts = cabi.Parts(None, None, None, None, None, None)
    args = lambda **kw: (C.byref(kw.get("shape", shape)), kw.get("latent", 256), None, 256, 256, C.byref(noise), cabi.FLAG_PHILOX,
                         C.byref(kw.get("parts", parts)), C.byref(kw.get("outs", outs)), kw.get("ws", 256), 1 << 30, None)
    assert L.fvae_heads_parts(*args(latent=None)) == -1                                   # FVAE_ERR_NULL
    assert L.fvae_heads_parts(*args(shape=cabi.Shape(64, 1, 1, 1, 65, 20, 128))) == -3    # H > 64: FVAE_ERR_LIMIT
    assert L.fvae_heads_parts(*args(ws=257)) == -5                                        # workspace not 256-byte aligned
    assert L.fvae_heads_parts(*args(parts=cabi.Parts(256, None, None, None, None, None))) == -1   # z_mu without z_sigma
    assert L.fvae_heads_parts(*args(outs=cabi.Outputs(*([None] * 9)))) == -1              # the decoder outputs are required


def test_next_row_modules_have_no_cpu_fallback(cabi):
    """Resident panel, fused optimizer and metric fail loudly on CPU tensors / devices instead of falling back."""
    import numpy as np
    from factorvae_b200.metrics import rank_ic
    from factorvae_b200.optim import FlatAdam
    from factorvae_b200.panel import PanelIndex, ResidentPanel
    with pytest.raises(RuntimeError):
        FlatAdam(torch.zeros(16))
    with pytest.raises(RuntimeError):
        rank_ic(torch.zeros(4), torch.zeros(4), torch.tensor([0, 4], dtype=torch.int32))
    idx = PanelIndex(np.zeros((2, 2), np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32), np.array([0, 1]), 4)
    with pytest.raises(RuntimeError):
        ResidentPanel(np.zeros((4, 3), np.float32), idx, 2, "cpu")


def test_dropin_module_reexports_what_the_reference_module_leaks():
    """utils.py:6 does `from module import *` and uses pd / np / torch that the reference's module.py imports at its top
    (module.py:2-8, no __all__): the drop-in must leak the same names."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); ns = {}; exec('from module import *', ns); "
            "missing = [n for n in ('torch','nn','F','optim','DataLoader','Dataset','TensorDataset','pd','np','FactorVAE',"
            "'FeatureExtractor','FactorEncoder','FactorDecoder','FactorPredictor','AlphaLayer','BetaLayer','AttentionLayer') if n not in ns]; "
            "print(missing); sys.exit(1 if missing else 0)") % os.path.join(ROOT, "dropin")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, (r.stdout, r.stderr[-800:])
