import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_golden(name):
    """Returns dict(params, grads, out, pred, inputs, dims) of torch tensors for one fixture."""
    z = np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz"))
    g = dict(params={}, grads={}, out={}, pred={}, inp={})
    for k in z.files:
        kind, key = k.split(":", 1)
        t = torch.from_numpy(z[k])
        if kind == "param":
            g["params"][key] = t
        elif kind == "grad":
            g["grads"][key] = t
        elif kind == "out":
            g["out"][key] = t
        elif kind == "pred":
            g["pred"][key] = t
        elif kind == "in":
            g["inp"][key] = t
        elif kind == "meta":
            C, H, K, M, T, train = [int(v) for v in z[k]]
            g["dims"] = dict(C=C, H=H, K=K, M=M, T=T, train=bool(train))
    return g


def split_by_date(g):
    """Per-date python lists (xs, ys, epss, masks) from a loaded fixture."""
    ptr = g["inp"]["date_ptr"].tolist()
    xs, ys, epss, masks = [], [], [], []
    for d in range(len(ptr) - 1):
        a, b = ptr[d], ptr[d + 1]
        xs.append(g["inp"]["x"][a:b])
        ys.append(g["inp"]["y"][a:b])
        epss.append(g["inp"]["eps"][a:b])
        masks.append(g["inp"]["keep_mask"][:, a:b] if "keep_mask" in g["inp"] else None)
    if "keep_mask" not in g["inp"]:
        masks = None
    return xs, ys, epss, masks


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
