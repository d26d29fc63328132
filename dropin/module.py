"""Put this directory first on sys.path and the reference's `main.py`, `train_model.py` and
`utils.py` (`from module import ...`, main.py:12, utils.py:6) pick up the B200-native modules.

`utils.py:6` does `from module import *` and then uses `pd`, `np`, `torch` ... that the reference's module.py happens to
import at its top (module.py:2-8) and -- having no `__all__` -- re-exports.  A drop-in has to re-export the same names, or
`utils.generate_prediction_scores` / `utils.RankIC` fail with NameError: this file therefore defines NO `__all__` and imports
those libraries under the reference's aliases."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

# the library aliases the reference's module.py leaks to `from module import *` (module.py:2-8)
import numpy as np  # noqa: F401,E402
import pandas as pd  # noqa: F401,E402
import torch  # noqa: F401,E402
import torch.nn as nn  # noqa: F401,E402
import torch.nn.functional as F  # noqa: F401,E402
import torch.optim as optim  # noqa: F401,E402
from torch.utils.data import DataLoader, Dataset, TensorDataset  # noqa: F401,E402

from factorvae_b200.module import (AlphaLayer, AttentionLayer, BetaLayer, FactorDecoder, FactorEncoder,  # noqa: F401,E402
                                   FactorPredictor, FactorVAE, FeatureExtractor, inject_noise, set_default_precision)
