"""Put this directory first on sys.path and the reference's `main.py`, `train_model.py` and
`utils.py` (`from module import ...`, main.py:12, utils.py:6) pick up the B200-native modules."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from factorvae_b200.module import *  # noqa: F401,F403,E402
from factorvae_b200.module import (AlphaLayer, AttentionLayer, BetaLayer, FactorDecoder, FactorEncoder,  # noqa: F401,E402
                                   FactorPredictor, FactorVAE, FeatureExtractor)
