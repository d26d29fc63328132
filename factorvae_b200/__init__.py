"""factorvae_b200 -- B200-native (sm_100a) ELBO-step hot path for FactorVAE.

Public surface:
  * `factorvae_b200.module`  : drop-in for the reference's module.py (same classes / state_dict)
  * `factorvae_b200.engine`  : flat-parameter layout + date-batched forward / backward over the C ABI
  * `factorvae_b200.batched` : date-batched, date-sharded (multi-GPU) training step
  * `include/fvae_b200.h`    : the C ABI itself (libfvae_b200.so)
"""
from .module import (AlphaLayer, AttentionLayer, BetaLayer, FactorDecoder, FactorEncoder, FactorPredictor, FactorVAE,
                     FeatureExtractor, inject_noise, set_default_precision)

__all__ = ["FeatureExtractor", "FactorEncoder", "AlphaLayer", "BetaLayer", "FactorDecoder", "AttentionLayer",
           "FactorPredictor", "FactorVAE", "inject_noise", "set_default_precision"]
