"""dahitra_b200 — B200-native (sm_100a) implementation of DAHiTra's ``newUNetTrans`` bitemporal forward pass.

Public surface mirrors the reference for this path:
    from dahitra_b200.networks import define_G, BASE_Transformer_UNet
"""
from .networks import BASE_Transformer_UNet, define_G, init_net, init_weights  # noqa: F401

__version__ = "0.1.0"
