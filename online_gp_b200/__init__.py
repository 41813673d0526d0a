"""online_gp_b200 — B200-native (sm_100a) implementation of the WISKI online-update hot path of wjmaddox/online_gp.

Host layer mirrors the reference's Python surface (``online_gp.models``, ``online_gp.lazy``, ``online_gp.mlls``,
``online_gp.likelihoods``, ``online_gp.settings``); all arithmetic runs in ``csrc/libwiski_b200.so`` through the C
ABI of ``include/wiski_b200.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
