"""tweediemix_b200 — B200-native (sm_100a) implementation of the TweedieMix fusion-sampling hot path.

Host side mirrors the reference's own interface for this path (``fusion_generation/``):
``utils_custom`` / ``utils_lora`` (hook API), ``fusion_sampling`` / ``fusion_sampling_lora``
(``Tweediemix`` + CLI); the arithmetic lives in ``libtmx.so`` (``csrc/``, C ABI in ``include/tmx.h``).
"""
__version__ = "0.1.0"
