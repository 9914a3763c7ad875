"""LoRA variant of the sampler — drop-in for ``fusion_generation/fusion_sampling_lora.py``.

The reference file differs from ``fusion_sampling.py`` in 11 hunks: concept U-Nets
carry ``LoRAAttnProcessor_base`` layers (``:203-210``), the fused window is
``t_stop_cur <= t <= t_cond_cur`` (``:324,378``) while the hook window is ``timesteps[t_cond:t_stop]``
(``:477``), and ``--t_stop`` exists (``:547``).  All of that is ``Tweediemix(variant="lora")``.
"""
from __future__ import annotations

from .fusion_sampling import FusionComponents, Tweediemix as _Tweediemix, build_parser as _build_parser, main as _main


class Tweediemix(_Tweediemix):
    def __init__(self, config, components=None, **kw):
        kw.setdefault("variant", "lora")
        super().__init__(config, components, **kw)


def build_parser():
    return _build_parser(lora=True)


def main(argv=None):
    return _main(argv, lora=True)


__all__ = ["Tweediemix", "FusionComponents", "build_parser", "main"]

if __name__ == "__main__":
    main()
