"""CPU oracle for the TweedieMix fusion-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tweediemix_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the timed CPU baseline, never as the shipped path.

Parity pinning:
  * the reference ships no tests and no golden vectors.  Its sampler
    files do not import as they are (diffusers / sentence_transformers absent), but
    with those imports stubbed the reference's UNMODIFIED ``Tweediemix.init_fusion``,
    ``alpha`` and ``denoise_step`` run on CPU: ``tests/golden/make_golden_sampler.py``
    records their latents step by step (custom 5-step config 1, LoRA 10-step with
    ``t_stop``, one step per phase) and ``tests/test_oracle_vs_reference_sampler.py``
    holds ``sampler_ref`` / ``step_math`` / ``schedule`` / ``masks_ref`` to them at
    2e-5 relative: **pinned**;
  * only ``unet_ref`` (the diffusers U-Net body, [D]) remains **parity unpinned**:
    diffusers cannot be installed here, and the reference treats it as a black box;
  * the hook restatement (``hooks_ref``) IS pinned: the reference's own
    ``fusion_generation/utils_custom.py`` and ``utils_lora.py`` import and run
    in the build container, and ``tests/golden/make_golden.py`` records their
    outputs on seeded inputs into ``tests/golden/*.pt``; ``tests/test_oracle_hooks.py``
    replays them bit-for-bit against ``hooks_ref``.

Every function cites the reference file:line it follows (paths relative to the
reference checkout, e.g. ``fusion_generation/fusion_sampling.py:305-307``).
Lines marked [D] restate diffusers==0.29.2 (``requirements.txt:4``), which is
not vendored in the reference and not installed in this image.
"""
