"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements (numpy / torch-CPU fp32, plus one C file) of the reference
algorithms on the hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker or the CPU baseline --
never as a product code path.  `isegprobe_b200/` must not import it.

Pinning status (see DESIGN.md "Oracle"):
  * distmaps, loftup, lift, head, vit, patch_embed: pinned against the
    reference's own modules executed in the authoring container
    (`oracle/make_golden.py` -> `tests/golden/*.npz`).
  * jbu (FeatUp JBUStack / AdaptiveConv): the arithmetic lives in the
    un-vendored, un-pinned third-party repo mhamilton723/FeatUp
    (reference call site core/model/upsamplers/JBUFeatUp.py:30-32).
    PARITY UNPINNED: restated from the published algorithm; there is no
    reference-side vector to check it against.
"""
