"""CPU oracle for the audio -> FLAME-vertex hot path of sunyasheng/AVI-Talking.

TEST INFRASTRUCTURE ONLY.  Nothing under ``avi_talking_b200/`` imports this
package.  The only permitted users are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, and there
only as the checker or as the timed CPU reference arm - never as the thing
shipped.

The reference is pure Python/PyTorch, so the oracle is a plain fp32 torch/numpy
restatement of the reference's arithmetic, function by function, each citing the
reference file:line it follows (paths relative to the upstream repository root).

Pinning status (see DESIGN.md "Oracle"):
  * FLAME / lbs, wav2vec2 wrapper, bias masks, PPE, Faceformer.predict/forward_ff:
    PINNED - ``oracle/make_golden.py`` imports the reference's own classes from
    /root/reference in the build container and commits their outputs on seeded
    inputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the
    restatement against them.
  * dalle2_pytorch-based diffusion prior sampler and the INFERNO EMOTE decoder
    (un-importable here: dalle2_pytorch / pytorch_lightning / omegaconf are not
    vendored nor installed): PARITY UNPINNED - restated from the reference's
    call sites and the published algorithm.
"""
