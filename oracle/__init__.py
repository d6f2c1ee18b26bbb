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
  * INFERNO EMOTE decoder (emote_oracle.py): PINNED - the reference's own classes
    run through oracle/_inferno_import.py (pytorch_lightning / omegaconf / munch
    stubbed at import) -> tests/golden/emote.npz.
  * faceformer_vert training step (train_oracle.py): PINNED - the reference's own
    forward_switch_frame + loss.backward() + torch.optim.Adam -> tests/golden/train.npz
    (deterministic mode: dropout / SpecAugment / LayerDrop inactive).
  * CLIP text tower (clip_oracle.py): PINNED on transformers.CLIPTextModel, the class
    models/diffusion_prior.py:37 instantiates -> tests/golden/clip_text.npz.
  * FanEncoder image branch (fan_oracle.py): PINNED - equals the reference's own
    FanEncoder class (omegaconf YAML read stubbed) bit for bit -> tests/golden/fan.npz.
  * Audio front end (avi_talking_b200/frontend.py is host logic, checked directly):
    PINNED on the reference's own process_audio / create_base_sample compiled from
    source -> tests/golden/frontend.npz.
  * Diffusion prior (prior_oracle.py): the reference's own class sources are pinned
    (executed over oracle/dalle2_standin.py -> tests/golden/prior.npz); the stand-in
    for the un-vendored, un-pinned dalle2_pytorch / rotary_embedding_torch is a
    restatement of the published modules: PARITY UNPINNED for that part.
"""
