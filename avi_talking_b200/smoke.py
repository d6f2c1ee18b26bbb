"""Builders of the drop-in models loaded with the seeded synthetic weights / assets (avi_talking_b200.synth): used by
__graft_entry__.smoke(), bench.py, the profiles/ scripts and the tests. Nothing here touches the oracle."""
from __future__ import annotations

import numpy as np
import torch


def build_models(precision: str, fd: int = 64, device: str = "cuda", seed_ff: int = 74, flame_dir: str = "/tmp/avi_flame_assets"):
    """Drop-in wav2vec2 + Faceformer + FLAME loaded with the seeded synthetic weights/assets (avi_talking_b200.synth)."""
    from transformers import Wav2Vec2Config

    from . import synth
    from .faceformer import Faceformer, make_args
    from .flame import FLAME_mediapipe
    from .wav2vec import Wav2Vec2Model

    w2v = Wav2Vec2Model(Wav2Vec2Config())
    w2v.load_state_dict(synth.wav2vec2_state(0), strict=False)
    flame = FLAME_mediapipe(synth.write_flame_assets(flame_dir))
    rng = np.random.default_rng(53)
    model = Faceformer(make_args(feature_dim=fd), audio_encoder=w2v, flame=flame,
                       coeff_mean=rng.normal(0, 0.3, size=53).astype("float32"),
                       coeff_std=(0.3 + rng.uniform(size=53)).astype("float32"))
    model.load_state_dict(synth.faceformer_state(fd=fd, seed=seed_ff), strict=False)
    model.precision = precision
    w2v.precision = precision
    flame.precision = precision
    return model.to(device).eval()


def build_prior(precision: str = "fp32", samples_per_cta: int = 0, device: str = "cuda", seed: int = 30):
    """Drop-in InstructDiffusionPrior (prior network + BrainNetwork) as constructed at train_diffusion_prior.py:961-991, loaded
    with the seeded synthetic weights."""
    from . import synth
    from .diffusion_prior import BrainNetwork, InstructDiffusionPrior, VersatileDiffusionPriorNetwork
    brain = BrainNetwork(in_dim=768, out_dim=128, clip_size=128, use_projector=True)
    net = VersatileDiffusionPriorNetwork(dim=128, depth=6, dim_head=64, heads=8, causal=False, num_tokens=1, learned_query_mode="pos_emb")
    prior = InstructDiffusionPrior(net=net, image_embed_dim=128, condition_on_text_encodings=False, timesteps=100, cond_drop_prob=0.2,
                                   image_embed_scale=None, voxel2clip=brain)
    prior.load_state_dict(synth.prior_state(seed), strict=False)
    brain.precision = precision
    prior.samples_per_cta = samples_per_cta
    return prior.to(device).eval()


def build_talking_head(precision: str = "fp32", device: str = "cuda", flame_dir: str = "/tmp/avi_flame_assets_emote"):
    """Drop-in TalkingHeadWrapper (EMOTE, Path B) with the seeded synthetic wav2vec2 / decoder weights and FLAME(300, 50)."""
    from transformers import Wav2Vec2Config

    from . import synth
    from .flame import FLAME
    from .talking_head import TalkingHeadWrapper, emote_cfg
    from .wav2vec import Wav2Vec2Model
    w2v = Wav2Vec2Model(Wav2Vec2Config())
    w2v.load_state_dict(synth.wav2vec2_state(0), strict=False)
    fcfg = synth.write_flame_assets(flame_dir)
    fcfg.n_shape, fcfg.n_exp = synth.EMOTE.n_shape, synth.EMOTE.n_exp
    flame = FLAME(fcfg)
    m = TalkingHeadWrapper.from_parts(w2v, flame, emote_cfg(n_identities=synth.EMOTE.n_identities))
    missing, unexpected = m.talking_head_model.load_state_dict(synth.emote_state(), strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("audio_model.") or ".flame." in k for k in missing), missing
    m.talking_head_model.precision = precision
    w2v.precision = precision
    flame.precision = precision
    return m.to(device).eval()
