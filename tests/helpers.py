"""Builders shared by the GPU tests: drop-in models loaded with the seeded synthetic weights."""
import torch

from avi_talking_b200 import synth


def build_wav2vec(precision, device="cuda", seed=0):
    from transformers import Wav2Vec2Config
    from avi_talking_b200.wav2vec import Wav2Vec2Model
    m = Wav2Vec2Model(Wav2Vec2Config())
    missing, unexpected = m.load_state_dict(synth.wav2vec2_state(seed), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    m.precision = precision
    return m.to(device).eval()


def build_faceformer(precision, fd=64, seed=74, device="cuda", w2v=None, variant="disentangle", head_std=1e-3):
    from avi_talking_b200.faceformer import Faceformer, FaceformerVert, make_args
    cls = Faceformer if variant == "disentangle" else FaceformerVert
    if w2v is None:
        w2v = build_wav2vec(precision, device="cpu")
    template = synth.flame_buffers()["v_template"].reshape(1, 1, 15069)
    m = cls(make_args(feature_dim=fd), audio_encoder=w2v, template=template)
    sd = synth.faceformer_state(fd=fd, seed=seed, head_std=head_std, variant=variant)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("audio_encoder.") or k in ("obj_embedding",) or k.startswith("PPE.") for k in missing), missing
    m.precision = precision
    m.audio_encoder.precision = precision
    return m.to(device).eval()


def build_flame(n_shape=100, device="cuda", mediapipe=True, tmpdir="/tmp/avi_flame_assets_test", precision="fp32"):
    from avi_talking_b200.flame import FLAME, FLAME_mediapipe
    cfg = synth.write_flame_assets(tmpdir)
    cfg.n_shape = n_shape
    m = (FLAME_mediapipe(cfg) if mediapipe else FLAME(cfg)).to(device)
    m.precision = precision
    return m
