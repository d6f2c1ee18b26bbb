"""Builders shared by the GPU tests: drop-in models loaded with the seeded synthetic weights."""
import numpy as np
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


# ---- prior-training fixtures (tests/golden/prior_train_*.npz hold strided fingerprints of every gradient / updated parameter)
def fingerprint(t, n=2048):
    f = t.detach().reshape(-1).double()
    step = max(1, f.numel() // n)
    return f[::step][:n].float().numpy(), np.array([float(f.sum()), float(f.norm())])


def check_against_golden(g, got_grads, got_new, grad_tol, param_tol, names=None):
    """got_* : {state-dict key: tensor}. Gradients are compared relative to the tensor's own l2 norm per element count."""
    worst_g = worst_p = 0.0
    for n in (names if names is not None else [str(x) for x in g["names"]]):
        gs, gn = g["g:" + n], g["gs:" + n]
        s, sn = fingerprint(got_grads[n])
        scale = max(float(np.abs(gs).max()), float(gn[1]) / max(1.0, got_grads[n].numel()) ** 0.5, 1e-12)
        worst_g = max(worst_g, float(np.abs(s - gs).max()) / scale, abs(sn[1] - gn[1]) / max(gn[1], 1e-12))
        ps = g["p:" + n]
        worst_p = max(worst_p, float(np.abs(fingerprint(got_new[n])[0] - ps).max()))
    assert worst_g < grad_tol, f"worst relative gradient error {worst_g}"
    assert worst_p < param_tol, f"worst absolute error of an AdamW-updated parameter {worst_p}"
    return worst_g, worst_p


