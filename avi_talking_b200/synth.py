"""Seeded synthetic assets, weights and inputs (shared by bench.py, smoke(), the tests and the oracle).

Pure data generation (numpy/torch on the host), no arithmetic of the hot path lives here.

Everything is drawn from ``numpy.random.default_rng`` (PCG64 + ziggurat normal),
which is bit-reproducible across machines, so the build container (where the
golden fixtures are minted from the real reference) and the GPU box (where the
CUDA path is checked) see identical tensors without shipping 377 MB of weights.

Shapes/keys follow the reference:
  FLAME pickle keys and slicing ........ third_party/inferno/inferno/models/DecaFLAME.py:53-106
  wav2vec2-base architecture ........... transformers Wav2Vec2Config() defaults (models/lib/wav2vec.py:76-78)
  Faceformer heads / decoder ........... models/faceformer_disentangle.py:158-241,327
"""
from __future__ import annotations

import math
import pickle
from types import SimpleNamespace

import numpy as np
import torch

N_VERT = 5023
N_FACE = 9976
N_JOINT = 5


def _t(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


# --------------------------------------------------------------------------- FLAME assets
def flame_model_dict(seed: int = 0) -> dict:
    """Synthetic stand-in for FLAME ``generic_model.pkl`` (licensed, absent upstream).

    Same key set / shapes / dtypes the reference unpickles (DecaFLAME.py:53-74).
    Magnitudes: template ~8 cm, blendshape directions ~1 mm per unit coefficient.
    """
    rng = np.random.default_rng(seed)
    v_template = rng.normal(0.0, 0.08, size=(N_VERT, 3))
    shapedirs = rng.normal(0.0, 1e-3, size=(N_VERT, 3, 400))
    posedirs = rng.normal(0.0, 1e-3, size=(N_VERT, 3, 36))
    jr = np.abs(rng.normal(size=(N_JOINT, N_VERT)))
    jr /= jr.sum(1, keepdims=True)
    w = np.abs(rng.normal(size=(N_VERT, N_JOINT))) ** 4  # peaky, like real skinning weights
    w /= w.sum(1, keepdims=True)
    f = rng.integers(0, N_VERT, size=(N_FACE, 3)).astype(np.uint32)
    kintree = np.array([[4294967295, 0, 1, 1, 1], [0, 1, 2, 3, 4]], dtype=np.int64)
    return dict(f=f, v_template=v_template, shapedirs=shapedirs, posedirs=posedirs,
                J_regressor=jr, kintree_table=kintree, weights=w)


def flame_lmk_dict(seed: int = 1) -> dict:
    """Synthetic ``landmark_embedding.npy`` payload (DecaFLAME.py:87-98)."""
    rng = np.random.default_rng(seed)

    def bary(*shape):
        b = np.abs(rng.normal(size=shape + (3,))) + 0.05
        return (b / b.sum(-1, keepdims=True)).astype(np.float32)

    return dict(
        static_lmk_faces_idx=rng.integers(0, N_FACE, size=(51,)).astype(np.int64),
        static_lmk_bary_coords=bary(51),
        dynamic_lmk_faces_idx=rng.integers(0, N_FACE, size=(79, 17)).astype(np.int64),
        dynamic_lmk_bary_coords=bary(79, 17),
        full_lmk_faces_idx=rng.integers(0, N_FACE, size=(1, 68)).astype(np.int64),
        full_lmk_bary_coords=bary(1, 68),
    )


def flame_mediapipe_lmk_dict(seed: int = 2) -> dict:
    """Synthetic mediapipe embedding (DecaFLAME.py:276-283): 105 static landmarks."""
    rng = np.random.default_rng(seed)
    b = np.abs(rng.normal(size=(105, 3))) + 0.05
    return dict(lmk_face_idx=rng.integers(0, N_FACE, size=(105,)).astype(np.int64),
                lmk_b_coords=(b / b.sum(-1, keepdims=True)).astype(np.float32),
                landmark_indices=np.arange(105))


def write_flame_assets(dirpath: str, seed: int = 0) -> SimpleNamespace:
    """Write the three asset files and return a config like ``misc/flame_cfg.pkl`` would give."""
    import os
    os.makedirs(dirpath, exist_ok=True)
    p_model = os.path.join(dirpath, "generic_model.pkl")
    p_lmk = os.path.join(dirpath, "landmark_embedding.npy")
    p_mp = os.path.join(dirpath, "mediapipe_landmark_embedding.npz")
    with open(p_model, "wb") as fh:
        pickle.dump(flame_model_dict(seed), fh)
    np.save(p_lmk, flame_lmk_dict(seed + 1), allow_pickle=True)
    np.savez(p_mp, **flame_mediapipe_lmk_dict(seed + 2))
    return SimpleNamespace(flame_model_path=p_model, flame_lmk_embedding_path=p_lmk,
                           flame_mediapipe_lmk_embedding_path=p_mp, n_shape=100, n_exp=50)


def flame_buffers(n_shape: int = 100, n_exp: int = 50, seed: int = 0) -> dict:
    """The registered buffers FLAME.__init__ derives from the pickle (DecaFLAME.py:60-106)."""
    m = flame_model_dict(seed)
    lm = flame_lmk_dict(seed + 1)
    mp = flame_mediapipe_lmk_dict(seed + 2)
    sd = np.concatenate([m["shapedirs"][:, :, :n_shape], m["shapedirs"][:, :, 300:300 + n_exp]], 2)
    pd = np.reshape(m["posedirs"], [-1, 36]).T
    parents = torch.tensor([-1, 0, 1, 1, 1], dtype=torch.long)
    return dict(
        faces_tensor=torch.from_numpy(m["f"].astype(np.int64)),
        v_template=_t(m["v_template"]), shapedirs=_t(sd), posedirs=_t(pd),
        J_regressor=_t(m["J_regressor"]), parents=parents, lbs_weights=_t(m["weights"]),
        lmk_faces_idx=torch.from_numpy(lm["static_lmk_faces_idx"]),
        lmk_bary_coords=_t(lm["static_lmk_bary_coords"]),
        dynamic_lmk_faces_idx=torch.from_numpy(lm["dynamic_lmk_faces_idx"]),
        dynamic_lmk_bary_coords=_t(lm["dynamic_lmk_bary_coords"]),
        full_lmk_faces_idx=torch.from_numpy(lm["full_lmk_faces_idx"]),
        full_lmk_bary_coords=_t(lm["full_lmk_bary_coords"]),
        lmk_faces_idx_mediapipe=torch.from_numpy(mp["lmk_face_idx"]),
        lmk_bary_coords_mediapipe=_t(mp["lmk_b_coords"]),
        neck_kin_chain=torch.tensor([1, 0], dtype=torch.long),
        eye_pose=torch.zeros(1, 6), neck_pose=torch.zeros(1, 3),
    )


def flame_params(n_frames: int, n_shape: int = 100, n_exp: int = 50, seed: int = 3,
                 zero_shape: bool = False, global_pose: bool = True) -> dict:
    """exp ~ N(0,1), jaw ~ N(0,0.1^2) rad, small global rotation."""
    rng = np.random.default_rng(seed)
    shape = np.zeros((n_frames, n_shape)) if zero_shape else rng.normal(size=(n_frames, n_shape))
    exp = rng.normal(size=(n_frames, n_exp))
    pose = np.zeros((n_frames, 6))
    if global_pose:
        pose[:, :3] = rng.normal(0, 0.2, size=(n_frames, 3))
    pose[:, 3:] = rng.normal(0, 0.1, size=(n_frames, 3))
    eye = rng.normal(0, 0.1, size=(n_frames, 6))
    return dict(shape=_t(shape), exp=_t(exp), pose=_t(pose), eye=_t(eye))


# --------------------------------------------------------------------------- audio
def audio(n_clips: int, n_samples: int, seed: int = 1234) -> torch.Tensor:
    """z-normalised noise, as Wav2Vec2Processor leaves it (dataset/voca_data_loader.py:59-60)."""
    out = np.empty((n_clips, n_samples), dtype=np.float32)
    for c in range(n_clips):
        x = np.random.default_rng(seed + c).normal(size=n_samples)
        # mild low-pass so neighbouring samples correlate like speech does
        x = np.convolve(x, np.array([0.25, 0.5, 0.25]), mode="same")
        out[c] = (x - x.mean()) / np.sqrt(x.var() + 1e-7)
    return torch.from_numpy(out)


# --------------------------------------------------------------------------- wav2vec2 weights
W2V = SimpleNamespace(conv_dim=(512,) * 7, conv_kernel=(10, 3, 3, 3, 3, 2, 2), conv_stride=(5, 2, 2, 2, 2, 2, 2),
                      hidden=768, heads=12, ffn=3072, layers=12, pos_k=128, pos_groups=16, eps=1e-5)


def wav2vec2_state(seed: int = 0, layers: int = 12) -> dict:
    """State dict with transformers-5.x key names for Wav2Vec2Model(Wav2Vec2Config()).

    Not the HF init (zero biases / std-0.02 Linears give near-uniform attention and
    leave bias paths untested): Linears ~ N(0, 0.7/sqrt(fan_in)), small non-zero biases,
    LayerNorm gains 1 +- 0.1, convs Kaiming-normal.
    """
    rng = np.random.default_rng(seed)
    sd = {}

    def lin(name, out_f, in_f, gain=0.7):
        sd[name + ".weight"] = _t(rng.normal(0, gain / math.sqrt(in_f), size=(out_f, in_f)))
        sd[name + ".bias"] = _t(rng.normal(0, 0.02, size=(out_f,)))

    def ln(name, c):
        sd[name + ".weight"] = _t(1.0 + 0.1 * rng.normal(size=(c,)))
        sd[name + ".bias"] = _t(0.05 * rng.normal(size=(c,)))

    cin = 1
    for i, (c, k) in enumerate(zip(W2V.conv_dim, W2V.conv_kernel)):
        sd[f"feature_extractor.conv_layers.{i}.conv.weight"] = _t(
            rng.normal(0, math.sqrt(2.0 / (cin * k)), size=(c, cin, k)))
        cin = c
    ln("feature_extractor.conv_layers.0.layer_norm", 512)
    ln("feature_projection.layer_norm", 512)
    lin("feature_projection.projection", 768, 512)
    sd["masked_spec_embed"] = _t(rng.uniform(size=(768,)))
    sd["encoder.pos_conv_embed.conv.bias"] = _t(rng.normal(0, 0.02, size=(768,)))
    v = rng.normal(0, 2 * math.sqrt(1.0 / (128 * 48)), size=(768, 48, 128))
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = _t(v)
    g = np.sqrt((v ** 2).sum(axis=(0, 1), keepdims=True)) * (1.0 + 0.1 * rng.normal(size=(1, 1, 128)))
    sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = _t(g)
    ln("encoder.layer_norm", 768)
    for l in range(layers):
        p = f"encoder.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(p + "attention." + nm, 768, 768, gain=1.0 if nm in ("q_proj", "k_proj") else 0.7)
        ln(p + "layer_norm", 768)
        lin(p + "feed_forward.intermediate_dense", 3072, 768)
        lin(p + "feed_forward.output_dense", 768, 3072)
        ln(p + "final_layer_norm", 768)
    return sd


# --------------------------------------------------------------------------- Faceformer (Path A) weights
def faceformer_state(fd: int = 64, n_subjects: int = 8, vertice_dim: int = N_VERT * 3, seed: int = 10,
                     head_std: float = 1e-3, variant: str = "disentangle") -> dict:
    """Heads + 1-layer nn.TransformerDecoder, torch key names (faceformer_disentangle.py:171-202,241,327).

    ``vertice_map_r`` is zero-initialised upstream (:201-202) which would make parity vacuous;
    it is re-randomised here with ``head_std`` (1e-3 -> ~8 mm rms displacement at fd=64,
    the magnitude of real facial motion).
    ``variant='vert'`` follows models/faceformer_vert.py:162 (dim_feedforward = 2*fd, no merge layer).
    """
    rng = np.random.default_rng(seed)
    sd = {}

    def lin(name, out_f, in_f, bias=True, std=None):
        s = std if std is not None else 0.7 / math.sqrt(in_f)
        sd[name + ".weight"] = _t(rng.normal(0, s, size=(out_f, in_f)))
        if bias:
            sd[name + ".bias"] = _t(rng.normal(0, 0.02, size=(out_f,)))

    def ln(name, c):
        sd[name + ".weight"] = _t(1.0 + 0.1 * rng.normal(size=(c,)))
        sd[name + ".bias"] = _t(0.05 * rng.normal(size=(c,)))

    lin("audio_feature_map", fd, 768)
    # vertice_map sees displacements of ~head_std*sqrt(fd): scale so the fed-back embedding is O(0.3)
    lin("vertice_map", fd, vertice_dim, std=0.3 / (head_std * math.sqrt(fd) * math.sqrt(vertice_dim)))
    lin("vertice_map_r", vertice_dim, fd, std=head_std)
    sd["vertice_map_r.bias"] = _t(rng.normal(0, head_std, size=(vertice_dim,)))
    lin("obj_vector", fd, n_subjects, bias=False, std=0.35)
    if variant == "disentangle":
        lin("v_merge2hidden", fd, 36 + fd)
        sd["learnable_eye_embed"] = _t(rng.normal(0, 0.3, size=(1, 1, 6)))
    p = "transformer_decoder.layers.0."
    sd[p + "self_attn.in_proj_weight"] = _t(rng.normal(0, 1.0 / math.sqrt(fd), size=(3 * fd, fd)))
    sd[p + "self_attn.in_proj_bias"] = _t(rng.normal(0, 0.02, size=(3 * fd,)))
    lin(p + "self_attn.out_proj", fd, fd)
    sd[p + "multihead_attn.in_proj_weight"] = _t(rng.normal(0, 1.0 / math.sqrt(fd), size=(3 * fd, fd)))
    sd[p + "multihead_attn.in_proj_bias"] = _t(rng.normal(0, 0.02, size=(3 * fd,)))
    lin(p + "multihead_attn.out_proj", fd, fd)
    lin(p + "linear1", 2 * fd, fd)
    lin(p + "linear2", fd, 2 * fd)
    for i in (1, 2, 3):
        ln(p + f"norm{i}", fd)
    return sd


def fan_embeddings(n_frames: int, seed: int = 20):
    """Deterministic stand-in for the FanEncoder image branch (OUT of scope):
    the 4-tuple (head[6], eye[6], emo[30], None) per frame that predict() consumes (:783-797)."""
    rng = np.random.default_rng(seed)
    return dict(head=_t(np.zeros((n_frames, 6))), eye=_t(np.zeros((n_frames, 6))),
                emo=_t(rng.normal(size=(n_frames, 30))))


# --------------------------------------------------------------------------- diffusion prior (text -> style) weights
def prior_state(seed: int = 30, dim: int = 128, depth: int = 6, heads: int = 8, dim_head: int = 64, ff_mult: int = 4,
                brain_in: int = 768, brain_h: int = 4096, brain_blocks: int = 4, with_brain: bool = True) -> dict:
    """State dict with the key names InstructDiffusionPrior(net=VersatileDiffusionPriorNetwork(...), voxel2clip=BrainNetwork(...))
    registers (models/diffusion_prior.py:58-117,119-313; construction train_diffusion_prior.py:963-991):
    ``net.*`` (dalle2_pytorch Attention / FeedForward / RelPosBias / LayerNorm / MLP parameter names) and ``voxel2clip.*``."""
    rng = np.random.default_rng(seed)
    sd = {}

    def w(name, out_f, in_f, gain=1.0):
        sd[name] = _t(rng.normal(0, gain / math.sqrt(in_f), size=(out_f, in_f)))

    def g(name, c):
        sd[name] = _t(1.0 + 0.1 * rng.normal(size=(c,)))

    inner = dim_head * heads
    ff_inner = ff_mult * dim
    t = "net.to_time_embeds.0.1.net."
    w(t + "0.0.weight", 2 * dim, dim); sd[t + "0.0.bias"] = _t(rng.normal(0, 0.02, size=(2 * dim,)))
    w(t + "1.0.weight", 2 * dim, 2 * dim); sd[t + "1.0.bias"] = _t(rng.normal(0, 0.02, size=(2 * dim,)))
    w(t + "2.weight", dim, 2 * dim); sd[t + "2.bias"] = _t(rng.normal(0, 0.02, size=(dim,)))
    sd["net.learned_query"] = _t(rng.normal(size=(1, dim)) * dim ** -0.5)
    sd["net.null_brain_embeds"] = _t(rng.normal(size=(1, dim)))
    sd["net.null_image_embed"] = _t(rng.normal(size=(1, dim)))
    c = "net.causal_transformer."
    sd[c + "rel_pos_bias.relative_attention_bias.weight"] = _t(rng.normal(size=(32, heads)))
    freqs = 1.0 / (10000.0 ** (np.arange(0, 32, 2)[:16].astype(np.float32) / 32.0))
    for l in range(depth):
        a = c + f"layers.{l}.0."
        g(a + "norm.g", dim)
        sd[a + "null_kv"] = _t(rng.normal(size=(2, dim_head)))
        w(a + "to_q.weight", inner, dim)
        w(a + "to_kv.weight", 2 * dim_head, dim)
        w(a + "to_out.0.weight", dim, inner)
        g(a + "to_out.1.g", dim)
        sd[a + "rotary_emb.freqs"] = _t(freqs)
        f = c + f"layers.{l}.1."
        g(f + "0.g", dim)
        w(f + "1.weight", 2 * ff_inner, dim)
        w(f + "5.weight", dim, ff_inner)
    g(c + "norm.g", dim)
    w(c + "project_out.weight", dim, dim)
    if with_brain:
        b = "voxel2clip."

        def lin(name, out_f, in_f):
            w(name + ".weight", out_f, in_f)
            sd[name + ".bias"] = _t(rng.normal(0, 0.02, size=(out_f,)))

        def ln(name, cdim):
            sd[name + ".weight"] = _t(1.0 + 0.1 * rng.normal(size=(cdim,)))
            sd[name + ".bias"] = _t(0.05 * rng.normal(size=(cdim,)))

        lin(b + "lin0.0", brain_h, brain_in); ln(b + "lin0.1", brain_h)
        for i in range(brain_blocks):
            lin(b + f"mlp.{i}.0", brain_h, brain_h); ln(b + f"mlp.{i}.1", brain_h)
        lin(b + "lin1", dim, brain_h)
        ln(b + "projector.0", dim); lin(b + "projector.2", 2048, dim)
        ln(b + "projector.3", 2048); lin(b + "projector.5", 2048, 2048)
        ln(b + "projector.6", 2048); lin(b + "projector.8", dim, 2048)
    return sd


def prior_inputs(batch: int, steps: int, seed: int = 7, dim: int = 128) -> dict:
    """voxel [B,768] (stand-in for mean-pooled CLIP-L token embeddings), initial image_embed [B,1,128], per-step noise."""
    rng = np.random.default_rng(seed)
    return dict(voxel=_t(rng.normal(size=(batch, 768))), image_embed=_t(rng.normal(size=(batch, 1, dim))),
                noises=_t(rng.normal(size=(steps, batch, 1, dim))))


# --------------------------------------------------------------------------- EMOTE (Path B) decoder weights and sample dict
EMOTE = SimpleNamespace(feature_dim=128, nhead=8, bottleneck=256, latent_frame=8, quant_factor=3, l2l_ff=384, n_out=53,
                        n_expression=8, n_intensities=3, n_identities=32, n_shape=300, n_exp=50)


def emote_state(seed: int = 40) -> dict:
    """Weights of the EMOTE talking-head model below wav2vec2, with the parameter names the reference's module tree registers
    (TalkingHeadBase: sequence_encoder / sequence_decoder; BertPriorDecoder FaceFormerDecoder.py:987-1075; L2lDecoder
    L2lMotionPrior.py:361-455). The reference zero-initialises ``decoder`` (:1050-1051); it is re-randomised here."""
    rng = np.random.default_rng(seed)
    E = EMOTE
    sd = {}

    def lin(name, out_f, in_f, gain=0.7):
        sd[name + ".weight"] = _t(rng.normal(0, gain / math.sqrt(in_f), size=(out_f, in_f)))
        sd[name + ".bias"] = _t(rng.normal(0, 0.02, size=(out_f,)))

    def ln(name, c):
        sd[name + ".weight"] = _t(1.0 + 0.1 * rng.normal(size=(c,)))
        sd[name + ".bias"] = _t(0.05 * rng.normal(size=(c,)))

    def enc_layer(p, d, ff):
        sd[p + "self_attn.in_proj_weight"] = _t(rng.normal(0, 1.0 / math.sqrt(d), size=(3 * d, d)))
        sd[p + "self_attn.in_proj_bias"] = _t(rng.normal(0, 0.02, size=(3 * d,)))
        lin(p + "self_attn.out_proj", d, d)
        lin(p + "linear1", ff, d)
        lin(p + "linear2", d, ff)
        ln(p + "norm1", d)
        ln(p + "norm2", d)

    lin("sequence_encoder.linear", E.feature_dim, 768)
    cond = E.n_expression + E.n_intensities + E.n_identities
    lin("sequence_decoder.obj_vector.map", E.feature_dim, cond, gain=1.0)
    enc_layer("sequence_decoder.bert_decoder.layers.0.", E.feature_dim, E.feature_dim)
    lin("sequence_decoder.decoder", E.bottleneck, E.feature_dim)
    lin("sequence_decoder.squasher_2.linear", E.bottleneck, E.bottleneck * E.latent_frame)
    m = "sequence_decoder.motion_prior.motion_decoder."
    for i in range(E.quant_factor):
        shape = (E.bottleneck, E.bottleneck, 5)   # ConvTranspose1d weight is [in, out, k]; Conv1d [out, in, k] - both square here
        sd[m + f"expander.{i}.0.weight"] = _t(rng.normal(0, 1.0 / math.sqrt(E.bottleneck * (2.5 if i == 0 else 5)), size=shape))
        sd[m + f"expander.{i}.0.bias"] = _t(rng.normal(0, 0.02, size=(E.bottleneck,)))
        sd[m + f"expander.{i}.2.weight"] = _t(1.0 + 0.1 * rng.normal(size=(E.bottleneck,)))
        sd[m + f"expander.{i}.2.bias"] = _t(0.05 * rng.normal(size=(E.bottleneck,)))
        sd[m + f"expander.{i}.2.running_mean"] = _t(0.1 * rng.normal(size=(E.bottleneck,)))
        sd[m + f"expander.{i}.2.running_var"] = _t(0.5 + rng.uniform(size=(E.bottleneck,)))
        sd[m + f"expander.{i}.2.num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
    lin(m + "decoder_linear_embedding", E.bottleneck, E.bottleneck)
    enc_layer(m + "decoder_transformer.layers.0.", E.bottleneck, E.l2l_ff)
    sd[m + "cross_smooth_layer.weight"] = _t(rng.normal(0, 0.25 / math.sqrt(E.bottleneck * 5), size=(E.n_out, E.bottleneck, 5)))
    sd[m + "cross_smooth_layer.bias"] = _t(rng.normal(0, 0.02, size=(E.n_out,)))
    return sd


def emote_sample(batch: int, n_frames: int, seed: int = 50) -> dict:
    """The sample dict TalkingHeadWrapper.forward consumes (evaluation_functions.py:141-161,218-275): raw_audio [B,T,640] (int16 view
    of the waveform as float), samplerate, gt_shape [B,300], gt_exp [B,T,50], gt_jaw [B,T,3], and the three per-frame condition
    one-hots (expression / intensity / identity)."""
    rng = np.random.default_rng(seed)
    E = EMOTE
    wav = audio(batch, n_frames * 640, seed=seed + 1)
    raw = torch.round(wav * 3000.0).clamp(-32767, 32767).reshape(batch, n_frames, 640)

    def one_hot(n):
        idx = rng.integers(0, n, size=(batch,))
        return torch.nn.functional.one_hot(torch.from_numpy(idx), n).float()[:, None, :].expand(batch, n_frames, n).contiguous()

    return dict(raw_audio=raw, samplerate=[16000] * batch,
                gt_shape=_t(rng.normal(size=(batch, E.n_shape))), gt_exp=_t(rng.normal(size=(batch, n_frames, E.n_exp))),
                gt_jaw=_t(0.1 * rng.normal(size=(batch, n_frames, 3))),
                gt_expression_label_condition=one_hot(E.n_expression), gt_expression_intensity_condition=one_hot(E.n_intensities),
                gt_expression_identity_condition=one_hot(E.n_identities))


# --------------------------------------------------------------------------- CLIP-L text tower
CLIP_TEXT = SimpleNamespace(vocab=49408, hidden=768, heads=12, ffn=3072, layers=12, max_pos=77, eps=1e-5)


def clip_text_state(seed: int = 60, layers: int = 12, vocab: int = CLIP_TEXT.vocab) -> dict:
    """State dict with transformers' CLIPTextModel key names (text_model.*), seeded; same spirit as wav2vec2_state (non-trivial
    biases / LayerNorm gains so every path is exercised)."""
    rng = np.random.default_rng(seed)
    sd = {}
    C, Fd = CLIP_TEXT.hidden, CLIP_TEXT.ffn
    sd["text_model.embeddings.token_embedding.weight"] = _t(rng.normal(0, 0.5, size=(vocab, C)))
    sd["text_model.embeddings.position_embedding.weight"] = _t(rng.normal(0, 0.3, size=(CLIP_TEXT.max_pos, C)))

    def lin(name, out_f, in_f, gain=0.7):
        sd[name + ".weight"] = _t(rng.normal(0, gain / math.sqrt(in_f), size=(out_f, in_f)))
        sd[name + ".bias"] = _t(rng.normal(0, 0.02, size=(out_f,)))

    def ln(name, c):
        sd[name + ".weight"] = _t(1.0 + 0.1 * rng.normal(size=(c,)))
        sd[name + ".bias"] = _t(0.05 * rng.normal(size=(c,)))

    for l in range(layers):
        p = f"text_model.encoder.layers.{l}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(p + "self_attn." + nm, C, C, gain=1.0 if nm in ("q_proj", "k_proj") else 0.7)
        ln(p + "layer_norm1", C)
        lin(p + "mlp.fc1", Fd, C)
        lin(p + "mlp.fc2", C, Fd)
        ln(p + "layer_norm2", C)
    ln("text_model.final_layer_norm", C)
    return sd


def clip_tokens(batch: int, seed: int = 61, vocab: int = CLIP_TEXT.vocab) -> torch.Tensor:
    """Synthetic token ids [B, 77] shaped like CLIPTokenizer output: BOS, a few words, EOS, padding (= EOS id)."""
    rng = np.random.default_rng(seed)
    ids = np.full((batch, CLIP_TEXT.max_pos), vocab - 1, dtype=np.int64)
    ids[:, 0] = vocab - 2
    for b in range(batch):
        n = int(rng.integers(3, 40))
        ids[b, 1:1 + n] = rng.integers(0, vocab - 2, size=n)
    return torch.from_numpy(ids)


# --------------------------------------------------------------------------- FanEncoder image branch
def fan_state(seed: int = 80) -> dict:
    """Seeded state dict for FanEncoder (keys of the reference module tree): Kaiming-normal convolutions / Linears (gain chosen so that
    activations stay O(1) through 60 layers), non-trivial BatchNorm affine parameters AND running statistics (eval mode uses them)."""
    import torch.nn as nn
    from .fan_encoder import FanEncoder
    rng = np.random.default_rng(seed)
    sd = {}
    for mod_name, mod in FanEncoder().named_modules():
        pre = mod_name + "." if mod_name else ""
        if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
            c = mod.num_features
            sd[pre + "weight"] = _t(1.0 + 0.1 * rng.normal(size=c))
            sd[pre + "bias"] = _t(0.05 * rng.normal(size=c))
            sd[pre + "running_mean"] = _t(0.1 * rng.normal(size=c))
            sd[pre + "running_var"] = _t(1.0 + 0.2 * rng.uniform(-1, 1, size=c))
            sd[pre + "num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
        elif isinstance(mod, (nn.Conv2d, nn.Linear)):
            shape = tuple(mod.weight.shape)
            sd[pre + "weight"] = _t(rng.normal(0, math.sqrt(1.6 / int(np.prod(shape[1:]))), size=shape))
            if mod.bias is not None:
                sd[pre + "bias"] = _t(0.02 * rng.normal(size=shape[0]))
    return sd


def fan_images(n: int, seed: int = 81, size: int = 224) -> torch.Tensor:
    """[n,3,size,size] in [-1,1]-ish: smooth random fields (low-pass noise) so that neighbouring pixels correlate like an image."""
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, 3, size // 8, size // 8)).astype(np.float32)
    return torch.nn.functional.interpolate(torch.from_numpy(x), size=(size, size), mode="bilinear", align_corners=False).contiguous()


# ------------------------------------------------------------------------------------------------ training-step regularisers
def train_regularisers(B: int, T: int, fd: int, seed: int = 300, p: float = 0.1, n_layers: int = 12, heads: int = 12, hidden: int = 768,
                       ffn: int = 3072, dec_heads: int = 4, drop_layers=(4, 9), feat_proj_dropout: float = 0.0) -> dict:
    """Every stochastic draw of ONE faceformer_vert training step in train mode as explicit tensors (masks are Bernoulli(1 - p) / (1 - p),
    i.e. already scaled the way nn.Dropout scales): HF Wav2Vec2 feat_proj / hidden / attention / activation dropout, LayerDrop,
    SpecAugment (models/lib/wav2vec.py:16-63,120-139), the PPE dropout (faceformer_vert.py PeriodicPositionalEncoding) and the five
    dropouts of nn.TransformerDecoderLayer. `order` lists the dropout sites in the order the reference CALLS them. The feature
    projection's dropout is a site only when the config enables it (Wav2Vec2Config default feat_proj_dropout = 0.0)."""
    g = torch.Generator().manual_seed(seed)

    def mask(*shape):
        return (torch.rand(shape, generator=g) >= p).float() / (1.0 - p)

    M = B * T
    keep = [l not in set(drop_layers) for l in range(n_layers)]
    spec = torch.zeros(B, T, dtype=torch.bool)
    for b in range(B):                                          # two short spans per clip (min_masks = 2, wav2vec.py:127)
        for s in torch.randint(0, max(1, T - 3), (2,), generator=g).tolist():
            spec[b, s:s + 3] = True
    masks, order = {}, []

    def add(name, *shape):
        masks[name] = mask(*shape)
        order.append(name)

    if feat_proj_dropout > 0:
        add("featproj", M, hidden)
    add("enc_in", M, hidden)
    for l in range(n_layers):
        if keep[l]:
            add(f"l{l}.attn", B, heads, T, T)
            add(f"l{l}.h1", M, hidden)
            add(f"l{l}.act", M, ffn)
            add(f"l{l}.h3", M, hidden)
    add("ppe", M, fd)
    add("dec.sa", B, dec_heads, T, T)
    add("dec.d1", M, fd)
    add("dec.ca", B, dec_heads, T, T)        # memory attention: only the diagonal (the one visible key of every query) matters
    add("dec.d2", M, fd)
    add("dec.act", M, 2 * fd)
    add("dec.d3", M, fd)
    return {"p": p, "spec_mask": spec, "layer_keep": keep, "masks": masks, "order": order}
