"""Oracle: wav2vec2 audio encoder as the reference wraps it (TEST INFRASTRUCTURE; see oracle/__init__.py).

fp32 functional-torch restatement, driven by a plain state dict, of
  models/lib/wav2vec.py  linear_interpolation :67-73, Wav2Vec2Model.forward :80-157 (eval, no attention_mask)
  third_party/inferno/inferno/models/temporal/AudioEncoders.py :16-24,38-90 (Path-B twin; ceil instead of floor)
and of the un-vendored dependency it subclasses, transformers' Wav2Vec2 (reference pins transformers==4.6.1,
requirements.txt:8; restated from the published architecture of wav2vec2-base = Wav2Vec2Config() defaults):
  conv feature extractor (7 Conv1d no bias, GroupNorm(512,512) on layer 0, exact-erf GELU),
  feature projection (LayerNorm(512) -> Linear(512,768)),
  positional conv (weight-normed grouped Conv1d k=128 pad=64 g=16, drop last, GELU) + LayerNorm,
  12 post-LN encoder layers (12 heads x 64, scale 1/8, FFN 3072 GELU).
Pinned by tests/golden/w2v_*.npz (outputs of the reference's own Wav2Vec2Model subclass on the same state dict).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .synth import W2V


def feat_lengths(n: int) -> list[int]:
    out = []
    for k, s in zip(W2V.conv_kernel, W2V.conv_stride):
        n = (n - k) // s + 1
        out.append(n)
    return out


def output_frames(n_samples: int, mode: str = "floor") -> int:
    """T = int(T50/50*25) (wav2vec.py:69-71) or ceil (AudioEncoders.py:33-36)."""
    t50 = feat_lengths(n_samples)[-1]
    return int(t50 / 50.0 * 25) if mode == "floor" else int(math.ceil(t50 / 50.0 * 25))


def feature_extractor(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """[B,N] -> [B,512,T50]."""
    h = x[:, None]
    for i, s in enumerate(W2V.conv_stride):
        h = F.conv1d(h, sd[f"feature_extractor.conv_layers.{i}.conv.weight"], stride=s)
        if i == 0:
            h = F.group_norm(h, 512, sd["feature_extractor.conv_layers.0.layer_norm.weight"],
                             sd["feature_extractor.conv_layers.0.layer_norm.bias"], eps=1e-5)
        h = F.gelu(h)
    return h


def linear_interpolation(features: torch.Tensor, output_len: int) -> torch.Tensor:
    """[B,T50,C] -> [B,T,C], align_corners=True (wav2vec.py:67-73)."""
    return F.interpolate(features.transpose(1, 2), size=output_len, align_corners=True,
                         mode="linear").transpose(1, 2)


def pos_conv_weight(sd: dict) -> torch.Tensor:
    """weight_norm(dim=2): w = g * v / ||v|| with the norm over dims (0,1) per tap."""
    g = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    v = sd["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    return g * v / v.norm(p=2, dim=(0, 1), keepdim=True)


def _drop(reg, name, x):
    """nn.Dropout in train mode with the (pre-scaled) mask of site `name` as an input; identity without regularisers."""
    if reg is None or name not in reg["masks"]:
        return x
    return x * reg["masks"][name].reshape(x.shape)


def encoder_layer(sd: dict, l: int, h: torch.Tensor, reg=None) -> torch.Tensor:
    """HF Wav2Vec2EncoderLayer; reg (train mode): attention-probability dropout, the hidden dropout after out_proj, the activation
    dropout inside the feed-forward and its output dropout (modeling_wav2vec2.py Wav2Vec2EncoderLayer / Wav2Vec2FeedForward)."""
    p = f"encoder.layers.{l}."
    B, T, C = h.shape
    H, D = W2V.heads, C // W2V.heads

    def lin(name, x):
        return F.linear(x, sd[p + name + ".weight"], sd[p + name + ".bias"])

    q = lin("attention.q_proj", h).view(B, T, H, D).transpose(1, 2)
    k = lin("attention.k_proj", h).view(B, T, H, D).transpose(1, 2)
    v = lin("attention.v_proj", h).view(B, T, H, D).transpose(1, 2)
    a = _drop(reg, f"l{l}.attn", torch.softmax(torch.matmul(q, k.transpose(2, 3)) * (D ** -0.5), dim=-1))
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, T, C)
    h = h + _drop(reg, f"l{l}.h1", lin("attention.out_proj", o))
    h = F.layer_norm(h, (C,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
    ff = lin("feed_forward.output_dense", _drop(reg, f"l{l}.act", F.gelu(lin("feed_forward.intermediate_dense", h))))
    h = h + _drop(reg, f"l{l}.h3", ff)
    return F.layer_norm(h, (C,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)


def encoder(sd: dict, h: torch.Tensor, layers: int = 12, reg=None) -> torch.Tensor:
    pc = F.conv1d(h.transpose(1, 2), pos_conv_weight(sd), sd["encoder.pos_conv_embed.conv.bias"],
                  padding=W2V.pos_k // 2, groups=W2V.pos_groups)[:, :, :-1]
    h = h + F.gelu(pc).transpose(1, 2)
    h = _drop(reg, "enc_in", F.layer_norm(h, (768,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], 1e-5))
    for l in range(layers):
        if reg is None or reg["layer_keep"][l]:          # LayerDrop: a skipped layer is the identity (Wav2Vec2Encoder.forward)
            h = encoder_layer(sd, l, h, reg)
    return h


@torch.no_grad()
def wav2vec2_forward(sd: dict, input_values: torch.Tensor, frame_num: int | None = None,
                     mode: str = "floor", layers: int = 12, return_stages: bool = False, reg=None):
    """Wav2Vec2Model.forward(input_values, dataset, frame_num=...) -> last_hidden_state [B,T,768]."""
    feats = feature_extractor(sd, input_values).transpose(1, 2)                      # wav2vec.py:97-98
    T = frame_num if frame_num is not None else output_frames(input_values.shape[1], mode)
    h = linear_interpolation(feats, T)                                               # :108
    hn = F.layer_norm(h, (512,), sd["feature_projection.layer_norm.weight"],
                      sd["feature_projection.layer_norm.bias"], 1e-5)
    proj = F.linear(hn, sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"])  # :120
    if reg is not None:                                                              # train mode
        proj = _drop(reg, "featproj", proj)                                          # Wav2Vec2FeatureProjection.dropout
        if reg.get("spec_mask") is not None:                                         # SpecAugment along time, wav2vec.py:122-131
            proj = torch.where(reg["spec_mask"][:, :, None], sd["masked_spec_embed"].to(proj.dtype)[None, None, :], proj)
    out = encoder(sd, proj, layers, reg)                                             # :142-148
    if return_stages:
        return dict(feats=feats, interp=h, proj=proj, out=out)
    return out
