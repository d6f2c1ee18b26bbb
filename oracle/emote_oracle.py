"""Oracle: EMOTE talking-head inference (Path B, what experiments/diffusion_test.sh runs) - TEST INFRASTRUCTURE (oracle/__init__.py).

fp32 torch-on-CPU restatement of (paths under third_party/inferno/)
  inferno_apps/TalkingHead/evaluation/TalkingHeadWrapper.py:123-138   TalkingHeadWrapper.forward (renderer None)
  inferno/models/talkinghead/TalkingHeadBase.py:503-553               forward: preprocess -> audio -> encode -> decode
  inferno/models/temporal/Preprocessors.py:62-186                     FlamePreprocessor._forward (gt_shape given, no texture)
  inferno/models/temporal/AudioEncoders.py:16-24,38-90,168-201        z-norm processor, Wav2Vec2ModelResampled(desired_output_length)
  inferno/models/temporal/SequenceEncoders.py:180-197                 LinearSequenceEncoder
  inferno/models/talkinghead/FaceFormerDecoder.py:166-267             EmotionCondition._gather_condition / LinearEmotionCondition
  .../FaceFormerDecoder.py:598-612,652-682,967-985,1104-1224          FeedForwardDecoder.forward/_style, StackLinearSquash,
                                                                      BertPriorDecoder._decode/_apply_motion_prior/_neutral_shape
  inferno/models/temporal/motion_prior/L2lMotionPrior.py:460-495      L2lDecoder.forward
  inferno/models/temporal/motion_prior/MotionPrior.py:316-351,376-380 decompose_sequential_output / postprocess / decoding_step
  inferno/models/temporal/TransformerMasking.py:80-98                 init_alibi_biased_mask_future
with the EMOTE configuration (talkinghead_conf/model/sequence_decoder/bertprior_wild.yaml, motion_prior_conf l2l_*): feature_dim 128,
8 heads, 1 post-LN GELU encoder layer (ff 128), no positional encoding, style_op 'add', post_bug_fix, squash_after with
stack_linear over 8 frames, L2L decoder (quant_factor 3, d 256, ff 384, alibi_future), 50 expression + 3 jaw outputs.
Pinned by tests/golden/emote.npz: outputs of the reference's own classes/methods (oracle/make_golden.py, stub-imported).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import flame_oracle as fo
from . import wav2vec2_oracle as wo
from .faceformer_oracle import get_slopes


def znorm(raw_audio: torch.Tensor) -> torch.Tensor:
    """Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm as called from AudioEncoders.py:170-178: the reference passes one
    [B, N] tensor, which HF treats as ONE example (joint statistics); the reference only ever calls it with B = 1."""
    x = raw_audio.reshape(raw_audio.shape[0], -1).float()
    return (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-7)


def encoder_layer(sd, p, x, nhead, mask=None):
    """torch.nn.TransformerEncoderLayer(batch_first, post-LN, GELU), eval mode; mask [nhead, T, T] additive or None."""
    B, T, D = x.shape
    hd = D // nhead
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = (t.view(B, T, nhead, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if mask is not None:
        s = s + mask[None]
    a = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, T, D)
    x = F.layer_norm(x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"]), (D,),
                     sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    f = F.linear(F.gelu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])), sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return F.layer_norm(x + f, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)


def alibi_future_mask(nhead, T):
    """TransformerMasking.py:80-98 sliced [:T,:T]: -slope_h * |i - j|."""
    slopes = torch.tensor(get_slopes(nhead))
    i = torch.arange(T)
    return -slopes[:, None, None] * (i[:, None] - i[None, :]).abs().float()[None]


def l2l_decoder(sd, p, z, nhead=8):
    """L2lDecoder.forward (L2lMotionPrior.py:460-495): z [B, T/8, 256] -> [B, T, 53]."""
    x = z
    for i in range(3):
        xc = x.permute(0, 2, 1)
        if i == 0:
            y = F.conv_transpose1d(xc, sd[p + "expander.0.0.weight"], sd[p + "expander.0.0.bias"], stride=2, padding=2, output_padding=1)
        else:
            y = F.conv1d(F.pad(xc, (2, 2), mode="replicate"), sd[p + f"expander.{i}.0.weight"], sd[p + f"expander.{i}.0.bias"])
        y = F.leaky_relu(y, 0.2)
        y = F.batch_norm(y, sd[p + f"expander.{i}.2.running_mean"], sd[p + f"expander.{i}.2.running_var"],
                         sd[p + f"expander.{i}.2.weight"], sd[p + f"expander.{i}.2.bias"], False, 0.0, 1e-5)
        x = y.permute(0, 2, 1)
        if i > 0:
            x = x.repeat_interleave(2, dim=1)
    x = F.linear(x, sd[p + "decoder_linear_embedding.weight"], sd[p + "decoder_linear_embedding.bias"])
    x = encoder_layer(sd, p + "decoder_transformer.layers.0.", x, nhead, alibi_future_mask(nhead, x.shape[1]))
    return F.conv1d(x.permute(0, 2, 1), sd[p + "cross_smooth_layer.weight"], sd[p + "cross_smooth_layer.bias"], padding=2).permute(0, 2, 1)


def flame_from_coeffs(buf, shape, exp, jaw):
    """FlamePreprocessor._forward (Preprocessors.py:62-186) with gt_shape [B, n_shape]: vertices [B,T,V*3] + template [B,V*3]."""
    B, T = exp.shape[:2]
    pose = torch.cat([torch.zeros_like(jaw), jaw], dim=-1).reshape(B * T, 6)
    shp = shape[:, None].expand(B, T, shape.shape[1]).reshape(B * T, -1)
    verts = fo.flame_forward(buf, shp, exp.reshape(B * T, -1), pose)[0]
    template = fo.flame_forward(buf, shape, torch.zeros(B, exp.shape[-1]), None)[0]
    return verts.reshape(B, T, -1), template.reshape(B, -1)


@torch.no_grad()
def talking_head_forward(sd, sd_w2v, buf, sample, style_emb=None, is_external_style_emb=False, per_clip_znorm=False):
    """TalkingHeadWrapper.forward(sample, style_emb, is_external_style_emb) -> dict with gt_vertices, template, predicted_exp,
    predicted_jaw, predicted_vertices. `per_clip_znorm`: normalise every clip on its own (what the batched drop-in does; equal
    to the reference for B = 1)."""
    out = dict(sample)
    B, T = sample["raw_audio"].shape[:2]
    out["gt_vertices"], out["template"] = flame_from_coeffs(buf, sample["gt_shape"], sample["gt_exp"], sample["gt_jaw"])   # preprocess
    raw = sample["raw_audio"].reshape(B, -1)
    audio = torch.cat([znorm(raw[b:b + 1]) for b in range(B)]) if per_clip_znorm else znorm(raw)
    out["processed_audio"] = audio
    feat = wo.wav2vec2_forward(sd_w2v, audio, frame_num=T)                                                                  # AudioEncoders.py:186-188
    out["audio_feature"] = feat
    h = F.linear(feat, sd["sequence_encoder.linear.weight"], sd["sequence_encoder.linear.bias"])                            # SequenceEncoders.py:189-197
    if not (style_emb is not None and is_external_style_emb):
        cond = torch.cat([sample["gt_expression_label_condition"], sample["gt_expression_intensity_condition"],
                          sample["gt_expression_identity_condition"]], dim=-1).float()                                      # FaceFormerDecoder.py:200-241
        style_emb = F.linear(cond, sd["sequence_decoder.obj_vector.map.weight"], sd["sequence_decoder.obj_vector.map.bias"])
    h = h + style_emb                                                                                                       # :669-670
    d = "sequence_decoder."
    h = encoder_layer(sd, d + "bert_decoder.layers.0.", h, 8)                                                               # :1210
    z = F.linear(h, sd[d + "decoder.weight"], sd[d + "decoder.bias"])                                                       # :1221 (post_bug_fix)
    Tp = int(math.ceil(T / 8) * 8)
    z = F.pad(z, (0, 0, 0, Tp - T))                                                                                         # :1111-1125
    z = F.linear(z.reshape(B, Tp // 8, -1), sd[d + "squasher_2.linear.weight"], sd[d + "squasher_2.linear.bias"])           # :975-985
    out["prior_input_sequence"] = z
    seq = l2l_decoder(sd, d + "motion_prior.motion_decoder.", z)[:, :T]                                                    # MotionPrior.py:376-380
    exp, jaw = seq[..., :50], seq[..., 50:53]                                                                               # :316-329
    verts, _ = flame_from_coeffs(buf, sample["gt_shape"], exp, jaw)                                                         # :331-351
    neutral = fo.flame_forward(buf, sample["gt_shape"], torch.zeros(B, 50))[0].reshape(B, 1, -1)                             # :1184-1192
    out["predicted_exp"], out["predicted_jaw"] = exp, jaw
    out["predicted_vertices"] = (verts - neutral) + out["template"][:, None]                                                # :1173-1175, :690-694
    return out
