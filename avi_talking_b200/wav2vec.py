"""Drop-in for models/lib/wav2vec.py: ``Wav2Vec2Model`` with the reference's forward signature
(models/lib/wav2vec.py:80-157) whose eval-mode computation runs in libavi_b200.so.

It subclasses the same transformers class the reference subclasses, so ``from_pretrained``, ``config``,
``feature_extractor._freeze_parameters()`` and every ``state_dict`` key are inherited unchanged; only ``forward`` is
replaced. Pipeline (time-major activations):

  conv0 + GroupNorm + GELU                      avi_w2v_conv0_gn_gelu
  conv1..6 (+GELU) as GEMMs over time           avi_gemm_bf16_tc (tcgen05) | avi_gemm_f32
  50 Hz -> 25 Hz lerp + LayerNorm(512)          avi_w2v_lerp_layernorm
  projection 512 -> 768                         GEMM
  positional grouped conv + GELU + res + LN     avi_w2v_posconv_ln
  12 x [QKV GEMM, MHA, out-proj GEMM(+res), LN, FFN GEMM(GELU), FFN GEMM(+res), LN]

precision = "bf16": GEMM operands bf16 (fp32 accumulate in TMEM), residual stream / LayerNorm / softmax fp32.
precision = "fp32": every operand fp32 on CUDA cores (the <=1e-5 mode).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F  # noqa: F401  (kept for parity with the reference module's namespace)
from transformers import Wav2Vec2Model as _HFWav2Vec2Model
from transformers.modeling_outputs import BaseModelOutput

from . import ops
from .ops import ACT_GELU, ACT_NONE


def default_precision() -> str:
    p = os.environ.get("AVI_B200_PRECISION", "bf16").lower()
    if p not in ("bf16", "fp32"):
        raise ValueError("AVI_B200_PRECISION must be bf16 or fp32")
    return p


def linear_interpolation_length(t50: int, input_fps=50, output_fps=25, output_len=None) -> int:
    """models/lib/wav2vec.py:67-71: output_len = int(seq_len / input_fps * output_fps) unless given."""
    if output_len is not None:
        return int(output_len)
    return int(t50 / float(input_fps) * output_fps)


def linear_interpolation(features, input_fps, output_fps, output_len=None):
    """models/lib/wav2vec.py:67-73, same signature: align_corners linear resample of [B, T, C] features along time to
    int(T / input_fps * output_fps) (or output_len) frames. Inside Wav2Vec2Model.forward the resample is fused with the feature
    projection's LayerNorm (avi_w2v_lerp_layernorm); this standalone form exists for callers of the function itself."""
    if not features.is_cuda:
        raise RuntimeError("avi_talking_b200.linear_interpolation runs on CUDA only (no CPU fallback)")
    B, T_in, Cc = features.shape
    T_out = linear_interpolation_length(T_in, input_fps, output_fps, output_len)
    x = features.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return ops.w2v_lerp(x, T_in * Cc, B, T_in, T_out, Cc).view(B, T_out, Cc)


def pack_posconv_band(w, groups):
    """Grouped conv weight [C, 48, k] fp32 -> the bf16 operand of avi_w2v_posconv_tc, [groups/4][k][3][96][64] (include/avi_b200.h)."""
    C, cg, k = w.shape
    wg = w.float().reshape(groups, cg, cg, k)                  # [group, co, ci, tap]
    band = torch.zeros(groups // 4, k, 3, 96, 64, dtype=torch.float32, device=w.device)
    for q in range(groups // 4):
        Wq = torch.zeros(4 * cg, 4 * cg, k, dtype=torch.float32, device=w.device)
        for gl in range(4):
            Wq[cg * gl:cg * (gl + 1), cg * gl:cg * (gl + 1)] = wg[4 * q + gl]
        for i, c in enumerate((0, 2, 1)):
            band[q, :, i] = Wq[48 * c:48 * c + 96, 64 * c:64 * c + 64].permute(2, 0, 1)
    return ops.cast_bf16(band.contiguous())


class Wav2Vec2Model(_HFWav2Vec2Model):
    def __init__(self, config):
        super().__init__(config)
        self.precision = default_precision()
        self._packed = None
        self._packed_key = None

    # ------------------------------------------------------------------ weight packing (once per weight change)
    def _pack_key(self):
        # ops.WEIGHT_EPOCH: bumped by the fused optimizer (train.FlatAdam), whose kernel updates weights without touching _version
        return (self.precision, ops.WEIGHT_EPOCH) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _posconv_weight(self):
        conv = self.encoder.pos_conv_embed.conv
        if hasattr(conv, "parametrizations"):
            g = conv.parametrizations.weight.original0
            v = conv.parametrizations.weight.original1
        else:  # transformers 4.x naming
            g, v = conv.weight_g, conv.weight_v
        return g * v / v.norm(p=2, dim=(0, 1), keepdim=True)

    @torch.no_grad()
    def _pack_extractor(self):
        """The (frozen, faceformer_vert.py:154) conv feature extractor's operands; cached on its own parameters only so that the
        training step, which rewrites every other weight each iteration, does not repack it."""
        cfg = self.config
        key = (self.precision,) + tuple((p.data_ptr(), p._version) for p in self.feature_extractor.parameters())
        if getattr(self, "_packed_fx", None) is not None and self._packed_fx_key == key:
            return self._packed_fx
        if cfg.feat_extract_norm != "group" or cfg.do_stable_layer_norm or cfg.conv_bias:
            raise NotImplementedError("only the wav2vec2-base layout (group-norm extractor, post-LN encoder, no conv bias)")
        bf16 = self.precision == "bf16"
        wdt = (lambda t: ops.cast_bf16(t)) if bf16 else (lambda t: t.detach().float().contiguous())
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        P = {}
        convs = self.feature_extractor.conv_layers
        P["conv0_w"] = f32(convs[0].conv.weight.reshape(cfg.conv_dim[0], cfg.conv_kernel[0]))
        P["gn_w"], P["gn_b"] = f32(convs[0].layer_norm.weight), f32(convs[0].layer_norm.bias)
        if bf16 and cfg.conv_dim[0] == 512 and cfg.conv_kernel[0] == 10 and cfg.conv_stride[0] == 5:
            P["conv0_w_tc"] = ops.conv0_pack_tc(P["conv0_w"])      # split-bf16 operand of the tensor-core conv0
        P["conv_w"] = []
        for i in range(1, len(convs)):
            w = convs[i].conv.weight  # [Cout, Cin, k] -> [Cout, k*Cin] tap-major (matches time-major activations)
            P["conv_w"].append(wdt(w.permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()))
        self._packed_fx, self._packed_fx_key = P, key
        return P

    @torch.no_grad()
    def _pack(self):
        key = self._pack_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        cfg = self.config
        bf16 = self.precision == "bf16"
        wdt = (lambda t: ops.cast_bf16(t)) if bf16 else (lambda t: t.detach().float().contiguous())
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        P = dict(self._pack_extractor())
        fp = self.feature_projection
        P["fp_ln_w"], P["fp_ln_b"] = f32(fp.layer_norm.weight), f32(fp.layer_norm.bias)
        P["fp_w"], P["fp_b"] = wdt(fp.projection.weight), f32(fp.projection.bias)
        k, g = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
        cg = cfg.hidden_size // g
        w = self._posconv_weight().float()  # [C, cg, k] -> [g][k][ci][co]
        P["pos_w"] = w.reshape(g, cg, cg, k).permute(0, 3, 2, 1).contiguous()
        P["pos_b"] = f32(self.encoder.pos_conv_embed.conv.bias)
        if bf16 and cg == 48 and g % 4 == 0 and k == 128:
            # tensor-core route (csrc/posconv_tc.cu): per quad of 4 groups (192 channels = 3 k-blocks of 64) the block-diagonal weight
            # Wq[n, ch, j] = w[192q + n, ch % 48, j] if ch // 48 == n // 48 else 0 meets k-block c only in output columns
            # [48c, 48c + 96): that band is packed as [quad][tap][i][96][64] with the k-blocks in stream order c = 0, 2, 1
            P["pos_w_band"] = pack_posconv_band(w, g)
        P["enc_ln_w"], P["enc_ln_b"] = f32(self.encoder.layer_norm.weight), f32(self.encoder.layer_norm.bias)
        P["layers"] = []
        for lyr in self.encoder.layers:
            a = lyr.attention
            L = {
                "qkv_w": wdt(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                "qkv_b": f32(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                "o_w": wdt(a.out_proj.weight), "o_b": f32(a.out_proj.bias),
                "ln1_w": f32(lyr.layer_norm.weight), "ln1_b": f32(lyr.layer_norm.bias),
                "ff1_w": wdt(lyr.feed_forward.intermediate_dense.weight), "ff1_b": f32(lyr.feed_forward.intermediate_dense.bias),
                "ff2_w": wdt(lyr.feed_forward.output_dense.weight), "ff2_b": f32(lyr.feed_forward.output_dense.bias),
                "ln2_w": f32(lyr.final_layer_norm.weight), "ln2_b": f32(lyr.final_layer_norm.bias),
            }
            P["layers"].append(L)
        self._packed, self._packed_key = P, key
        return P

    # ------------------------------------------------------------------ stages
    def _feature_extractor(self, x, P):
        """[B, N] fp32 -> time-major features [B, La, 512] (La = even-padded length), returns (tensor, T50, La)."""
        cfg = self.config
        B, n = x.shape
        adt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        Cc = cfg.conv_dim[0]
        L = (n - cfg.conv_kernel[0]) // cfg.conv_stride[0] + 1
        La = L + (L & 1)
        h = torch.empty((B, La, Cc), dtype=adt, device=x.device)
        if "conv0_w_tc" in P:
            ops.conv0_gn_gelu_tc(x, P["conv0_w"], P["conv0_w_tc"], P["gn_w"], P["gn_b"], h, La * Cc)
        else:
            ops.conv0_gn_gelu(x, P["conv0_w"], P["gn_w"], P["gn_b"], h, La * Cc)
        for i in range(1, len(cfg.conv_kernel)):
            k, s = cfg.conv_kernel[i], cfg.conv_stride[i]
            Lo = (L - k) // s + 1
            Loa = Lo + (Lo & 1)
            Co = cfg.conv_dim[i]
            o = torch.empty((B, Loa, Co), dtype=adt, device=x.device)
            ops.gemm(h, P["conv_w"][i - 1], None, o, batch=B, rows=Lo, N=Co, K=k * Cc, act=ACT_GELU, conv_taps=k, conv_stride=s,
                     a_ld=Cc, a_batch_stride=La * Cc, a_rows_alloc=La, c_ld=Co, c_batch_stride=Loa * Co)
            h, L, La, Cc = o, Lo, Loa, Co
        return h, L, La

    def _posconv_tc(self, proj, P, B, T):
        """Positional grouped conv on the tcgen05 path as an implicit convolution: the zero-padded bf16 activation slab stays in shared
        memory and the 128 taps walk it by descriptor row offset; only the non-zero band of the grouped weight is contracted."""
        cfg = self.config
        k = cfg.num_conv_pos_embeddings
        Tp = T + k                                              # rows t-64 .. t+63 around every output row, zero padded
        xpad = ops.pad_cast_bf16(proj, B, T, k // 2, Tp)
        pc = ops.posconv_tc(xpad, P["pos_w_band"], P["pos_b"], B, T, cfg.num_conv_pos_embedding_groups, k)
        return ops.posconv_merge_ln(proj, pc, P["enc_ln_w"], P["enc_ln_b"], want_bf16=True, eps=cfg.layer_norm_eps)

    def _encoder_layer(self, h32, h16, Lw, B, T, inplace=False):
        """inplace (bf16 tensor-core path only): the two residual additions update the fp32 residual stream where it lies - the
        GEMM epilogue leaves through TMA reduce-add, so the residual is never read by an SM nor written to a second buffer."""
        bf16 = self.precision == "bf16"
        H = self.config.num_attention_heads
        D = self.config.hidden_size // H
        a_in = h16 if bf16 else h32
        adt = torch.bfloat16 if bf16 else torch.float32
        qkv = ops.linear(a_in, Lw["qkv_w"], Lw["qkv_b"], out_dtype=adt)
        att = ops.mha(qkv, B, T, H, D, D ** -0.5)
        y = ops.linear(att, Lw["o_w"], Lw["o_b"], residual=h32, out_dtype=torch.float32, out=h32 if inplace else None)
        h1_32, h1_16 = ops.layernorm(y, Lw["ln1_w"], Lw["ln1_b"], want_bf16=bf16, eps=self.config.layer_norm_eps)
        f = ops.linear(h1_16 if bf16 else h1_32, Lw["ff1_w"], Lw["ff1_b"], act=ACT_GELU, out_dtype=adt)
        y2 = ops.linear(f, Lw["ff2_w"], Lw["ff2_b"], residual=h1_32, out_dtype=torch.float32, out=h1_32 if inplace else None)
        return ops.layernorm(y2, Lw["ln2_w"], Lw["ln2_b"], want_bf16=bf16, eps=self.config.layer_norm_eps)

    # ------------------------------------------------------------------ reference signature
    @torch.no_grad()
    def forward(self, input_values, dataset=None, attention_mask=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, frame_num=None):
        if self.training and (self.config.apply_spec_augment or self.config.layerdrop > 0):
            raise NotImplementedError("train-mode SpecAugment / LayerDrop are not on the inference hot path; call .eval()")
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is never passed on the AVI-Talking path (faceformer_disentangle.py:775)")
        if not input_values.is_cuda:
            raise RuntimeError("avi_talking_b200.Wav2Vec2Model runs on CUDA only (no CPU fallback)")
        cfg = self.config
        P = self._pack()
        x = input_values.contiguous().float()
        B = x.shape[0]
        bf16 = self.precision == "bf16"
        feats, T50, La = self._feature_extractor(x, P)                                      # wav2vec.py:97
        T = linear_interpolation_length(T50, 50, 25, frame_num)                             # :108
        Cf = cfg.conv_dim[-1]
        hn32, hn16 = ops.lerp_layernorm(feats, La * Cf, B, T50, T, P["fp_ln_w"], P["fp_ln_b"], want_f32=not bf16,
                                        want_bf16=bf16, eps=cfg.layer_norm_eps)
        proj = ops.linear(hn16 if bf16 else hn32, P["fp_w"], P["fp_b"], out_dtype=torch.float32)   # :120
        if "pos_w_band" in P:
            h32, h16 = self._posconv_tc(proj, P, B, T)
        else:
            h32, h16 = ops.posconv_ln(proj, P["pos_w"], P["pos_b"], P["enc_ln_w"], P["enc_ln_b"], B, T,
                                      cfg.num_conv_pos_embedding_groups, cfg.num_conv_pos_embeddings, want_bf16=bf16,
                                      eps=cfg.layer_norm_eps)
        all_hidden = (h32.view(B, T, -1),) if output_hidden_states else None
        for Lw in P["layers"]:                                                               # :142-148
            h32, h16 = self._encoder_layer(h32, h16, Lw, B, T, inplace=bf16 and not output_hidden_states)
            if output_hidden_states:
                all_hidden = all_hidden + (h32.view(B, T, -1),)
        out = h32.view(B, T, cfg.hidden_size)
        self.last_hidden_state_bf16 = h16.view(B, T, cfg.hidden_size) if bf16 else None   # for bf16 consumers (EMOTE encoder)
        # The reference forces output_attentions=True (:90) but no caller reads them (faceformer_disentangle.py:535,638,775
        # use .last_hidden_state only); the fused attention kernel never materialises the [B,12,T,T] maps.
        if return_dict is False:
            return (out,) + ((all_hidden,) if all_hidden is not None else ())
        return BaseModelOutput(last_hidden_state=out, hidden_states=all_hidden, attentions=None)
