"""Drop-in for the CLIP text tower of models/diffusion_prior.py:30-55: ``CLIPTextModel`` subclasses the transformers
class the reference instantiates (same config, ``state_dict`` keys and ``from_pretrained``); ``forward(input_ids)`` runs in
libavi_b200.so and returns ``last_hidden_state [B, 77, 768]``; ``FrozenCLIPEmbedder`` keeps the reference wrapper's interface
(a tokenizer is needed only when strings are passed; token ids are accepted directly, there is no network here to fetch a vocabulary).

Pipeline per layer (pre-LN):  LN1 -> fused QKV GEMM -> causal attention (12 x 64) -> out-proj GEMM accumulating into the fp32 residual
stream in place (TMA reduce-add) -> LN2 -> fc1 GEMM with quick_gelu in the epilogue -> fc2 GEMM in place; final LN; `text_to_voxel`
adds the 77-token mean of train_diffusion_prior.py:439,711.
precision "bf16": tcgen05 GEMMs (bf16 operands, fp32 accumulate), everything else fp32; "fp32": CUDA-core GEMMs.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from transformers import CLIPTextModel as _HFCLIPTextModel
from transformers.modeling_outputs import BaseModelOutputWithPooling

from . import ops
from .ops import ACT_QUICK_GELU
from .wav2vec import default_precision


class CLIPTextModel(_HFCLIPTextModel):
    def __init__(self, config):
        super().__init__(config)
        self.precision = default_precision()
        self._packed, self._packed_key = None, None

    @torch.no_grad()
    def _pack(self):
        key = (self.precision, ops.WEIGHT_EPOCH) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and key == self._packed_key:
            return self._packed
        cfg = self.config
        if cfg.hidden_act != "quick_gelu":
            raise NotImplementedError("only the quick_gelu CLIP text tower (openai/clip-vit-large-patch14) is built")
        bf16 = self.precision == "bf16"
        wdt = (lambda t: ops.cast_bf16(t)) if bf16 else (lambda t: t.detach().float().contiguous())
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        tm = self.text_model
        P = {"tok": f32(tm.embeddings.token_embedding.weight), "pos": f32(tm.embeddings.position_embedding.weight),
             "fln_w": f32(tm.final_layer_norm.weight), "fln_b": f32(tm.final_layer_norm.bias), "layers": []}
        for lyr in tm.encoder.layers:
            a = lyr.self_attn
            P["layers"].append({
                "qkv_w": wdt(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                "qkv_b": f32(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                "o_w": wdt(a.out_proj.weight), "o_b": f32(a.out_proj.bias),
                "ln1_w": f32(lyr.layer_norm1.weight), "ln1_b": f32(lyr.layer_norm1.bias),
                "ln2_w": f32(lyr.layer_norm2.weight), "ln2_b": f32(lyr.layer_norm2.bias),
                "fc1_w": wdt(lyr.mlp.fc1.weight), "fc1_b": f32(lyr.mlp.fc1.bias),
                "fc2_w": wdt(lyr.mlp.fc2.weight), "fc2_b": f32(lyr.mlp.fc2.bias)})
        self._packed, self._packed_key = P, key
        return P

    @torch.no_grad()
    def hidden(self, input_ids):
        """[B, T<=77] token ids -> final-LayerNorm hidden states, fp32 [B*T, 768]."""
        if not input_ids.is_cuda:
            raise RuntimeError("avi_talking_b200.CLIPTextModel runs on CUDA only (no CPU fallback)")
        cfg = self.config
        B, T = input_ids.shape
        if T > cfg.max_position_embeddings or T > 128:
            raise ValueError(f"sequence length {T} exceeds max_position_embeddings")
        P = self._pack()
        bf16 = self.precision == "bf16"
        adt = torch.bfloat16 if bf16 else torch.float32
        H = cfg.num_attention_heads
        D = cfg.hidden_size // H
        eps = cfg.layer_norm_eps
        x = ops.embed_tokens(input_ids, P["tok"], P["pos"])
        for L in P["layers"]:
            h32, h16 = ops.layernorm(x, L["ln1_w"], L["ln1_b"], want_f32=not bf16, want_bf16=bf16, eps=eps)
            qkv = ops.linear(h16 if bf16 else h32, L["qkv_w"], L["qkv_b"], out_dtype=torch.float32)
            att, _ = ops.attn_train_fwd(qkv, B, T, H, D, bias_mode=2, want_p=False)       # causal, scale 1/sqrt(D)
            att = ops.cast_bf16(att) if bf16 else att
            x = ops.linear(att, L["o_w"], L["o_b"], residual=x, out_dtype=torch.float32, out=x if bf16 else None)
            h32, h16 = ops.layernorm(x, L["ln2_w"], L["ln2_b"], want_f32=not bf16, want_bf16=bf16, eps=eps)
            f = ops.linear(h16 if bf16 else h32, L["fc1_w"], L["fc1_b"], act=ACT_QUICK_GELU, out_dtype=adt)
            x = ops.linear(f, L["fc2_w"], L["fc2_b"], residual=x, out_dtype=torch.float32, out=x if bf16 else None)
        return ops.layernorm(x, P["fln_w"], P["fln_b"], eps=eps)[0]

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, **kw):
        """CLIPTextModel.forward(input_ids=...) as called at models/diffusion_prior.py:49. Padding attention masks are not used on
        that path (the tokenizer pads with EOS and the model runs causally over all 77 positions)."""
        if attention_mask is not None or position_ids is not None:
            raise NotImplementedError("attention_mask / position_ids are not passed on the AVI-Talking path (diffusion_prior.py:49)")
        B, T = input_ids.shape
        last = self.hidden(input_ids).view(B, T, -1)
        eos = (input_ids.to(torch.int) == self.config.eos_token_id).int().argmax(dim=-1)
        pooled = last[torch.arange(B, device=last.device), eos]
        return BaseModelOutputWithPooling(last_hidden_state=last, pooler_output=pooled)

    @torch.no_grad()
    def text_to_voxel(self, input_ids):
        """last_hidden_state.mean(dim=1): the [B, 768] `voxel` fed to BrainNetwork (train_diffusion_prior.py:439,711)."""
        B, T = input_ids.shape
        return ops.token_mean(self.hidden(input_ids), B, T)


class FrozenCLIPEmbedder(nn.Module):
    """models/diffusion_prior.py:30-55, same constructor `(version, device, max_length)`: tokenizer and text tower come from
    `from_pretrained(version)` (the transformers cache; there is no network on the B200 boxes), the tower being the drop-in above.
    `transformer=` / `tokenizer=` inject already-built parts (tests, synthetic weights); `forward` also accepts a LongTensor of token ids."""

    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77, *, transformer: CLIPTextModel = None, tokenizer=None):
        super().__init__()
        if transformer is None:
            from transformers import CLIPTokenizer
            tokenizer = CLIPTokenizer.from_pretrained(version) if tokenizer is None else tokenizer      # :36
            transformer = CLIPTextModel.from_pretrained(version)                                         # :37
        self.tokenizer, self.transformer, self.device, self.max_length = tokenizer, transformer, device, max_length
        self.freeze()

    def freeze(self):
        self.transformer = self.transformer.eval()
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, text):
        if torch.is_tensor(text):
            tokens = text.to(self.device)
        else:
            if self.tokenizer is None:
                raise RuntimeError("FrozenCLIPEmbedder: pass token ids, or construct it with a CLIPTokenizer")
            enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True, return_overflowing_tokens=False,
                                 padding="max_length", return_tensors="pt")
            tokens = enc["input_ids"].to(self.device)
        return self.transformer(input_ids=tokens).last_hidden_state

    def encode(self, text):
        return self(text)
