"""Oracle: CLIP-L text tower as the reference uses it (TEST INFRASTRUCTURE; see oracle/__init__.py).

Reference: models/diffusion_prior.py:30-55  FrozenCLIPEmbedder.forward = transformers.CLIPTextModel(input_ids).last_hidden_state
([B, 77, 768]), then the 77-token mean at train_diffusion_prior.py:439,711.  The arithmetic lives in the `transformers` dependency
(reference pins 4.6.1 in requirements.txt; the container has 5.5.0): restated here from the published CLIP text architecture -
token + position embeddings, 12 pre-LN layers (12 heads x 64, scale 1/8, CAUSAL mask, quick_gelu MLP 3072), final LayerNorm.
Pinned by tests/golden/clip_text.npz = outputs of the installed transformers.CLIPTextModel (the class the reference imports) on the
seeded state dict of avi_talking_b200.synth.clip_text_state.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

H, D = 12, 64


def clip_text_forward(sd: dict, input_ids: torch.Tensor, layers: int = 12) -> torch.Tensor:
    p = "text_model."
    B, T = input_ids.shape
    x = sd[p + "embeddings.token_embedding.weight"][input_ids] + sd[p + "embeddings.position_embedding.weight"][:T][None]
    causal = torch.full((T, T), float("-inf")).triu(1)
    for l in range(layers):
        q = f"{p}encoder.layers.{l}."

        def lin(name, t):
            return F.linear(t, sd[q + name + ".weight"], sd[q + name + ".bias"])

        h = F.layer_norm(x, (x.shape[-1],), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], 1e-5)
        qq = lin("self_attn.q_proj", h).view(B, T, H, D).transpose(1, 2) * (D ** -0.5)
        kk = lin("self_attn.k_proj", h).view(B, T, H, D).transpose(1, 2)
        vv = lin("self_attn.v_proj", h).view(B, T, H, D).transpose(1, 2)
        a = torch.softmax(qq @ kk.transpose(2, 3) + causal, dim=-1)
        x = x + lin("self_attn.out_proj", (a @ vv).transpose(1, 2).reshape(B, T, H * D))
        h = F.layer_norm(x, (x.shape[-1],), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], 1e-5)
        f = lin("mlp.fc1", h)
        x = x + lin("mlp.fc2", f * torch.sigmoid(1.702 * f))
    return F.layer_norm(x, (x.shape[-1],), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)


def text_to_voxel(sd: dict, input_ids: torch.Tensor, layers: int = 12) -> torch.Tensor:
    """train_diffusion_prior.py:439,711: mean over the 77 tokens -> [B, 768] (the `voxel` BrainNetwork consumes)."""
    return clip_text_forward(sd, input_ids, layers).mean(dim=1)
