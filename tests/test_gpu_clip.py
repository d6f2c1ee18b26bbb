"""GPU parity of the CLIP text tower drop-in against transformers.CLIPTextModel goldens (tests/golden/clip_text.npz)
and the CPU oracle. fp32 mode: |err| <= 5e-5 on LayerNorm-ed O(1) hidden states; bf16 GEMM mode: relative L2 error <= 1e-2."""
import numpy as np
import pytest
import torch

from avi_talking_b200 import synth

pytestmark = pytest.mark.gpu


def build(precision, layers):
    from transformers import CLIPTextConfig
    from avi_talking_b200.clip_text import CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=synth.CLIP_TEXT.vocab, hidden_size=768, intermediate_size=3072, num_hidden_layers=layers,
                         num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=768)
    m = CLIPTextModel(cfg)
    missing, unexpected = m.load_state_dict(synth.clip_text_state(60, layers), strict=False)
    assert not unexpected and all("position_ids" in k for k in missing)
    m.precision = precision
    return m.cuda().eval()


@pytest.mark.parametrize("tag,layers,B", [("l12", 12, 3), ("l2", 2, 2)])
def test_clip_text_matches_transformers_golden(golden, tag, layers, B):
    g = golden("clip_text")
    ids = synth.clip_tokens(B, seed=61).cuda()
    m = build("fp32", layers)
    out = m(input_ids=ids)
    last = out.last_hidden_state.cpu()
    assert np.abs(last[:, ::4, ::3].numpy() - g[f"{tag}_last_sub"]).max() <= 5e-5
    assert np.abs(m.text_to_voxel(ids).cpu().numpy() - g[f"{tag}_voxel"]).max() <= 2e-5
    eos = (ids == m.config.eos_token_id).int().argmax(-1)
    assert torch.equal(out.pooler_output, out.last_hidden_state[torch.arange(B, device="cuda"), eos])
    m16 = build("bf16", layers)
    v16 = m16.text_to_voxel(ids).cpu().numpy()
    rel = np.linalg.norm(v16 - g[f"{tag}_voxel"]) / np.linalg.norm(g[f"{tag}_voxel"])
    l16 = m16(input_ids=ids).last_hidden_state.cpu()[:, ::4, ::3].numpy()
    rel_h = np.linalg.norm(l16 - g[f"{tag}_last_sub"]) / np.linalg.norm(g[f"{tag}_last_sub"])
    print(f"CLIP text {tag}: bf16 relative L2 error voxel {rel:.2e}, hidden {rel_h:.2e}")
    assert rel <= 1e-2 and rel_h <= 1e-2


def test_text_to_style_end_to_end_on_device():
    """token ids -> CLIP text tower -> 77-token mean -> BrainNetwork -> DDIM prior -> style embedding [B,1,128], all in libavi_b200.so,
    equal to the prior driven by the oracle's CLIP output (fp32 mode)."""
    from oracle import clip_oracle as co
    from avi_talking_b200.diffusion_prior import voxel2style_emb
    from avi_talking_b200.smoke import build_prior
    B, layers = 4, 2
    ids = synth.clip_tokens(B, seed=62)
    m = build("fp32", layers)
    prior = build_prior("fp32")
    inp = synth.prior_inputs(B, 64)
    x0, noise = inp["image_embed"].cuda(), inp["noises"][:63].cuda()
    voxel = m.text_to_voxel(ids.cuda())
    style = voxel2style_emb(voxel, prior, timesteps_prior=64, image_embed=x0, noise=noise)
    with torch.no_grad():
        ref_voxel = co.text_to_voxel(synth.clip_text_state(60, layers), ids, layers)
    style_ref = voxel2style_emb(ref_voxel.cuda(), prior, timesteps_prior=64, image_embed=x0, noise=noise)
    assert style.shape == (B, 1, 128)
    assert (style - style_ref).abs().max().item() <= 1e-3
