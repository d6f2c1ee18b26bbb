"""CLIP-L text tower -> 77-token mean -> BrainNetwork -> DDIM-64 prior: instruction tokens to style embedding,
batch 256 (BASELINE configs[3] with its text front end). Usage (GPU box): python profiles/clip_bench.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from transformers import CLIPTextConfig  # noqa: E402

from avi_talking_b200 import synth  # noqa: E402
from avi_talking_b200.clip_text import CLIPTextModel  # noqa: E402
from avi_talking_b200.diffusion_prior import voxel2style_emb  # noqa: E402
from avi_talking_b200.smoke import build_prior  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = CLIPTextConfig(vocab_size=synth.CLIP_TEXT.vocab, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                     num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=768)
ids = synth.clip_tokens(B, seed=61).cuda()
inp = synth.prior_inputs(B, 64)
x0, noise = inp["image_embed"].cuda(), inp["noises"][:63].cuda()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for prec in ("fp32", "bf16"):
    m = CLIPTextModel(cfg)
    m.load_state_dict(synth.clip_text_state(60, 12), strict=False)
    m.precision = prec
    m = m.cuda().eval()
    prior = build_prior(prec)
    prior.samples_per_cta = 2
    ms_clip = timeit(lambda: m.text_to_voxel(ids))
    ms_all = timeit(lambda: voxel2style_emb(m.text_to_voxel(ids), prior, timesteps_prior=64, image_embed=x0, noise=noise))
    flops = 2.0 * B * 77 * 12 * (768 * 2304 + 768 * 768 + 2 * 768 * 3072) + 4.0 * B * 12 * 12 * 77 * 77 * 64
    print(f"{prec} B={B}: CLIP text tower {ms_clip:.3f} ms ({flops / ms_clip / 1e9:.0f} TFLOP/s), tokens -> style embedding {ms_all:.3f} ms "
          f"({B / ms_all * 1e3:.0f} instructions/s)")
print("CPU oracle of the text tower: python bench.py --cpu-baseline clip")
