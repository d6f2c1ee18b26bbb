"""BASELINE configs[3]: diffusion-prior sampling, batch 256, DDIM-64 and DDPM-100, from 768-d instruction embeddings
(BrainNetwork -> one-launch sampler). Prints samples/s, ms per call and the achieved fraction of 12.8 MFLOP/sample-step; also
the CPU oracle of the same config is timed by `python bench.py --cpu-baseline prior`. Usage (GPU box): python profiles/prior_bench.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200 import synth  # noqa: E402
from avi_talking_b200.diffusion_prior import voxel2style_emb  # noqa: E402
from avi_talking_b200.smoke import build_prior  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
inp = synth.prior_inputs(B, 100)
voxel, x0 = inp["voxel"].cuda(), inp["image_embed"].cuda()
for prec in ("fp32", "bf16"):
    prior = build_prior(prec)
    for timesteps in (64, 100):
        steps = 63 if timesteps == 64 else 100
        noise = inp["noises"][:steps].cuda()
        for spc in (1, 2, 4):
            prior.samples_per_cta = spc
            fn = lambda: voxel2style_emb(voxel, prior, timesteps_prior=timesteps, image_embed=x0, noise=noise)  # noqa: E731
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{prec} B={B} timesteps={timesteps} samples/CTA={spc}: {ms:8.3f} ms/call  {B / ms * 1e3:10.0f} samples/s  "
                  f"{12.8e6 * B * steps / ms / 1e9:7.2f} TFLOP/s (denoiser, algorithmic)")
print("CPU oracle of the same sampler: python bench.py --cpu-baseline prior")
