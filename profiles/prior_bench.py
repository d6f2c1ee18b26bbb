"""BASELINE configs[3]: diffusion-prior sampling, batch 256, DDIM-64 and DDPM-100, from 768-d instruction embeddings
(BrainNetwork -> one-launch sampler). Prints samples/s, ms per call and the achieved fraction of 12.8 MFLOP/sample-step; also
times the CPU oracle on a bounded sample. Usage (GPU box): python profiles/prior_bench.py [batch]"""
import os
import sys
import time

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
# CPU oracle, bounded sample
from oracle import prior_oracle as po  # noqa: E402
sd = synth.prior_state()
nb = 16
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
po.voxel2style_emb(sd, inp["voxel"][:nb], inp["image_embed"][:nb], inp["noises"][:, :nb], timesteps_prior=64)
dt = time.perf_counter() - t0
print(f"CPU oracle ({torch.get_num_threads()} threads) DDIM-64 on {nb} samples: {dt * 1e3:.1f} ms -> {nb / dt:.1f} samples/s")
