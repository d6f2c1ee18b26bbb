"""FanEncoder image branch: images/s of the drop-in in both precisions, batch of 16 images (one chunk), CUDA events.
Usage (GPU box): python profiles/fan_bench.py [n_images]. CPU side: the reference module itself is PyTorch (time it with torch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200 import ops, synth  # noqa: E402
from avi_talking_b200.fan_encoder import FanEncoder  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = synth.fan_images(n, seed=81).cuda()
for prec in ("fp32", "tf32", "bf16"):
    m = FanEncoder()
    m.load_state_dict(synth.fan_state(80))
    m.precision = prec
    m = m.cuda().eval()
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ops.PROFILE = []
    m(x)
    torch.cuda.synchronize()
    agg = {}
    for name, a, b, work in ops.PROFILE:
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
        d[2] += work
    ops.PROFILE = None
    gk = {"bf16": "gemm_bf16_tc", "tf32": "gemm_tf32_tc", "fp32": "gemm_f32"}[prec]
    print(f"{prec}: {ms:.3f} ms per {n} images = {n / ms * 1e3:.0f} images/s; " +
          ", ".join(f"{k}: {v[0]} launches {v[1]:.3f} ms" for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])) +
          f"; GEMM {agg[gk][2] / agg[gk][1] / 1e9:.1f} TFLOP/s algorithmic ({agg[gk][2] / n / 1e9:.2f} GFLOP per image)")
