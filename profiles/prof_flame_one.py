"""convert_coeff2verts (FLAME blend + LBS, tensor-core path) on the configs[1] frame count, for `ncu -k regex:flame_tc`.
Usage: python profiles/prof_flame_one.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200.smoke import build_models  # noqa: E402

m = build_models("bf16")
F = 64 * 249
coeff = torch.randn(F, 53, device="cuda")
pose = 0.1 * torch.randn(F, 6, device="cuda")
shape = torch.randn(F, 100, device="cuda")
for _ in range(3):
    m.convert_coeff2verts(coeff, pose, shape)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m.convert_coeff2verts(coeff, pose, shape)
e1.record()
torch.cuda.synchronize()
print("ok flame %.4f ms" % (e0.elapsed_time(e1) / 10))
