import os, sys
sys.path.insert(0, os.getcwd())
import torch
from avi_talking_b200.smoke import build_models
from avi_talking_b200 import ops
m = build_models("bf16")
P = m._pack()
hidden = torch.randn(64, 249, 64, device="cuda")
tmpl = m.template.cuda()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for pad in (True, False):
    ops.PAD_VERTEX_ROWS = pad
    out = m._vertex_head(hidden, P, tmpl)
    print("pad", pad, out.stride(), "%.4f ms" % t(lambda: m._vertex_head(hidden, P, tmpl)))
rng = torch.Generator(device="cuda").manual_seed(0)
coeff = torch.randn(64*249, 53, device="cuda"); pose = 0.1*torch.randn(64*249, 6, device="cuda"); shape = torch.randn(64*249, 100, device="cuda")
for pad in (True, False):
    ops.PAD_VERTEX_ROWS = pad
    fv = m.convert_coeff2verts(coeff, pose, shape)
    print("flame pad", pad, fv.stride(), "%.4f ms" % t(lambda: m.convert_coeff2verts(coeff, pose, shape)))
