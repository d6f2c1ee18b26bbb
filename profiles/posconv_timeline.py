"""Developer experiment (library built by profiles/build_timeline_lib.sh): where the implicit positional conv spends its time.
Per unit of pair 0: MMA issue span, cycles the MMA thread waited for weight stages, cycles the producer waited for free stages,
epilogue span. AVI_PC_DBG=1 makes every tap read slab row 0 (timing of aligned descriptors; wrong numbers)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from avi_talking_b200 import _lib, ops  # noqa: E402

B, T = 64, 249
g = torch.Generator(device="cuda").manual_seed(0)
xpad = (torch.randn(B, T + 128, 768, device="cuda", generator=g)).bfloat16()
band = (torch.randn(4, 128, 3, 96, 64, device="cuda", generator=g) * 0.01).bfloat16()
pb = torch.randn(768, device="cuda", generator=g)
for _ in range(3):
    ops.posconv_tc(xpad, band, pb, B, T, 16, 128)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.posconv_tc(xpad, band, pb, B, T, 16, 128)
e1.record()
torch.cuda.synchronize()
print("posconv_tc %.4f ms per call, AVI_PC_DBG=%s" % (e0.elapsed_time(e1) / 5, os.environ.get("AVI_PC_DBG", "0")))
buf = (C.c_longlong * (16 * 8))()
assert _lib.load().avi_debug_posconv_timeline(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(16, 8)
t0 = t[0, 0]
print("unit | mma start, mma end (issue), mma span | mma waited on full | producer waited on empty | epilogue start, end  (clk, rel. to unit 0 MMA start; 192 k-blocks per unit)")
for u in range(7):
    print(u, t[u, 0] - t0, t[u, 1] - t0, t[u, 1] - t[u, 0], "|", t[u, 2], "|", t[u, 3], "|", t[u, 4] - t0, t[u, 5] - t0)
