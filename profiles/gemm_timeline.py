"""Developer experiment (needs a library built with AVI_NVCC_EXTRA=-DAVI_GEMM_TIMELINE): clock64() stamps of CTA 0's producer,
MMA issuer and first epilogue warp for the first tiles of one GEMM shape. Usage: python profiles/gemm_timeline.py vhead"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
which = sys.argv[1] if len(sys.argv) > 1 else "vhead"
sys.argv = [sys.argv[0], which]
exec(open(os.path.join(ROOT, "profiles", "prof_gemm_one.py")).read())
from avi_talking_b200 import _lib  # noqa: E402
buf = (C.c_longlong * (3 * 64 * 8))()
assert _lib.load().avi_debug_timeline(buf) == 0
import numpy as np  # noqa: E402
t = np.array(buf[:], dtype=np.int64).reshape(3, 64, 8)
t0 = t[2, 0, 0]
print("tile | producer: start, first-empty | mma: start, tmem_empty, first-full, last-full, commit | epi: start, bias-bar, tmem_full, ld0, -, ld1, -, end   (clk since epilogue start of tile 0)")
for i in range(2, 14):
    print(i, (t[0, i, :2] - t0).tolist(), (t[1, i, :5] - t0).tolist(), (t[2, i, [0, 1, 2, 3, 5, 7]] - t0).tolist())
