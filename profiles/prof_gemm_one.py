"""One GEMM shape of the configs[1] step, launched 3 times (for `ncu --launch-skip 2 --launch-count 1`).
Usage: python profiles/prof_gemm_one.py {conv1|qkv|oproj|ffn1|ffn2|vhead}"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200 import ops  # noqa: E402
from avi_talking_b200.ops import ACT_GELU, ACT_NONE  # noqa: E402

which = sys.argv[1]
dev = "cuda"
B, T = 64, 249
M = B * T
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, dtype=torch.bfloat16, scale=1.0: (torch.randn(*s, device=dev, generator=g) * scale).to(dtype)  # noqa: E731

if which == "conv1":
    L, k, s, c = 31999, 3, 2, 512
    La, Lo = L + 1, (L - k) // s + 1
    Loa = Lo + (Lo & 1)
    h, w = rnd(B, La, c), rnd(c, k * c, scale=0.03)
    o = torch.empty((B, Loa, c), dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(h, w, None, o, batch=B, rows=Lo, N=c, K=k * c, act=ACT_GELU, conv_taps=k, conv_stride=s, a_ld=c,  # noqa: E731
                          a_batch_stride=La * c, a_rows_alloc=La, c_ld=c, c_batch_stride=Loa * c)
else:
    n, k, act, odt, res = {"qkv": (2304, 768, ACT_NONE, torch.bfloat16, False), "oproj": (768, 768, ACT_NONE, torch.float32, True),
                           "ffn1": (3072, 768, ACT_GELU, torch.bfloat16, False), "ffn2": (768, 3072, ACT_NONE, torch.float32, True),
                           "vhead": (15069, 192, ACT_NONE, torch.float32, False), "vheadp": (15069, 192, ACT_NONE, torch.float32, False),
                           "vhead64": (15069, 64, ACT_NONE, torch.float32, False)}[which]
    a, w, bias = rnd(M, k), rnd(n, k, scale=0.03), rnd(n, dtype=torch.float32)
    r = rnd(M, n, dtype=torch.float32) if res else None
    ld = 15072 if which in ("vheadp", "vhead64") else n      # 16-byte aligned vertex rows (ops.empty_rows)
    o = r if res else torch.empty((M, ld), dtype=odt, device=dev)   # residual GEMMs update the fp32 stream in place, as in the step
    fn = lambda: ops.gemm(a, w, bias, o, rows=M, N=n, K=k, act=act, residual=r, a_rows_alloc=M, c_ld=ld)  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record()
torch.cuda.synchronize()
print("ok", which, "%.4f ms" % (e0.elapsed_time(e1) / 10))
