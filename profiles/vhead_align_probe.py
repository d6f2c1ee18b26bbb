import os, sys
sys.path.insert(0, os.getcwd())
import torch
from avi_talking_b200 import ops
M=15936
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, dtype=torch.bfloat16, scale=1.0: (torch.randn(*s, device="cuda", generator=g) * scale).to(dtype)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/(3*reps)
a = rnd(M, 192)
for N, ld in ((15069, 15069), (15069, 15072), (15072, 15072), (15104, 15104)):
    w = rnd(N, 192, scale=0.02); bias = rnd(N, dtype=torch.float32)
    o = torch.empty((M, ld), dtype=torch.float32, device="cuda")
    ms = timeit(lambda: ops.gemm(a, w, bias, o, rows=M, N=N, K=192, a_rows_alloc=M, c_ld=ld))
    print(f"N={N} c_ld={ld}: {ms:.4f} ms  {M*N*4/ms/1e6:.0f} GB/s")
