import sys, math, torch, numpy as np
sys.path.insert(0, "/root/repo")
from avi_talking_b200 import ops
rows, N, K, ld = 390, 15069, 64, 15072
g = torch.Generator().manual_seed(0)
A = torch.randn(rows, K, generator=g).bfloat16().cuda(); W = (torch.randn(N, K, generator=g) / 8).bfloat16().cuda(); b = torch.randn(N, generator=g).cuda()
for dt in (torch.float32, torch.bfloat16):
    for N2, ld2 in ((15069, 15072), (221, 224), (13, 16), (45, 48)):
        buf = torch.full((rows + 3, ld2), 7.0, dtype=dt, device="cuda")
        ops.gemm(A, W[:N2].contiguous(), b[:N2].contiguous(), buf, rows=rows, N=N2, K=K, a_rows_alloc=rows, c_ld=ld2)
        got = buf.float().cpu()
        print(dt, N2, ld2, "pad cols untouched:", bool(torch.all(got[:rows, N2:] == 7.0)), "tail rows untouched:", bool(torch.all(got[rows:] == 7.0)), got[0, N2-2:].tolist())
