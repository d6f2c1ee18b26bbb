import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
x = torch.empty(15936, 15069, device="cuda")
y = torch.empty(15936, 15069, device="cuda")
ms = t(lambda: x.fill_(1.0)); print("fill 960MB: %.3f ms  %.0f GB/s" % (ms, x.numel()*4/ms/1e6))
ms = t(lambda: y.copy_(x)); print("copy 960MB: %.3f ms  %.0f GB/s (r+w)" % (ms, 2*x.numel()*4/ms/1e6))
z = torch.empty(64, 32000, 512, device="cuda", dtype=torch.bfloat16)
ms = t(lambda: z.fill_(1.0)); print("fill 2.1GB: %.3f ms  %.0f GB/s" % (ms, z.numel()*2/ms/1e6))
a = torch.randn(15936, 64, device="cuda"); w = torch.randn(64, 15069, device="cuda")
ms = t(lambda: torch.mm(a, w, out=x)); print("cublas fp32 mm vhead: %.3f ms  %.0f GB/s" % (ms, x.numel()*4/ms/1e6))
ab = a.bfloat16(); wb = w.bfloat16(); xb = torch.empty(15936, 15069, device="cuda", dtype=torch.bfloat16)
ms = t(lambda: torch.mm(ab, wb, out=xb)); print("cublas bf16 mm vhead (bf16 out): %.3f ms  %.0f GB/s" % (ms, xb.numel()*2/ms/1e6))
