"""Yardstick only (never on the product path): cuBLAS bf16 (torch.matmul) next to avi_gemm_bf16_tc on the plain-GEMM shapes of the
configs[1] step, alternating in one process so both see the same clocks. Shows how much of the gap to the 8192^3 'measured peak' is the
SHAPE (L2-bound operand stream, tile quantisation) rather than the kernel. Usage: python profiles/cublas_yardstick.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, device=dev, generator=g) * scale).bfloat16()  # noqa: E731


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps)


print(f"{'shape':34s} {'M':>8s} {'N':>6s} {'K':>6s} {'ours ms':>9s} {'TF/s':>7s} {'cuBLAS ms':>10s} {'TF/s':>7s}")
for name, M, N, K in (("conv1-like (plain GEMM)", 1023936, 512, 1536), ("conv3-like", 255936, 512, 1536), ("encoder qkv", 15936, 2304, 768),
                      ("encoder out-proj (bf16 out)", 15936, 768, 768), ("encoder ffn1 (no GELU)", 15936, 3072, 768),
                      ("encoder ffn2 (bf16 out)", 15936, 768, 3072), ("8192^3", 8192, 8192, 8192)):
    a, w = rnd(M, K), rnd(N, K, scale=0.03)
    o = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    ours = timeit(lambda: ops.gemm(a, w, None, o, rows=M, N=N, K=K, a_rows_alloc=M))
    wt = w.t()
    ref = timeit(lambda: torch.matmul(a, wt, out=o))
    fl = 2.0 * M * N * K
    print(f"{name:34s} {M:8d} {N:6d} {K:6d} {ours:9.4f} {fl / ours / 1e9:7.1f} {ref:10.4f} {fl / ref / 1e9:7.1f}")
