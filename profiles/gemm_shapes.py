"""Per-shape timing of the tcgen05 GEMM (avi_gemm_bf16_tc) on every distinct contraction of the BASELINE configs[1] step
(64 clips x 10 s): conv1..6 as conv-mode GEMMs, feature projection, positional-conv block GEMM, the four encoder GEMMs and
the split-bf16 vertex head. CUDA events on the launching stream, `reps` back-to-back launches after 3 warm-ups; operands of
every shape exceed or rotate through more than the 126 MB L2 only for the large ones - small shapes are L2-resident, as they
are inside the real step (the producer kernel has just written them).

Usage (GPU box):  python profiles/gemm_shapes.py [clips] > gpurun_out/gemm_shapes.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from avi_talking_b200 import ops  # noqa: E402
from avi_talking_b200.ops import ACT_GELU, ACT_NONE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = 10
dev = "cuda"
T = 249
M = B * T
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


def timeit(fn):
    """`reps` launches captured in one CUDA graph (no host launch gaps between them), graph replayed 3 times after a warm-up."""
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


rows = []


def conv_case(name, L_in, k, s, cin=512, cout=512):
    La = L_in + (L_in & 1)
    Lo = (L_in - k) // s + 1
    Loa = Lo + (Lo & 1)
    h = rnd(B, La, cin)
    w = rnd(cout, k * cin, scale=0.03)
    o = torch.empty((B, Loa, cout), dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm(h, w, None, o, batch=B, rows=Lo, N=cout, K=k * cin, act=ACT_GELU, conv_taps=k, conv_stride=s, a_ld=cin,  # noqa: E731
                          a_batch_stride=La * cin, a_rows_alloc=La, c_ld=cout, c_batch_stride=Loa * cout)
    ms = timeit(fn)
    rows.append((name, B * Lo, cout, k * cin, 2.0 * B * Lo * cout * k * cin, ms))
    return Lo


def lin_case(name, m, n, k, act=ACT_NONE, out_dtype=torch.bfloat16, residual=False):
    a = rnd(m, k)
    w = rnd(n, k, scale=0.03)
    bias = rnd(n, dtype=torch.float32)
    res = rnd(m, n, dtype=torch.float32) if residual else None
    # residual GEMMs update the fp32 residual stream in place (TMA reduce-add epilogue), exactly as the encoder layer calls them
    o = res if residual else torch.empty((m, n), dtype=out_dtype, device=dev)
    fn = lambda: ops.gemm(a, w, bias, o, rows=m, N=n, K=k, act=act, residual=res, a_rows_alloc=m)  # noqa: E731
    ms = timeit(fn)
    rows.append((name, m, n, k, 2.0 * m * n * k, ms))


L = 31999
for i, (k, s) in enumerate(zip((3, 3, 3, 3, 2, 2), (2, 2, 2, 2, 2, 2)), start=1):
    L = conv_case(f"conv{i} (k={k}, s={s})", L, k, s)
lin_case("feature projection", M, 768, 512, out_dtype=torch.float32)
# positional conv: the whole grouped conv (16 groups x 48 channels, 128 taps) as the slab-resident implicit kernel; algorithmic FLOPs
Tp = T + 128
xpad = rnd(B, Tp, 768)
band = rnd(4, 128, 3, 96, 64, scale=0.01)
pb = rnd(768, dtype=torch.float32)
fn = lambda: ops.posconv_tc(xpad, band, pb, B, T, 16, 128)  # noqa: E731
rows.append(("pos-conv (implicit, whole layer; algorithmic FLOPs)", M, 768, 48 * 128, 2.0 * M * 768 * 48 * 128, timeit(fn)))
lin_case("encoder qkv", M, 2304, 768)
lin_case("encoder out-proj (+res, fp32)", M, 768, 768, out_dtype=torch.float32, residual=True)
lin_case("encoder ffn1 (GELU)", M, 3072, 768, act=ACT_GELU)
lin_case("encoder ffn2 (+res, fp32)", M, 768, 3072, out_dtype=torch.float32, residual=True)
# vertex head: split-bf16 operands, K = 3*64, ragged fp32 rows of 15069
a = rnd(M, 192)
w = rnd(15069, 192, scale=0.02)
bias = rnd(15069, dtype=torch.float32)
o = torch.empty((M, 15069), dtype=torch.float32, device=dev)
ms = timeit(lambda: ops.gemm(a, w, bias, o, rows=M, N=15069, K=192, a_rows_alloc=M))
rows.append(("vertex head (fp32 rows of 15069; HBM-write bound)", M, 15069, 192, 2.0 * M * 15069 * 192, ms))

print(f"{'shape':70s} {'M':>9s} {'N':>6s} {'K':>6s} {'ms':>8s} {'TFLOP/s':>8s} {'out GB/s':>9s}")
for name, m, n, k, fl, ms in rows:
    ob = m * n * (4 if ("fp32" in name or "projection" in name or "pos-conv" in name) else 2)
    print(f"{name:70s} {m:9d} {n:6d} {k:6d} {ms:8.4f} {fl / ms / 1e9:8.1f} {ob / ms / 1e6:9.1f}")
