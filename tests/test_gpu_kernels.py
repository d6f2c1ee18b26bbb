"""Per-kernel parity tests on the B200 (call through the C ABI via avi_talking_b200.ops).
References are float64 torch-on-CPU restatements of the same op; tolerances are written per test."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from avi_talking_b200 import ops as _ops
    return _ops


def _rng(seed):
    return np.random.default_rng(seed)


def _t(a, dev="cuda", dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)


def _act(x, act):
    if act == 1:
        return F.gelu(x)
    if act == 2:
        return F.relu(x)
    return x


def _gemm_ref(A, W, bias, act, residual):
    y = A.double() @ W.double().t()
    if bias is not None:
        y = y + bias.double()
    y = _act(y, act)
    if residual is not None:
        y = y + residual.double()
    return y


@pytest.mark.parametrize("rows,N,K,act,use_res", [(70, 100, 36, 0, False), (129, 64, 100, 2, True), (257, 130, 768, 1, True)])
def test_gemm_f32_plain(ops, rows, N, K, act, use_res):
    r = _rng(1)
    A, W, b = r.normal(size=(rows, K)), r.normal(size=(N, K)) / math.sqrt(K), r.normal(size=(N,))
    res = r.normal(size=(rows, N)) if use_res else None
    out = ops.linear(_t(A), _t(W), _t(b), act=act, residual=None if res is None else _t(res))
    ref = _gemm_ref(_t(A, "cpu"), _t(W, "cpu"), _t(b, "cpu"), act, None if res is None else _t(res, "cpu"))
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5


def _conv_ref(x, w, k, s):
    # x [B, L, C] time-major ; w [Cout, Cin, k]
    return F.gelu(F.conv1d(x.double().transpose(1, 2), w.double(), stride=s)).transpose(1, 2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_conv_mode(ops, dtype):
    r = _rng(2)
    B, L, Cin, Cout, k, s = 3, 301, 128, 192, 3, 2
    La = L + (L & 1)
    x = torch.zeros(B, La, Cin)
    x[:, :L] = torch.from_numpy(r.normal(size=(B, L, Cin)).astype(np.float32))
    w = torch.from_numpy((r.normal(size=(Cout, Cin, k)) / math.sqrt(Cin * k)).astype(np.float32))
    xq, wq = x.to(dtype).float(), w.to(dtype).float()
    Lo = (L - k) // s + 1
    Loa = Lo + (Lo & 1)
    out = torch.full((B, Loa, Cout), float("nan"), dtype=dtype, device="cuda")
    wp = w.permute(0, 2, 1).reshape(Cout, k * Cin).contiguous().to(dtype).cuda()
    ops.gemm(x.to(dtype).cuda(), wp, None, out, batch=B, rows=Lo, N=Cout, K=k * Cin, act=1, conv_taps=k, conv_stride=s,
             a_ld=Cin, a_batch_stride=La * Cin, a_rows_alloc=La, c_ld=Cout, c_batch_stride=Loa * Cout)
    ref = _conv_ref(xq[:, :L], wq, k, s)
    got = out[:, :Lo].float().cpu().double()
    tol = 2e-5 if dtype == torch.float32 else 2e-2  # bf16 output rounding (2^-9 relative on |y| <~ 4)
    assert (got - ref).abs().max().item() < tol


@pytest.mark.parametrize("rows,N,K,act,use_res,out_dt,out2", [
    (300, 512, 128, 0, False, torch.float32, False),
    (1000, 768, 768, 1, False, torch.bfloat16, False),
    (515, 2304, 768, 0, True, torch.float32, True),
    (390, 15069, 64, 0, False, torch.float32, False),   # vertex head: ragged N, unaligned rows
    (128, 256, 3072, 0, True, torch.float32, False),
])
def test_gemm_bf16_tc(ops, rows, N, K, act, use_res, out_dt, out2):
    r = _rng(3)
    A = torch.from_numpy(r.normal(size=(rows, K)).astype(np.float32)).bfloat16()
    W = torch.from_numpy((r.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32)).bfloat16()
    b = torch.from_numpy(r.normal(size=(N,)).astype(np.float32))
    res = torch.from_numpy(r.normal(size=(rows, N)).astype(np.float32)) if use_res else None
    o = ops.linear(A.cuda(), W.cuda(), b.cuda(), act=act, residual=None if res is None else res.cuda(), out_dtype=out_dt,
                   out2_dtype=(torch.bfloat16 if out_dt == torch.float32 else torch.float32) if out2 else None)
    ref = _gemm_ref(A.float(), W.float(), b, act, res)
    outs = o if isinstance(o, tuple) else (o,)
    for t in outs:
        tol = 1e-4 if t.dtype == torch.float32 else 4e-2
        err = (t.float().cpu().double() - ref).abs().max().item()
        assert err < tol, (t.dtype, err)


def test_gemm_bf16_tc_many_tiles_persistent(ops):
    """More tiles than SMs and >2 tiles per CTA: exercises the smem ring wrap and both TMEM accumulator stages."""
    r = _rng(4)
    rows, N, K = 128 * 40, 1024, 256   # 160 tiles... x4 n-tiles = 160 -> >148
    A = torch.from_numpy(r.normal(size=(rows, K)).astype(np.float32)).bfloat16()
    W = torch.from_numpy((r.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32)).bfloat16()
    o = ops.linear(A.cuda(), W.cuda(), None, out_dtype=torch.float32)
    ref = A.float().double() @ W.float().double().t()
    assert (o.cpu().double() - ref).abs().max().item() < 1e-4
    rows = 128 * 148 * 3 + 17
    A = torch.from_numpy(r.normal(size=(rows, 64)).astype(np.float32)).bfloat16()
    W = torch.from_numpy((r.normal(size=(256, 64)) / 8).astype(np.float32)).bfloat16()
    o = ops.linear(A.cuda(), W.cuda(), None, out_dtype=torch.float32)
    ref = A.float().double() @ W.float().double().t()
    assert (o.cpu().double() - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("rows,N,K,act,out_dt", [
    (256 * 51 + 9, 768, 192, 0, torch.float32),     # 52 m-tiles x 3 n-tiles: even super-tile count, ragged last rows
    (256 * 62 + 64, 800, 128, 1, torch.bfloat16),    # 63 m-tiles (odd: the last super-tile has an idle pair), narrow last n-tile (32 columns)
    (256 * 75, 256, 64, 0, torch.float32),           # one n-tile, 75 m-tiles -> 38 super-tiles on <= 37 clusters, single k-block
])
def test_gemm_bf16_tc_two_pair_clusters_multicast(ops, rows, N, K, act, out_dt):
    """More than one wave of tiles: the opt-in 4-CTA cluster kernel (two pairs on consecutive m-tiles, W quarters multicast)
    against the reference product AND bit for bit against the default pair kernel."""
    from avi_talking_b200 import _lib
    r = _rng(41)
    A = torch.from_numpy(r.normal(size=(rows, K)).astype(np.float32)).bfloat16()
    W = torch.from_numpy((r.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32)).bfloat16()
    b = torch.from_numpy(r.normal(size=(N,)).astype(np.float32))
    o_pair = ops.linear(A.cuda(), W.cuda(), b.cuda(), act=act, out_dtype=out_dt)
    old = _lib.load().avi_gemm_set_multicast(1)
    try:
        o = ops.linear(A.cuda(), W.cuda(), b.cuda(), act=act, out_dtype=out_dt)
    finally:
        _lib.load().avi_gemm_set_multicast(old)
    ref = _gemm_ref(A.float(), W.float(), b, act, None)
    tol = 1e-4 if out_dt == torch.float32 else 4e-2
    assert (o.float().cpu().double() - ref).abs().max().item() < tol
    assert torch.equal(o, o_pair)


@pytest.mark.parametrize("multicast", [0, 1])
def test_gemm_conv_mode_many_tiles_batched(ops, multicast):
    """Strided conv as a GEMM over 5 clips x 13 m-tiles x 2 n-tiles (130 tiles: the cluster kernel, super-tiles straddling clips)."""
    r = _rng(42)
    B, L, Cin, Cout, k, s = 5, 2 * 256 * 13 - 100, 64, 512, 3, 2
    La = L + (L & 1)
    x = torch.zeros(B, La, Cin)
    x[:, :L] = torch.from_numpy(r.normal(size=(B, L, Cin)).astype(np.float32))
    w = torch.from_numpy((r.normal(size=(Cout, Cin, k)) / math.sqrt(Cin * k)).astype(np.float32))
    dtype = torch.bfloat16
    xq, wq = x.to(dtype).float(), w.to(dtype).float()
    Lo = (L - k) // s + 1
    Loa = Lo + (Lo & 1)
    out = torch.full((B, Loa, Cout), float("nan"), dtype=dtype, device="cuda")
    wp = w.permute(0, 2, 1).reshape(Cout, k * Cin).contiguous().to(dtype).cuda()
    from avi_talking_b200 import _lib
    old = _lib.load().avi_gemm_set_multicast(multicast)
    try:
        ops.gemm(x.to(dtype).cuda(), wp, None, out, batch=B, rows=Lo, N=Cout, K=k * Cin, act=1, conv_taps=k, conv_stride=s,
                 a_ld=Cin, a_batch_stride=La * Cin, a_rows_alloc=La, c_ld=Cout, c_batch_stride=Loa * Cout)
    finally:
        _lib.load().avi_gemm_set_multicast(old)
    ref = _conv_ref(xq[:, :L], wq, k, s)
    got = out[:, :Lo].float().cpu().double()
    assert (got - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("out_dt", [torch.float32, torch.bfloat16])
def test_conv0_gn_gelu(ops, out_dt):
    r = _rng(5)
    B, n, C = 2, 4000, 512
    x = torch.from_numpy(r.normal(size=(B, n)).astype(np.float32))
    w = torch.from_numpy((r.normal(size=(C, 1, 10)) * 0.4).astype(np.float32))
    g = torch.from_numpy((1 + 0.1 * r.normal(size=C)).astype(np.float32))
    bt = torch.from_numpy((0.1 * r.normal(size=C)).astype(np.float32))
    L = (n - 10) // 5 + 1
    La = L + (L & 1)
    out = torch.zeros(B, La, C, dtype=out_dt, device="cuda")
    ops.conv0_gn_gelu(x.cuda(), w.reshape(C, 10).cuda(), g.cuda(), bt.cuda(), out, La * C)
    ref = F.gelu(F.group_norm(F.conv1d(x.double()[:, None], w.double(), stride=5), C, g.double(), bt.double(), 1e-5)).transpose(1, 2)
    tol = 2e-5 if out_dt == torch.float32 else 3e-2
    assert (out[:, :L].float().cpu().double() - ref).abs().max().item() < tol


@pytest.mark.parametrize("T_in,T_out", [(49, 24), (199, 99), (499, 249), (49, 20), (10, 1)])
def test_lerp_layernorm(ops, T_in, T_out):
    r = _rng(6)
    B, C = 2, 512
    x = torch.from_numpy(r.normal(size=(B, T_in + 1, C)).astype(np.float32))
    g = torch.from_numpy((1 + 0.1 * r.normal(size=C)).astype(np.float32))
    bt = torch.from_numpy((0.1 * r.normal(size=C)).astype(np.float32))
    o32, o16 = ops.lerp_layernorm(x.cuda(), (T_in + 1) * C, B, T_in, T_out, g.cuda(), bt.cuda(), True, True)
    y = F.interpolate(x[:, :T_in].transpose(1, 2), size=T_out, align_corners=True, mode="linear").transpose(1, 2)
    ref = F.layer_norm(y, (C,), g, bt, 1e-5).reshape(B * T_out, C)
    assert (o32.cpu() - ref).abs().max().item() < 1e-5
    assert (o16.float().cpu() - ref).abs().max().item() < 3e-2


def test_layernorm_with_residual(ops):
    r = _rng(7)
    x = torch.from_numpy(r.normal(size=(77, 768)).astype(np.float32))
    rs = torch.from_numpy(r.normal(size=(77, 768)).astype(np.float32))
    g = torch.from_numpy((1 + 0.1 * r.normal(size=768)).astype(np.float32))
    bt = torch.from_numpy((0.1 * r.normal(size=768)).astype(np.float32))
    o, _ = ops.layernorm(x.cuda(), g.cuda(), bt.cuda())
    assert (o.cpu() - F.layer_norm(x, (768,), g, bt, 1e-5)).abs().max().item() < 1e-5
    o, o16 = ops.layernorm(x.cuda(), g.cuda(), bt.cuda(), res=rs.cuda(), want_bf16=True)
    ref = F.layer_norm(x + rs, (768,), g, bt, 1e-5)
    assert (o.cpu() - ref).abs().max().item() < 1e-5
    assert (o16.float().cpu() - ref).abs().max().item() < 3e-2
    x64 = torch.from_numpy(r.normal(size=(5, 64)).astype(np.float32))
    o, _ = ops.layernorm(x64.cuda(), g[:64].cuda(), bt[:64].cuda())
    assert (o.cpu() - F.layer_norm(x64, (64,), g[:64], bt[:64], 1e-5)).abs().max().item() < 1e-5


@pytest.mark.parametrize("T", [24, 99, 130])
def test_posconv_ln(ops, T):
    r = _rng(8)
    B, C, G, K = 2, 768, 16, 128
    x = torch.from_numpy(r.normal(size=(B, T, C)).astype(np.float32))
    w = torch.from_numpy((r.normal(size=(C, C // G, K)) / math.sqrt(48 * 128) * 2).astype(np.float32))
    cb = torch.from_numpy((0.02 * r.normal(size=C)).astype(np.float32))
    g = torch.from_numpy((1 + 0.1 * r.normal(size=C)).astype(np.float32))
    bt = torch.from_numpy((0.1 * r.normal(size=C)).astype(np.float32))
    wp = w.reshape(G, 48, 48, K).permute(0, 3, 2, 1).contiguous()
    o32, o16 = ops.posconv_ln(x.reshape(B * T, C).cuda(), wp.cuda(), cb.cuda(), g.cuda(), bt.cuda(), B, T, G, K, True)
    pc = F.conv1d(x.double().transpose(1, 2), w.double(), cb.double(), padding=K // 2, groups=G)[:, :, :-1]
    ref = F.layer_norm(x.double() + F.gelu(pc).transpose(1, 2), (C,), g.double(), bt.double(), 1e-5).reshape(B * T, C)
    assert (o32.cpu().double() - ref).abs().max().item() < 2e-5
    assert (o16.float().cpu().double() - ref).abs().max().item() < 3e-2


@pytest.mark.parametrize("rows,K", [(7000, 768), (15936, 3072), (20000, 512)])
def test_gemm_inplace_residual_many_tiles(ops, rows, K):
    """The in-place residual GEMM (C += A W^T + bias through TMA reduce-add) with more tiles than CTA pairs: every output element is
    produced by ONE reduce-add, so the result is bit-reproducible run to run (a stream-K split of the tail wave was measured in round 2
    and was slower - profiles/r2/README.md - so whole tiles stay the schedule)."""
    r = _rng(100 + K)
    N = 768
    A = torch.from_numpy(r.normal(size=(rows, K)).astype(np.float32)).bfloat16().cuda()
    W = torch.from_numpy((r.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32)).bfloat16().cuda()
    bias = torch.from_numpy(r.normal(size=N).astype(np.float32)).cuda()
    res = torch.from_numpy(r.normal(size=(rows, N)).astype(np.float32)).cuda()
    ref = (A.float().double() @ W.float().double().t() + bias.double() + res.double()).cpu()

    def run():
        c = res.clone()
        ops.gemm(A, W, bias, c, rows=rows, N=N, K=K, residual=c, a_rows_alloc=rows)
        return c

    c0, c1 = run(), run()
    assert torch.equal(c0, c1)
    assert (c0.cpu().double() - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("B,T", [(2, 24), (3, 249), (1, 256), (2, 300), (5, 99)])
def test_posconv_tensor_core_implicit(ops, B, T):
    """avi_w2v_posconv_tc (slab-resident implicit grouped conv on tcgen05, banded weight, two tap halves meeting by reduce-add) ==
    fp64 torch conv1d(groups=16, padding=64)[..., :-1] on the bf16-rounded operands. Covers one / two time tiles per clip, T below
    and above one CTA's 128 rows, and more units than one wave of pairs would need zero-padding for."""
    from avi_talking_b200.wav2vec import pack_posconv_band
    r = _rng(80 + T)
    C, G, K = 768, 16, 128
    x = torch.from_numpy(r.normal(size=(B, T, C)).astype(np.float32))
    w = torch.from_numpy((r.normal(size=(C, C // G, K)) / math.sqrt(48 * 128) * 2).astype(np.float32))
    cb = torch.from_numpy((0.02 * r.normal(size=C)).astype(np.float32))
    xpad = ops.pad_cast_bf16(x.reshape(B * T, C).cuda(), B, T, K // 2, T + K)
    assert torch.equal(xpad[:, K // 2:K // 2 + T].cpu(), x.bfloat16()) and float(xpad[:, :K // 2].abs().max()) == 0.0
    band = pack_posconv_band(w.cuda(), G)
    pc = ops.posconv_tc(xpad, band, cb.cuda(), B, T, G, K)
    pc2 = ops.posconv_tc(xpad, band, cb.cuda(), B, T, G, K)
    assert torch.equal(pc, pc2)                                   # two commutative additions onto zero: run-to-run identical
    ref = F.conv1d(x.bfloat16().double().transpose(1, 2), w.bfloat16().double(), cb.double(), padding=K // 2, groups=G)[:, :, :-1]
    err = (pc.cpu().double().view(B, T, C) - ref.transpose(1, 2)).abs().max().item()
    assert err < 2e-4, err                                        # fp32 accumulation of 6144 bf16 products of magnitude ~ 0.03


@pytest.mark.parametrize("T,dtype", [(24, torch.float32), (249, torch.float32), (300, torch.bfloat16), (129, torch.bfloat16),
                                     (24, torch.bfloat16), (128, torch.bfloat16), (249, torch.bfloat16), (256, torch.bfloat16)])
def test_mha(ops, T, dtype):
    r = _rng(9)
    B, H, D = 2, 12, 64
    qkv = torch.from_numpy(r.normal(size=(B, T, 3 * H * D)).astype(np.float32)).to(dtype)
    out = ops.mha(qkv.cuda().reshape(B * T, -1), B, T, H, D, D ** -0.5)
    q, k, v = [t.reshape(B, T, H, D).transpose(1, 2) for t in qkv.float().double().split(H * D, dim=-1)]
    a = torch.softmax(q @ k.transpose(2, 3) * D ** -0.5, -1)
    ref = (a @ v).transpose(1, 2).reshape(B * T, H * D)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert (out.float().cpu().double() - ref).abs().max().item() < tol


def test_cast_bf16(ops):
    x = torch.from_numpy(_rng(10).normal(size=(1031,)).astype(np.float32))
    assert torch.equal(ops.cast_bf16(x.cuda()).cpu(), x.bfloat16())


@pytest.mark.parametrize("n", [4000, 16000, 1290])
def test_conv0_gn_gelu_tensor_core(ops, n):
    """conv0 as a split-bf16 MMA with analytic GroupNorm statistics == fp64 torch conv + group_norm + gelu (bf16 output rounding)."""
    r = _rng(6)
    B, C = 3, 512
    x = torch.from_numpy(r.normal(size=(B, n)).astype(np.float32))
    x[1] *= 3.0
    x[2] += 0.5          # non-zero mean: exercises the mean term of the analytic statistics
    w = torch.from_numpy((r.normal(size=(C, 1, 10)) * 0.4).astype(np.float32))
    g = torch.from_numpy((1 + 0.1 * r.normal(size=C)).astype(np.float32))
    bt = torch.from_numpy((0.1 * r.normal(size=C)).astype(np.float32))
    L = (n - 10) // 5 + 1
    La = L + (L & 1)
    out = torch.zeros(B, La, C, dtype=torch.bfloat16, device="cuda")
    w2 = w.reshape(C, 10).cuda()
    ops.conv0_gn_gelu_tc(x.cuda(), w2, ops.conv0_pack_tc(w2), g.cuda(), bt.cuda(), out, La * C)
    ref = F.gelu(F.group_norm(F.conv1d(x.double()[:, None], w.double(), stride=5), C, g.double(), bt.double(), 1e-5)).transpose(1, 2)
    got = out[:, :L].float().cpu().double()
    err = (got - ref).abs()
    # bf16 output rounding is 2^-9 relative; the arithmetic before it is fp32-accurate
    assert (err / (1.0 + ref.abs())).max().item() < 6e-3
    assert err.mean().item() < 1e-3


@pytest.mark.parametrize("rows,N,K,ld", [
    (390, 15069, 64, 15072),     # vertex rows on a 16-byte aligned pitch: TMA-store epilogue, last n-tile 221 -> MMA N = 224
    (517, 192, 128, 768),        # positional-conv block: N = 192 < one n-tile, output is a column window of a wider buffer
    (260, 96, 64, 96),           # N = 96: narrow MMA (n_eff 96), bf16-friendly pitch
    (130, 45, 64, 768),          # unaligned window inside a wider buffer: must fall back to per-thread stores (neighbours intact)
])
def test_gemm_bf16_tc_tma_store_windows_and_narrow_tiles(ops, rows, N, K, ld):
    """Nothing outside the [rows, N] window is touched - except the row's own padding when the pitch ends at the 16-byte granule of
    column N-1 (the TMA unit clips in 16-byte granules: include/avi_b200.h) - and narrow last n-tiles are exact."""
    r = _rng(7)
    A = torch.from_numpy(r.normal(size=(rows, K)).astype(np.float32)).bfloat16()
    W = torch.from_numpy((r.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32)).bfloat16()
    b = torch.from_numpy(r.normal(size=(N,)).astype(np.float32))
    ref = _gemm_ref(A.float(), W.float(), b, 0, None)
    for dt, tol in ((torch.float32, 1e-4), (torch.bfloat16, 4e-2)):
        buf = torch.full((rows + 3, ld), 7.0, dtype=dt, device="cuda")          # sentinel everywhere
        ops.gemm(A.cuda(), W.cuda(), b.cuda(), buf, rows=rows, N=N, K=K, a_rows_alloc=rows, c_ld=ld)
        got = buf.float().cpu()
        assert (got[:rows, :N].double() - ref).abs().max().item() < tol
        es = 4 if dt == torch.float32 else 2
        pad_end = -(-N * es // 16) * 16 // es                      # end of the 16-byte granule holding column N-1
        own_padding = pad_end == ld
        assert torch.all(got[rows:] == 7.0)
        assert torch.all(got[:rows, (pad_end if own_padding else N):] == 7.0)
        if own_padding:
            assert torch.all((got[:rows, N:pad_end] == 7.0) | (got[:rows, N:pad_end] == 0.0))


def test_gemm_bf16_tc_inplace_residual_reduce_add(ops):
    """residual is out: the epilogue updates the fp32 residual stream in place through TMA reduce-add; same numbers as the
    out-of-place residual epilogue (one fp32 addition either way)."""
    r = _rng(8)
    rows, N, K = 700, 768, 768
    A = torch.from_numpy(r.normal(size=(rows, K)).astype(np.float32)).bfloat16().cuda()
    W = torch.from_numpy((r.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32)).bfloat16().cuda()
    b = torch.from_numpy(r.normal(size=(N,)).astype(np.float32)).cuda()
    res = torch.from_numpy(r.normal(size=(rows, N)).astype(np.float32)).cuda()
    sep = ops.linear(A, W, b, residual=res, out_dtype=torch.float32)
    inplace = res.clone()
    out = ops.linear(A, W, b, residual=inplace, out_dtype=torch.float32, out=inplace)
    assert out.data_ptr() == inplace.data_ptr()
    ref = _gemm_ref(A.float().cpu(), W.float().cpu(), b.cpu(), 0, res.cpu())
    assert (inplace.cpu().double() - ref).abs().max().item() < 1e-4
    assert torch.equal(inplace, sep)


@pytest.mark.parametrize("rows,N,K,act,use_res", [(300, 512, 128, 0, False), (1000, 768, 96, 1, False), (515, 192, 2304, 0, True),
                                                  (390, 15069, 64, 0, False)])
def test_gemm_tf32_tc(ops, rows, N, K, act, use_res):
    """tcgen05.mma.kind::tf32 on fp32 operands (same pipeline as the bf16 GEMM, 32 elements per k-block): compared with the exact
    product of the TF32-truncated operands (1e-4) and with the fp32 product (10-bit significand: 2e-3 on O(1) outputs)."""
    r = _rng(9)
    A = torch.from_numpy(r.normal(size=(rows, K)).astype(np.float32))
    W = torch.from_numpy((r.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32))
    b = torch.from_numpy(r.normal(size=(N,)).astype(np.float32))
    res = torch.from_numpy(r.normal(size=(rows, N)).astype(np.float32)) if use_res else None
    ld = 15072 if N == 15069 else N
    out = torch.empty((rows, ld), dtype=torch.float32, device="cuda")
    ops.gemm(A.cuda(), W.cuda(), b.cuda(), out, rows=rows, N=N, K=K, act=act, residual=None if res is None else res.cuda(),
             a_rows_alloc=rows, c_ld=ld, tf32=True)
    got = out[:, :N].cpu().double()

    def trunc(t):   # TF32 keeps 10 explicit significand bits: the hardware ignores the low 13 bits of the fp32 word
        return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)

    ref_t = _gemm_ref(trunc(A), trunc(W), b, act, res)
    ref = _gemm_ref(A, W, b, act, res)
    e_t, e = (got - ref_t).abs().max().item(), (got - ref).abs().max().item()
    print(f"tf32 GEMM {rows}x{N}x{K}: max err vs truncated-operand product {e_t:.2e}, vs fp32 product {e:.2e}")
    assert e_t < 1e-4 and e < 1e-2


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_gemm_2d_taps_is_a_3x3_convolution(ops, mode):
    """AviGemmArgs.conv_taps_x / conv_row_pitch: a 3x3 / pad 1 convolution over zero-bordered NHWC lines of W+2 pixels as ONE contraction
    with 9 taps, against F.conv2d."""
    import torch.nn.functional as F
    r = _rng(11)
    N, H, W, Cin, Cout = 3, 13, 10, 64, 48
    x = torch.from_numpy(r.normal(size=(N, Cin, H, W)).astype(np.float32))
    w = torch.from_numpy((r.normal(size=(Cout, Cin, 3, 3)) / math.sqrt(9 * Cin)).astype(np.float32))
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    if mode == "bf16":
        x, w = x.bfloat16().float(), w.bfloat16().float()
    Wp = W + 2
    a = torch.zeros(N * (H + 2) * Wp + 2, Cin)
    a[: N * (H + 2) * Wp].view(N, H + 2, Wp, Cin)[:, 1:H + 1, 1:W + 1] = x.permute(0, 2, 3, 1)
    wm = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out = torch.empty(N * H * Wp, Cout, dtype=torch.float32, device="cuda")
    ops.gemm(a.to(dt).cuda(), wm.to(dt).cuda(), None, out, batch=N, rows=H * Wp, N=Cout, K=9 * Cin, conv_taps=9, conv_stride=1,
             conv_taps_x=3, conv_row_pitch=Wp, a_ld=Cin, a_batch_stride=(H + 2) * Wp * Cin, a_rows_alloc=(H + 2) * Wp + 2, c_ld=Cout,
             c_batch_stride=H * Wp * Cout, tf32=(mode == "tf32"))
    got = out.cpu().view(N, H, Wp, Cout)[:, :, :W].permute(0, 3, 1, 2)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    tol = {"fp32": 1e-5, "tf32": 5e-3, "bf16": 1e-4}[mode]      # bf16 operands are exact here (pre-rounded): fp32 accumulation only
    assert (got.double() - ref).abs().max().item() < tol


@pytest.mark.gpu
def test_compact_vertex_sink_roundtrip(ops):
    """avi_pack_disp_f16 on padded-pitch vertex rows and frontend.CompactVertexSink: fp16 displacement out, fp32 vertices back on
    the host within 2^-11 of the displacement (opt-in, outside the fp32 contract)."""
    from avi_talking_b200.frontend import CompactVertexSink
    r = _rng(51)
    rows, Cc, ld = 37, 15069, 15072
    tpl = torch.from_numpy(r.normal(size=(Cc,)).astype(np.float32) * 0.1)
    disp = torch.from_numpy(r.normal(size=(rows, Cc)).astype(np.float32) * 1e-2)
    buf = torch.zeros(rows, ld)
    buf[:, :Cc] = tpl + disp
    v = buf.cuda()[:, :Cc]                                   # non-contiguous rows, as the drop-in returns them
    h = ops.pack_disp_f16(v, tpl.cuda())
    assert h.dtype == torch.float16 and h.shape == (rows, Cc)
    want = (buf[:, :Cc] - tpl)
    assert (h.float().cpu() - want).abs().max().item() <= want.abs().max().item() * 2.0 ** -11 + 1e-9
    sink = CompactVertexSink(("verts",), {"verts": tpl})
    assert sink.push({"verts": v.view(1, rows, Cc)}) is None
    back = sink.unpack(sink.flush())["verts"]
    assert back.shape == (1, rows, Cc) and (back[0] - buf[:, :Cc]).abs().max().item() < 5e-5


def test_dynamic_tile_scheduler_is_bit_identical_to_the_static_walk(ops):
    """Cluster launch control (one cluster per tile in the grid, the resident CTAs take the pending ones) against the static
    tile = cta + k * ctas walk: the GEMM (plain, narrow last n-tile, conv mode, in-place reduce-add) and conv0, bit for bit."""
    r = _rng(11)
    prev = ops.set_dynamic_tiles(3)
    try:
        def both(fn):
            outs = []
            for mask in (3, 0):
                ops.set_dynamic_tiles(mask)
                outs.append(fn())
            torch.cuda.synchronize()
            return outs

        for rows, N, K, act, out_dt in [(256 * 62 + 64, 800, 128, 1, torch.bfloat16), (128 * 148 * 3 + 17, 256, 64, 0, torch.float32),
                                        (15936, 2304, 768, 0, torch.bfloat16), (300, 64, 64, 0, torch.float32)]:
            A = _t(r.normal(size=(rows, K)), dt=torch.bfloat16)
            W = _t(r.normal(size=(N, K)) / math.sqrt(K), dt=torch.bfloat16)
            b = _t(r.normal(size=N))
            d, s_ = both(lambda: ops.linear(A, W, b, act=act, out_dtype=out_dt))
            assert torch.equal(d, s_), (rows, N, K)
        # in-place residual (TMA reduce-add): the same tiles land on the same addresses whoever computes them
        A = _t(r.normal(size=(15936, 768)), dt=torch.bfloat16)
        W = _t(r.normal(size=(768, 768)) / 27.0, dt=torch.bfloat16)
        base = _t(r.normal(size=(15936, 768)))

        def inplace():
            h = base.clone()
            ops.linear(A, W, None, residual=h, out_dtype=torch.float32, out=h)
            return h
        d, s_ = both(inplace)
        assert torch.equal(d, s_)
        # conv mode, batched
        B, L, Cin, Cout, k, st = 5, 2000, 128, 512, 3, 2
        x = _t(r.normal(size=(B, L, Cin)), dt=torch.bfloat16)
        w = _t(r.normal(size=(Cout, k * Cin)) / 20.0, dt=torch.bfloat16)
        Lo = (L - k) // st + 1

        def conv():
            out = torch.empty((B, Lo, Cout), dtype=torch.bfloat16, device="cuda")
            ops.gemm(x, w, None, out, batch=B, rows=Lo, N=Cout, K=k * Cin, conv_taps=k, conv_stride=st, a_batch_stride=L * Cin,
                     a_rows_alloc=L, c_batch_stride=Lo * Cout, act=1)
            return out
        d, s_ = both(conv)
        assert torch.equal(d, s_)
        # conv0 (one CTA per 128-row tile)
        Bc, n, C = 5, 48000, 512
        xa = _t(r.normal(size=(Bc, n)))
        w0 = _t(r.normal(size=(C, 10)) * 0.4)
        g, bt = _t(1 + 0.1 * r.normal(size=C)), _t(0.1 * r.normal(size=C))
        L0 = (n - 10) // 5 + 1
        La = L0 + (L0 & 1)
        wp = ops.conv0_pack_tc(w0)

        def conv0():
            out = torch.zeros(Bc, La, C, dtype=torch.bfloat16, device="cuda")
            ops.conv0_gn_gelu_tc(xa, w0, wp, g, bt, out, La * C)
            return out
        d, s_ = both(conv0)
        assert torch.equal(d, s_)
    finally:
        ops.set_dynamic_tiles(prev)
