"""fp32 operands on the bf16 tensor cores as split terms (ops.FP32_GEMM = "x3": avi_split_bf16_terms + the tcgen05 GEMM with fp32
accumulation) against (1) an fp64 product of the same fp32 operands and (2) the reference-minted goldens of the fp32 mode
(tests/golden/w2v.npz, faceformer.npz). Measured: 5e-6 of max|C| per GEMM, 3e-5 relative over the wav2vec2 stack, 1e-6 m on the
vertices - between the bf16 mode (1e-2) and the CUDA-core fp32 mode (<= 1e-5, which stays the default of precision="fp32").
A six-term split measured no better (the tensor core's fp32 accumulation truncates), see ops.FP32_GEMM."""
import numpy as np
import pytest
import torch

from avi_talking_b200 import ops, synth
from oracle.make_golden import COL_STRIDE

from helpers import build_faceformer, build_wav2vec

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_split_terms_bit_exact_vs_torch():
    x = torch.from_numpy(np.random.default_rng(1).normal(size=(37, 128)).astype(np.float32)).cuda() * 3.0
    s0 = x.bfloat16()
    r1 = x - s0.float()
    s1 = r1.bfloat16()
    s2 = (r1 - s1.float()).bfloat16()
    terms = (s0, s1, s2)
    for pattern in ((0, 1, 0), (0, 0, 1), (0, 0, 0, 1, 1, 2), (0, 1, 2, 0, 1, 0)):
        got = ops.split_bf16_terms(x, pattern)
        want = torch.cat([terms[p] for p in pattern], dim=1)
        assert torch.equal(got, want), pattern
    # the three terms carry the value to fp32 precision
    assert ((s0.float() + s1.float()) + s2.float() - x).abs().max().item() <= 2.0 ** -24 * x.abs().max().item()


@pytest.mark.parametrize("mode,tol", [("x3", 3e-5)])
def test_split_gemm_vs_fp64(monkeypatch, mode, tol):
    """Plain and conv-mode (3 taps, stride 2, bias + GELU off) GEMMs: error relative to the fp64 product of the same fp32 operands;
    the CUDA-core fp32 kernel is the yardstick."""
    rng = np.random.default_rng(2)
    A = torch.from_numpy(rng.normal(size=(1000, 768)).astype(np.float32)).cuda()
    W = torch.from_numpy((rng.normal(size=(520, 768)) / 27.0).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.normal(size=(520,)).astype(np.float32)).cuda()
    want = A.double() @ W.double().t() + bias.double()
    errs = {}
    for m_ in ("simt", mode):
        monkeypatch.setattr(ops, "FP32_GEMM", m_)
        got = ops.linear(A, W, bias)
        errs[m_] = ((got.double() - want).abs().max() / want.abs().max()).item()
    print(f"plain GEMM max-err / max|ref|: simt {errs['simt']:.2e}, {mode} {errs[mode]:.2e}")
    assert errs[mode] <= tol
    # conv mode: [B, L, 128] time-major, 3 taps, stride 2 -> rows = (L - 3) // 2 + 1
    B, L, Cin, Cout, k, s_ = 3, 64, 128, 96, 3, 2
    x = torch.from_numpy(rng.normal(size=(B, L, Cin)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.normal(size=(Cout, Cin, k)) / 20.0).astype(np.float32)).cuda()
    Lo = (L - k) // s_ + 1
    wt = w.permute(0, 2, 1).reshape(Cout, k * Cin).contiguous()                          # tap-major [Cout, k * Cin]
    want = torch.nn.functional.conv1d(x.double().transpose(1, 2), w.double(), stride=s_).transpose(1, 2)
    monkeypatch.setattr(ops, "FP32_GEMM", mode)
    out = torch.empty((B, Lo, Cout), dtype=torch.float32, device="cuda")
    ops.gemm(x, wt, None, out, batch=B, rows=Lo, N=Cout, K=k * Cin, conv_taps=k, conv_stride=s_, a_batch_stride=L * Cin, a_rows_alloc=L,
             c_batch_stride=Lo * Cout)
    e = ((out.double() - want).abs().max() / want.abs().max()).item()
    print(f"conv-mode GEMM {mode}: {e:.2e}")
    assert e <= tol


@pytest.mark.parametrize("mode,tol", [("x3", 1e-4)])
def test_wav2vec2_fp32_mode_on_tensor_cores_matches_reference_golden(monkeypatch, golden, mode, tol):
    """The fp32 mode of Wav2Vec2Model.forward with every dense contraction on tcgen05 as split terms, against the goldens minted from
    the reference's own class: 3e-5 measured (the CUDA-core fp32 mode holds 1e-5, the bf16 mode 1e-2)."""
    monkeypatch.setattr(ops, "FP32_GEMM", mode)
    g = golden("w2v")
    m = build_wav2vec("fp32")
    a1 = synth.audio(2, 16000, seed=1234).cuda()
    e1 = relerr(m(a1, "vocaset").last_hidden_state.cpu(), torch.from_numpy(g["hs_1s"]))
    a4 = synth.audio(1, 64000, seed=1234).cuda()
    e4 = relerr(m(a4, "vocaset").last_hidden_state.cpu(), torch.from_numpy(g["hs_4s"]))
    print(f"fp32-mode wav2vec2 on tensor cores ({mode}): relative error {e1:.2e} (1 s x 2), {e4:.2e} (4 s)")
    assert e1 < tol and e4 < tol


def test_predict_fp32_mode_on_tensor_cores_matches_reference_golden(monkeypatch, golden):
    monkeypatch.setattr(ops, "FP32_GEMM", "x3")
    g = golden("faceformer")
    m = build_faceformer("fp32", fd=64, seed=10 + 64)
    a = synth.audio(1, 16000, seed=1234).cuda()
    emo = synth.fan_embeddings(24, seed=20)["emo"][None].cuda()
    v = m.predict_from_embeddings(a, emo)
    err = (v[0, :, ::COL_STRIDE].cpu() - torch.from_numpy(g["predict_fd64_sub"])).abs().max().item()
    print("fp32-mode predict on tensor cores (x3): max abs vertex error (m):", err)
    assert err < 1e-5
