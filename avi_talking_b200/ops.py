"""Thin torch-tensor wrappers over the C ABI (include/avi_b200.h). torch is used for device memory and the
current stream only; every computation below runs in libavi_b200.so. No fallbacks."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_RELU, DT_BF16, DT_F32, AviDecoderWeights, AviGemmArgs, AviPriorNet  # noqa: F401


# bench.py's roofline pass: when set to a list, every launch made through `_timed` appends
# (kernel name, start event, end event, algorithmic work) with CUDA events recorded on the launching stream.
PROFILE = None

# incremented whenever a kernel of this library rewrites model weights in place (train.FlatAdam.step): part of every pack key
WEIGHT_EPOCH = 0


class _timed:
    def __init__(self, name, work):
        self.name, self.work = name, work

    def __enter__(self):
        if PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, self.e0, self.e1, self.work))
        return False


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DT_F32
    if t.dtype == torch.bfloat16:
        return DT_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("avi_talking_b200 ops need CUDA tensors (there is no CPU path)")


def cast_bf16(src: torch.Tensor, out=None) -> torch.Tensor:
    _need_cuda(src)
    src = src.contiguous().float()
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device) if out is None else out
    _lib.check(_lib.load().avi_cast_f32_to_bf16(_ptr(src), _ptr(dst), C.c_int64(src.numel()), _stream()), "avi_cast_f32_to_bf16")
    return dst


def pad_cast_bf16(x, B, T, front, rows_out):
    _need_cuda(x)
    Cc = x.shape[-1]
    out = torch.empty((B, rows_out, Cc), dtype=torch.bfloat16, device=x.device)
    with _timed("pad_cast", 0.0):
        _lib.check(_lib.load().avi_pad_cast_bf16(_ptr(x), _ptr(out), C.c_int32(B), C.c_int32(T), C.c_int32(Cc), C.c_int32(front),
                                                 C.c_int32(rows_out), _stream()), "avi_pad_cast_bf16")
    return out


def split_bf16x3(x2d):
    _need_cuda(x2d)
    rows, K = x2d.shape
    out = torch.empty((rows, 3 * K), dtype=torch.bfloat16, device=x2d.device)
    with _timed("split_bf16x3", 0.0):
        _lib.check(_lib.load().avi_split_bf16x3(_ptr(x2d), _ptr(out), C.c_int64(rows), C.c_int32(K), _stream()), "avi_split_bf16x3")
    return out


# How fp32 operands are contracted: "simt" = the CUDA-core fp32 kernel (avi_gemm_f32, the <= 1e-5 mode); "x3" = each operand split into
# bf16 hi / lo terms, A = [hi|lo|hi] against W = [hi|hi|lo] on the tcgen05 GEMM (avi_split_bf16_terms): 5e-6 of max|C| per GEMM,
# 3e-5 relative over the whole wav2vec2 stack at 5.4x the speed of "simt" (profiles/r2/bench_fp32_*.json) - an opt-in middle mode.
# A six-term split (every product down to 2^-24) was MEASURED NO BETTER (8.6e-6 per GEMM, 6e-5 over the stack): the floor is the
# tensor core's own fp32 accumulation, which truncates (error grows linearly with the number of accumulated k-blocks), not the
# dropped lo x lo products - so it is not offered.
FP32_GEMM = os.environ.get("AVI_B200_FP32_GEMM", "simt").lower()
_SPLIT_PATTERNS = {"x3": (3, (0, 1, 0), (0, 0, 1))}
if FP32_GEMM not in ("simt",) + tuple(_SPLIT_PATTERNS):
    raise ValueError("AVI_B200_FP32_GEMM must be simt or x3")


DYN_GEMM, DYN_CONV0 = 1, 2


def set_dynamic_tiles(mask: int) -> int:
    """Which persistent kernels take their tiles through cluster launch control (bit set) instead of a static walk; returns the
    previous mask. Results are bit-identical either way (include/avi_b200.h avi_set_dynamic_tiles)."""
    return int(_lib.load().avi_set_dynamic_tiles(C.c_int32(mask)))


def set_pdl(on: bool) -> bool:
    """Programmatic dependent launch of the GEMM / attention / LayerNorm kernels (include/avi_b200.h avi_set_pdl); returns the previous
    setting."""
    return bool(_lib.load().avi_set_pdl(C.c_int32(1 if on else 0)))


def split_bf16_terms(x2d, pattern):
    """[rows, K] fp32 -> [rows, len(pattern) * K] bf16 blocks of the bf16 terms named by `pattern` (0 = hi, 1 = mid, 2 = lo)."""
    _need_cuda(x2d)
    rows, K = x2d.shape
    out = torch.empty((rows, len(pattern) * K), dtype=torch.bfloat16, device=x2d.device)
    packed = 0
    for t, w in enumerate(pattern):
        packed |= int(w) << (2 * t)
    with _timed("split_bf16_terms", 0.0):
        _lib.check(_lib.load().avi_split_bf16_terms(_ptr(x2d), _ptr(out), C.c_int64(rows), C.c_int32(K), C.c_int32(len(pattern)),
                                                    C.c_uint32(packed), _stream()), "avi_split_bf16_terms")
    return out


def _gemm_f32_split(A, W, bias, out, mode, *, rows, N, K, batch, act, residual, out2, conv_taps, conv_stride, a_ld, a_batch_stride,
                    a_rows_alloc, c_ld, c_batch_stride, res_ld, res_batch_stride, algorithmic_flops, profile, conv_taps_x, conv_row_pitch):
    """fp32 operands on the bf16 tensor path as split terms; returns False when the layout is not one the split supports (the caller
    then takes the CUDA-core kernel)."""
    nt, pa, pw = _SPLIT_PATTERNS[mode]
    cin = K // conv_taps
    if cin % 64 != 0 or (a_ld is not None and a_ld != cin) or a_batch_stride % cin != 0 or not A.is_contiguous() or not W.is_contiguous():
        return False
    if A.numel() % cin != 0 or W.numel() != N * K or out.dtype != torch.float32 or out2 is not None:
        return False
    A3 = split_bf16_terms(A.view(-1, cin), pa)                       # every row of the activation buffer, channel blocks per term
    W3 = split_bf16_terms(W.view(-1, cin), pw).view(N, conv_taps * nt * cin)
    flops = 2.0 * batch * rows * N * K if algorithmic_flops is None else float(algorithmic_flops)
    gemm(A3, W3, bias, out, rows=rows, N=N, K=nt * K, batch=batch, act=act, residual=residual, conv_taps=conv_taps,
         conv_stride=conv_stride, a_ld=nt * cin, a_batch_stride=nt * a_batch_stride, a_rows_alloc=a_rows_alloc, c_ld=c_ld,
         c_batch_stride=c_batch_stride, res_ld=res_ld, res_batch_stride=res_batch_stride, algorithmic_flops=flops, profile=profile,
         conv_taps_x=conv_taps_x, conv_row_pitch=conv_row_pitch)
    return True


def gemm(A, W, bias, out, *, rows, N, K, batch=1, act=ACT_NONE, residual=None, out2=None, conv_taps=1, conv_stride=1,
         a_ld=None, a_batch_stride=0, a_rows_alloc=None, c_ld=None, c_batch_stride=0, res_ld=None, res_batch_stride=0,
         algorithmic_flops=None, profile=None, tf32=False, conv_taps_x=0, conv_row_pitch=0):
    """C[b,r,n] = act(sum_k A[b,r,k] W[n,k] + bias[n]) (+ residual). fp32 A/W -> CUDA-core kernel, bf16 -> tcgen05 kernel."""
    _need_cuda(A, W, bias, out, residual, out2)
    if A.dtype != W.dtype:
        raise TypeError("A and W must share a dtype")
    if A.dtype == torch.float32 and not tf32 and FP32_GEMM in _SPLIT_PATTERNS:
        if _gemm_f32_split(A, W, bias, out, FP32_GEMM, rows=rows, N=N, K=K, batch=batch, act=act, residual=residual, out2=out2,
                           conv_taps=conv_taps, conv_stride=conv_stride, a_ld=a_ld, a_batch_stride=a_batch_stride,
                           a_rows_alloc=a_rows_alloc, c_ld=c_ld, c_batch_stride=c_batch_stride, res_ld=res_ld,
                           res_batch_stride=res_batch_stride, algorithmic_flops=algorithmic_flops, profile=profile,
                           conv_taps_x=conv_taps_x, conv_row_pitch=conv_row_pitch):
            return out
    args = AviGemmArgs()
    args.A, args.W, args.bias, args.residual = A.data_ptr(), W.data_ptr(), (bias.data_ptr() if bias is not None else None), \
        (residual.data_ptr() if residual is not None else None)
    args.C = out.data_ptr()
    args.C2 = out2.data_ptr() if out2 is not None else None
    args.batch, args.rows, args.N, args.K = batch, rows, N, K
    args.conv_taps, args.conv_stride = conv_taps, conv_stride
    args.conv_taps_x, args.conv_row_pitch = conv_taps_x, conv_row_pitch
    args.a_ld = a_ld if a_ld is not None else K // conv_taps
    args.a_batch_stride = a_batch_stride
    args.a_rows_alloc = a_rows_alloc if a_rows_alloc is not None else (
        (rows - 1) * conv_stride + conv_taps if conv_taps_x == 0 else rows - 1 + (conv_taps // conv_taps_x - 1) * conv_row_pitch + conv_taps_x)
    args.c_ld = c_ld if c_ld is not None else N
    args.c_batch_stride = c_batch_stride
    args.res_ld = res_ld if res_ld is not None else N
    args.res_batch_stride = res_batch_stride
    args.a_dtype, args.c_dtype, args.act = _dt(A), _dt(out), act
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("bias must be fp32")
    if residual is not None and residual.dtype != torch.float32:
        raise TypeError("residual must be fp32")
    if out2 is not None and out2.dtype == out.dtype:
        raise TypeError("out2 must be the other dtype")
    lib = _lib.load()
    # bench.py's roofline counts ALGORITHMIC flops: callers that pad the contraction with structural zeros say so
    flops = 2.0 * batch * rows * N * K if algorithmic_flops is None else float(algorithmic_flops)
    if tf32:                     # fp32 operands on the tensor cores as TF32 (10-bit significand), same kernel as the bf16 path
        if A.dtype != torch.float32:
            raise TypeError("tf32 GEMM takes fp32 operands")
        with _timed("gemm_tf32_tc", flops):
            _lib.check(lib.avi_gemm_tf32_tc(C.byref(args), _stream()), "avi_gemm_tf32_tc")
        return out
    if profile is not None:      # (name, work): a launch whose roofline is not the tensor pipe (the HBM-write-bound vertex head)
        with _timed(profile[0], float(profile[1])):
            _lib.check((lib.avi_gemm_bf16_tc if A.dtype == torch.bfloat16 else lib.avi_gemm_f32)(C.byref(args), _stream()), "avi_gemm")
        return out
    if A.dtype == torch.bfloat16:
        with _timed("gemm_bf16_tc", flops):
            _lib.check(lib.avi_gemm_bf16_tc(C.byref(args), _stream()), "avi_gemm_bf16_tc")
    else:
        with _timed("gemm_f32", flops):
            _lib.check(lib.avi_gemm_f32(C.byref(args), _stream()), "avi_gemm_f32")
    return out


def linear(x2d: torch.Tensor, W: torch.Tensor, bias, *, act=ACT_NONE, residual=None, out_dtype=None, out2_dtype=None,
           out=None, tf32=False):
    """nn.Linear on [rows, K] activations: returns out (and out2 if requested)."""
    rows, K = x2d.shape
    N = W.shape[0]
    assert W.shape[1] == K and x2d.is_contiguous() and W.is_contiguous()
    if out is None:
        out = torch.empty((rows, N), dtype=out_dtype or torch.float32, device=x2d.device)
    out2 = torch.empty((rows, N), dtype=out2_dtype, device=x2d.device) if out2_dtype is not None else None
    gemm(x2d, W, bias, out, rows=rows, N=N, K=K, act=act, residual=residual, out2=out2, a_rows_alloc=rows, tf32=tf32)
    return (out, out2) if out2 is not None else out


def conv0_gn_gelu(audio, w, gn_w, gn_b, out, out_batch_stride, eps=1e-5):
    _need_cuda(audio, w, out)
    B, n = audio.shape
    Cc = w.shape[0]
    stats = torch.empty((B, Cc, 2), dtype=torch.float64, device=audio.device)
    with _timed("conv0_gn_gelu", float(out.numel() * out.element_size())):
        _lib.check(_lib.load().avi_w2v_conv0_gn_gelu(_ptr(audio), _ptr(w), _ptr(gn_w), _ptr(gn_b), _ptr(stats), _ptr(out),
                                                     C.c_int32(_dt(out)), C.c_int64(out_batch_stride), C.c_int32(B), C.c_int32(n),
                                                     C.c_int32(Cc), C.c_float(eps), _stream()), "avi_w2v_conv0_gn_gelu")
    return out


def conv0_pack_tc(w):
    _need_cuda(w)
    wp = torch.empty((512, 64), dtype=torch.bfloat16, device=w.device)
    _lib.check(_lib.load().avi_w2v_conv0_pack_tc(_ptr(w.contiguous().float()), _ptr(wp), _stream()), "avi_w2v_conv0_pack_tc")
    return wp


def conv0_gn_gelu_tc(audio, w, w_packed, gn_w, gn_b, out, out_batch_stride, eps=1e-5):
    """conv0 + GroupNorm + GELU on tensor cores (bf16 output, 512 channels)."""
    _need_cuda(audio, w, out)
    B, n = audio.shape
    Cc = w.shape[0]
    stats = torch.empty((B, Cc, 2), dtype=torch.float64, device=audio.device)
    with _timed("conv0_gn_gelu", float(out.numel() * out.element_size())):
        _lib.check(_lib.load().avi_w2v_conv0_gn_gelu_tc(_ptr(audio), _ptr(w), _ptr(w_packed), _ptr(gn_w), _ptr(gn_b), _ptr(stats),
                                                        _ptr(out), C.c_int64(out_batch_stride), C.c_int32(B), C.c_int32(n),
                                                        C.c_int32(Cc), C.c_float(eps), _stream()), "avi_w2v_conv0_gn_gelu_tc")
    return out


def lerp_layernorm(x, in_batch_stride, B, T_in, T_out, ln_w, ln_b, want_f32, want_bf16, eps=1e-5):
    _need_cuda(x)
    Cc = ln_w.numel()
    o32 = torch.empty((B * T_out, Cc), dtype=torch.float32, device=x.device) if want_f32 else None
    o16 = torch.empty((B * T_out, Cc), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    with _timed("lerp_layernorm", 0.0):
        _lib.check(_lib.load().avi_w2v_lerp_layernorm(_ptr(x), C.c_int32(_dt(x)), C.c_int64(in_batch_stride), _ptr(ln_w), _ptr(ln_b),
                                                      _ptr(o32), _ptr(o16), C.c_int32(B), C.c_int32(T_in), C.c_int32(T_out),
                                                      C.c_int32(Cc), C.c_float(eps), _stream()), "avi_w2v_lerp_layernorm")
    return o32, o16


def layernorm(x, w, b, *, res=None, want_f32=True, want_bf16=False, eps=1e-5):
    _need_cuda(x, res)
    assert x.dtype == torch.float32 and x.is_contiguous()
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    o32 = torch.empty_like(x) if want_f32 else None
    o16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    # algorithmic bytes: read x (+ residual), write the fp32 and / or bf16 result
    with _timed("layernorm", float(x.numel() * (4 + (4 if res is not None else 0) + (4 if want_f32 else 0) + (2 if want_bf16 else 0)))):
        _lib.check(_lib.load().avi_layernorm(_ptr(x), _ptr(res), _ptr(w), _ptr(b), _ptr(o32), _ptr(o16), C.c_int64(rows),
                                             C.c_int32(Cc), C.c_float(eps), _stream()), "avi_layernorm")
    return o32, o16


def posconv_ln(x, w_packed, conv_bias, ln_w, ln_b, B, T, groups, k, want_bf16, eps=1e-5):
    _need_cuda(x)
    Cc = x.shape[-1]
    o32 = torch.empty((B * T, Cc), dtype=torch.float32, device=x.device)
    o16 = torch.empty((B * T, Cc), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    with _timed("posconv_ln", 2.0 * B * T * Cc * (Cc // groups) * k):
        _lib.check(_lib.load().avi_w2v_posconv_ln(_ptr(x), _ptr(w_packed), _ptr(conv_bias), _ptr(ln_w), _ptr(ln_b), _ptr(o32),
                                                  _ptr(o16), C.c_int32(B), C.c_int32(T), C.c_int32(Cc), C.c_int32(groups),
                                                  C.c_int32(k), C.c_float(eps), _stream()), "avi_w2v_posconv_ln")
    return o32, o16


def ff_denorm_coeff(coeff, mean, std, pose, n_exp=50):
    """exp [F, n_exp] = coeff[:, :n_exp] * std + mean; pose[:, :3] zeroed in place (faceformer_disentangle.py:426-430)."""
    _need_cuda(coeff, mean, std, pose)
    assert coeff.is_contiguous() and pose.is_contiguous() and coeff.dtype == pose.dtype == torch.float32
    F_, nc = coeff.shape
    out = torch.empty((F_, n_exp), dtype=torch.float32, device=coeff.device)
    with _timed("ff_denorm_coeff", 0.0):
        _lib.check(_lib.load().avi_ff_denorm_coeff(_ptr(coeff), _ptr(mean), _ptr(std), _ptr(pose), _ptr(out), C.c_int32(F_), C.c_int32(nc),
                                                   C.c_int32(n_exp), C.c_int32(pose.shape[-1]), _stream()), "avi_ff_denorm_coeff")
    return out


def ff_fill_cond(eye, emo, out, B, T):
    """out[:, 0:6] = eye ([6] or [B*T, 6]), out[:, 6:36] = emo [B, >=T, 30]; out [B*T, ld] fp32."""
    _need_cuda(eye, emo, out)
    per_row = 0 if eye.numel() == 6 else 1
    with _timed("ff_fill_cond", 0.0):
        _lib.check(_lib.load().avi_ff_fill_cond(_ptr(eye), C.c_int32(per_row), _ptr(emo), C.c_int64(emo.stride(0)), _ptr(out), C.c_int32(B),
                                                C.c_int32(T), C.c_int32(out.stride(0)), _stream()), "avi_ff_fill_cond")
    return out


def posconv_tc(xpad, w_band, bias, B, T, groups, k):
    """Grouped positional conv (+ bias) on tcgen05, activation slab resident in shared memory: xpad bf16 [B, Tp, C] -> fp32 [B*T, C]."""
    _need_cuda(xpad, w_band, bias)
    Tp, Cc = xpad.shape[1], xpad.shape[2]
    pc = torch.empty((B * T, Cc), dtype=torch.float32, device=xpad.device)
    with _timed("posconv_tc", 2.0 * B * T * Cc * (Cc // groups) * k):
        _lib.check(_lib.load().avi_w2v_posconv_tc(_ptr(xpad), _ptr(w_band), _ptr(bias), _ptr(pc), C.c_int32(B), C.c_int32(T), C.c_int32(Tp),
                                                  C.c_int32(Cc), C.c_int32(groups), C.c_int32(k), _stream()), "avi_w2v_posconv_tc")
    return pc


def posconv_merge_ln(x, pc, ln_w, ln_b, want_bf16, eps=1e-5):
    """out = LayerNorm(x + GELU(pc)), written over pc (fp32) plus an optional bf16 copy."""
    _need_cuda(x, pc)
    rows, Cc = x.shape
    o16 = torch.empty((rows, Cc), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    with _timed("posconv_merge_ln", float(x.numel() * 12)):
        _lib.check(_lib.load().avi_w2v_posconv_merge_ln(_ptr(x), _ptr(pc), _ptr(ln_w), _ptr(ln_b), _ptr(pc), _ptr(o16),
                                                        C.c_int64(rows), C.c_int32(Cc), C.c_float(eps), _stream()),
                   "avi_w2v_posconv_merge_ln")
    return pc, o16


def mha(qkv, B, T, H, D, scale):
    _need_cuda(qkv)
    out = torch.empty((B * T, H * D), dtype=qkv.dtype, device=qkv.device)
    lib = _lib.load()
    if lib.avi_mha_fwd_tc_supported(C.c_int32(_dt(qkv)), C.c_int32(T), C.c_int32(D)):
        with _timed("mha_tc", 4.0 * B * H * T * T * D):
            _lib.check(lib.avi_mha_fwd_tc(_ptr(qkv), _ptr(out), C.c_int32(B), C.c_int32(T), C.c_int32(H), C.c_int32(D),
                                          C.c_float(scale), _stream()), "avi_mha_fwd_tc")
        return out
    with _timed("mha", 4.0 * B * H * T * T * D):
        _lib.check(_lib.load().avi_mha_fwd(_ptr(qkv), _ptr(out), C.c_int32(_dt(qkv)), C.c_int32(B), C.c_int32(T), C.c_int32(H),
                                           C.c_int32(D), C.c_float(scale), _stream()), "avi_mha_fwd")
    return out


def ff_decoder_ar(wstruct: AviDecoderWeights, cross, style, B, T, fd, period):
    _need_cuda(cross, style)
    hidden = torch.empty((B, T, fd), dtype=torch.float32, device=cross.device)
    kv = torch.empty((B, 2, T, fd + 1), dtype=torch.float32, device=cross.device)
    with _timed("ff_decoder_ar", 0.0):
        _lib.check(_lib.load().avi_ff_decoder_ar(C.byref(wstruct), _ptr(cross), _ptr(style), _ptr(hidden), _ptr(kv), C.c_int32(B),
                                                 C.c_int32(T), C.c_int32(fd), C.c_int32(period), _stream()), "avi_ff_decoder_ar")
    return hidden


def ff_biased_attn(qkv, B, T, fd, period):
    _need_cuda(qkv)
    out = torch.empty((B, T, fd), dtype=torch.float32, device=qkv.device)
    _lib.check(_lib.load().avi_ff_biased_attn(_ptr(qkv), _ptr(out), C.c_int32(B), C.c_int32(T), C.c_int32(fd), C.c_int32(period),
                                              _stream()), "avi_ff_biased_attn")
    return out


# Vertex rows are 15069 floats = 60276 bytes: only 4-byte aligned, which costs the HBM-write-bound kernels ~40 % (measured:
# profiles/r1/vhead_align_probe.txt). Outputs are therefore allocated with the row stride rounded up to a multiple of 4 floats and
# returned as a [..., 15069] view of that buffer (`.contiguous()` gives the dense layout).
PAD_VERTEX_ROWS = True


def padded_cols(n: int) -> int:
    return ((n + 3) // 4) * 4 if PAD_VERTEX_ROWS else n


def empty_rows(rows: int, cols: int, device) -> torch.Tensor:
    """fp32 [rows, cols] whose row stride is padded to 16 bytes (a view of a [rows, padded] buffer)."""
    return torch.empty((rows, padded_cols(cols)), dtype=torch.float32, device=device)[:, :cols]


def flame_pack(shapedirs, posedirs, v_template, J_regressor, K_pad):
    _need_cuda(shapedirs)
    V, _, NB = shapedirs.shape
    dirs = torch.empty((K_pad, V * 3), dtype=torch.float32, device=shapedirs.device)
    jreg = torch.empty((15, NB + 1), dtype=torch.float32, device=shapedirs.device)
    _lib.check(_lib.load().avi_flame_pack(_ptr(shapedirs), _ptr(posedirs), _ptr(v_template), _ptr(J_regressor), _ptr(dirs),
                                          _ptr(jreg), C.c_int32(V), C.c_int32(NB), C.c_int32(K_pad), _stream()), "avi_flame_pack")
    return dirs, jreg


def flame_lbs(betas, full_pose, dirs, jreg, lbs_weights, V, NB, K_pad, want_joints=False, want_dyn_rows=False, rotmat=False):
    """rotmat: full_pose is [F, 45] rotation matrices (lbs(pose2rot=False)) instead of [F, 15] axis-angle."""
    _need_cuda(betas, full_pose)
    assert full_pose.shape[1] == (45 if rotmat else 15)
    F = betas.shape[0]
    dev = betas.device
    coef = torch.empty((F, K_pad), dtype=torch.float32, device=dev)
    A = torch.empty((F, 5, 12), dtype=torch.float32, device=dev)
    verts = torch.empty((F, V, 3), dtype=torch.float32, device=dev)
    joints = torch.empty((F, 5, 3), dtype=torch.float32, device=dev) if want_joints else None
    rows = torch.empty((F,), dtype=torch.int32, device=dev) if want_dyn_rows else None
    with _timed("flame_lbs", float(F) * (V * 12 + 4 * (NB + 6))):  # algorithmic bytes: verts written + coefficients read
        _lib.check(_lib.load().avi_flame_lbs_fwd_ex(_ptr(betas), _ptr(full_pose), C.c_int32(1 if rotmat else 0), _ptr(dirs), _ptr(jreg),
                                                    _ptr(lbs_weights), _ptr(coef), _ptr(A), _ptr(verts), _ptr(joints), _ptr(rows),
                                                    C.c_int32(F), C.c_int32(V), C.c_int32(NB), C.c_int32(K_pad), _stream()),
                   "avi_flame_lbs_fwd")
    return verts, joints, rows


def flame_tc_supported(NB: int) -> bool:
    return bool(_lib.load().avi_flame_tc_supported(C.c_int32(NB)))


def flame_pack_tc(dirs32, V, NB):
    V_pad = ((V + 127) // 128) * 128
    dirs16 = torch.empty((3, V_pad, 192), dtype=torch.float16, device=dirs32.device)
    _lib.check(_lib.load().avi_flame_pack_tc(_ptr(dirs32), _ptr(dirs16), C.c_int32(V), C.c_int32(NB), C.c_int32(V_pad), _stream()),
               "avi_flame_pack_tc")
    return dirs16


def flame_lbs_tc(betas, full_pose, dirs16, jreg, lbs_weights, v_template, V, NB, K_pad, want_joints=False, want_dyn_rows=False,
                 padded=False, rotmat=False):
    """Same results contract as flame_lbs, blend + skinning on the tcgen05 path."""
    _need_cuda(betas, full_pose)
    assert full_pose.shape[1] == (45 if rotmat else 15)
    F = betas.shape[0]
    dev = betas.device
    coef = torch.empty((F, K_pad), dtype=torch.float32, device=dev)
    coef16 = torch.empty((F, 192), dtype=torch.float16, device=dev)
    A = torch.empty((F, 5, 12), dtype=torch.float32, device=dev)
    vrows = empty_rows(F, V * 3, dev) if padded else torch.empty((F, V * 3), dtype=torch.float32, device=dev)
    verts = vrows.view(F, V, 3)
    joints = torch.empty((F, 5, 3), dtype=torch.float32, device=dev) if want_joints else None
    rows = torch.empty((F,), dtype=torch.int32, device=dev) if want_dyn_rows else None
    lib = _lib.load()
    with _timed("flame_lbs", float(F) * (V * 12 + 4 * (NB + 6))):
        _lib.check(lib.avi_flame_prologue_ex(_ptr(betas), _ptr(full_pose), C.c_int32(1 if rotmat else 0), _ptr(jreg), _ptr(coef), _ptr(A),
                                             _ptr(joints), _ptr(rows), C.c_int32(F), C.c_int32(NB), C.c_int32(K_pad), _stream()),
                   "avi_flame_prologue")
        _lib.check(lib.avi_flame_blend_skin_tc_grouped(_ptr(coef), _ptr(A), _ptr(dirs16), _ptr(lbs_weights), _ptr(v_template),
                                                       C.c_int64(0), _ptr(coef16), _ptr(verts), C.c_int64(vrows.stride(0)), C.c_int32(F),
                                                       C.c_int32(V), C.c_int32(NB + 36), C.c_int32(0), C.c_int32(K_pad),
                                                       C.c_int32(dirs16.shape[1]), C.c_int32(F), _stream()),
                   "avi_flame_blend_skin_tc_grouped")
    return verts, joints, rows


def flame_landmarks(verts, faces, idx, bary, per_frame: bool):
    _need_cuda(verts, faces, idx, bary)
    F, V, _ = verts.shape
    L = idx.shape[-1]
    out = torch.empty((F, L, 3), dtype=torch.float32, device=verts.device)
    _lib.check(_lib.load().avi_flame_landmarks(_ptr(verts), _ptr(faces), _ptr(idx), _ptr(bary), _ptr(out), C.c_int32(F),
                                               C.c_int32(V), C.c_int32(L), C.c_int32(1 if per_frame else 0), _stream()),
               "avi_flame_landmarks")
    return out


# ------------------------------------------------------------------------------------------------ diffusion prior
def prior_layer_floats() -> int:
    return int(_lib.load().avi_prior_layer_floats())


def prior_time_embed(times, w0t, b0, w1t, b1, w2t, b2):
    _need_cuda(times, w0t)
    steps = times.numel()
    temb = torch.empty((steps, 128), dtype=torch.float32, device=times.device)
    with _timed("prior_time_embed", 0.0):
        _lib.check(_lib.load().avi_prior_time_embed(_ptr(times), _ptr(w0t), _ptr(b0), _ptr(w1t), _ptr(b1), _ptr(w2t), _ptr(b2),
                                                    _ptr(temb), C.c_int32(steps), _stream()), "avi_prior_time_embed")
    return temb


def prior_sample(net: AviPriorNet, temb, sched, text_embed, x_init, noise, out_scale, samples_per_cta=0, null_pred=None, cond_scale=1.0):
    """One launch = the whole DDPM/DDIM loop. text_embed/x_init [B,128], noise [steps,B,128], sched [steps,6] -> [B,128].
    null_pred [steps,128] + cond_scale: classifier-free guidance (the null pass of every step, precomputed)."""
    _need_cuda(temb, sched, text_embed, x_init, noise)
    B, steps = text_embed.shape[0], sched.shape[0]
    for t, shp in ((temb, (steps, 128)), (text_embed, (B, 128)), (x_init, (B, 128)), (noise, (steps, B, 128)), (sched, (steps, 6))):
        if tuple(t.shape) != shp or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"prior_sample: expected contiguous fp32 {shp}, got {tuple(t.shape)} {t.dtype}")
    out = torch.empty((B, 128), dtype=torch.float32, device=text_embed.device)
    # algorithmic work: 12.8 MFLOP per sample-step
    with _timed("prior_sample", 12.8e6 * B * steps):
        if null_pred is not None and (tuple(null_pred.shape) != (steps, 128) or null_pred.dtype != torch.float32 or not null_pred.is_contiguous()):
            raise ValueError("prior_sample: null_pred must be contiguous fp32 [steps, 128]")
        _lib.check(_lib.load().avi_prior_sample_cfg(C.byref(net), _ptr(temb), _ptr(sched), _ptr(text_embed), _ptr(x_init), _ptr(noise),
                                                    _ptr(null_pred), C.c_float(cond_scale), _ptr(out), C.c_int32(B), C.c_int32(steps),
                                                    C.c_float(out_scale), C.c_int32(samples_per_cta), _stream()), "avi_prior_sample_cfg")
    return out


def ln_gelu_res(x, w, b, res=None, want_bf16=False, eps=1e-5):
    """GELU(LayerNorm(x)) (+ res) over the last dim (<= 4096)."""
    _need_cuda(x, res)
    assert x.dtype == torch.float32 and x.is_contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    o32 = torch.empty_like(x)
    o16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    with _timed("ln_gelu_res", float(x.numel() * 8)):
        _lib.check(_lib.load().avi_ln_gelu_res(_ptr(x), _ptr(w), _ptr(b), _ptr(res), _ptr(o32), _ptr(o16), C.c_int64(rows),
                                               C.c_int32(Cc), C.c_float(eps), _stream()), "avi_ln_gelu_res")
    return o32, o16


# ------------------------------------------------------------------------------------------------ EMOTE decoder pieces
def audio_znorm(x, eps=1e-7):
    _need_cuda(x)
    x = x.contiguous().float()
    y = torch.empty_like(x)
    with _timed("audio_znorm", float(x.numel() * 8)):
        _lib.check(_lib.load().avi_audio_znorm(_ptr(x), _ptr(y), C.c_int32(x.shape[0]), C.c_int64(x.shape[1]), C.c_float(eps), _stream()),
                   "avi_audio_znorm")
    return y


def mha_small(qkv, B, T, H, D, slopes=None, want_bf16=False):
    """fp32 qkv [B*T, 3*H*D] -> (out fp32 [B*T, H*D], optional bf16 copy); additive -slope_h*|i-j| bias when `slopes` is given."""
    _need_cuda(qkv, slopes)
    assert qkv.dtype == torch.float32 and qkv.is_contiguous()
    o32 = torch.empty((B * T, H * D), dtype=torch.float32, device=qkv.device)
    o16 = torch.empty((B * T, H * D), dtype=torch.bfloat16, device=qkv.device) if want_bf16 else None
    with _timed("mha_small", 4.0 * B * H * T * T * D):
        _lib.check(_lib.load().avi_mha_small_fwd(_ptr(qkv), _ptr(o32), _ptr(o16), C.c_int32(B), C.c_int32(T), C.c_int32(H), C.c_int32(D),
                                                 C.c_float(D ** -0.5), _ptr(slopes), _stream()), "avi_mha_small_fwd")
    return o32, o16


def stage_rows(src, B, L, Lp, front, mode, dtype=torch.float32):
    """mode 0 zero pad, 1 replicate pad, 2 zero insertion (see include/avi_b200.h)."""
    _need_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    Cc = src.shape[-1]
    dst = torch.empty((B, Lp, Cc), dtype=dtype, device=src.device)
    with _timed("stage_rows", 0.0):
        _lib.check(_lib.load().avi_stage_rows(_ptr(src), _ptr(dst), C.c_int32(_dt(dst)), C.c_int32(B), C.c_int32(L), C.c_int32(Lp),
                                              C.c_int32(Cc), C.c_int32(front), C.c_int32(mode), _stream()), "avi_stage_rows")
    return dst


def lrelu_bn_repeat(x, bn_scale, bn_shift, B, L, repeat, slope=0.2):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    Cc = x.shape[-1]
    y = torch.empty((B, L * repeat, Cc), dtype=torch.float32, device=x.device)
    with _timed("lrelu_bn_repeat", 0.0):
        _lib.check(_lib.load().avi_lrelu_bn_repeat(_ptr(x), _ptr(bn_scale), _ptr(bn_shift), _ptr(y), C.c_int32(B), C.c_int32(L),
                                                   C.c_int32(Cc), C.c_int32(repeat), C.c_float(slope), _stream()), "avi_lrelu_bn_repeat")
    return y


def sub_add_rows_(a, neutral, tpl):
    """In place: a[b, t, :] = (a[b, t, :] - neutral[b, :]) + tpl[b, :]; `a` may have a padded row stride."""
    _need_cuda(a, neutral, tpl)
    B, T, Cc = a.shape
    assert a.dtype == torch.float32 and a.stride(2) == 1 and a.stride(0) == T * a.stride(1) and neutral.is_contiguous() and tpl.is_contiguous()
    with _timed("sub_add_rows", float(a.numel() * 8)):
        _lib.check(_lib.load().avi_sub_add_rows(_ptr(a), _ptr(neutral), _ptr(tpl), _ptr(a), C.c_int32(B), C.c_int32(T), C.c_int32(Cc),
                                                C.c_int64(a.stride(1)), _stream()), "avi_sub_add_rows")
    return a


def flame_pack_tc_rows(dirs32, V, row0, n_dirs):
    V_pad = ((V + 127) // 128) * 128
    dirs16 = torch.empty((3, V_pad, 192), dtype=torch.float16, device=dirs32.device)
    _lib.check(_lib.load().avi_flame_pack_tc_rows(_ptr(dirs32), _ptr(dirs16), C.c_int32(V), C.c_int32(row0), C.c_int32(n_dirs),
                                                  C.c_int32(V_pad), _stream()), "avi_flame_pack_tc_rows")
    return dirs16


def flame_lbs_tc_grouped(betas, full_pose, dirs16_exp, jreg, lbs_weights, templates, V, NB, n_shape, K_pad, frames_per_group):
    """FLAME with one shape per group of frames: the shape blendshapes are already folded into `templates` [G, V*3]; the
    tensor-core contraction covers expression + pose-corrective columns only. betas [F, NB] still carries the shape (joints)."""
    _need_cuda(betas, full_pose, templates)
    F = betas.shape[0]
    dev = betas.device
    coef = torch.empty((F, K_pad), dtype=torch.float32, device=dev)
    coef16 = torch.empty((F, 192), dtype=torch.float16, device=dev)
    A = torch.empty((F, 5, 12), dtype=torch.float32, device=dev)
    vrows = empty_rows(F, V * 3, dev)
    verts = vrows.view(F, V, 3)
    lib = _lib.load()
    n_dirs = NB - n_shape + 36
    with _timed("flame_lbs", float(F) * (V * 12 + 4 * (NB - n_shape + 6))):
        _lib.check(lib.avi_flame_prologue(_ptr(betas), _ptr(full_pose), _ptr(jreg), _ptr(coef), _ptr(A), None, None,
                                          C.c_int32(F), C.c_int32(NB), C.c_int32(K_pad), _stream()), "avi_flame_prologue")
        _lib.check(lib.avi_flame_blend_skin_tc_grouped(_ptr(coef), _ptr(A), _ptr(dirs16_exp), _ptr(lbs_weights), _ptr(templates),
                                                       C.c_int64(V * 3), _ptr(coef16), _ptr(verts), C.c_int64(vrows.stride(0)),
                                                       C.c_int32(F), C.c_int32(V), C.c_int32(n_dirs), C.c_int32(n_shape), C.c_int32(K_pad),
                                                       C.c_int32(dirs16_exp.shape[1]), C.c_int32(frames_per_group), _stream()),
                   "avi_flame_blend_skin_tc_grouped")
    return verts


def flame_set_max_ctas(n: int):
    _lib.check(_lib.load().avi_flame_set_max_ctas(C.c_int32(int(n))), "avi_flame_set_max_ctas")


# ------------------------------------------------------------------------------------------------ training-step pieces
def _chk(rc, name):
    _lib.check(rc, name)


def transpose_cast(x2d, dtype, R_pad=None):
    """fp32 [R, C] (row stride allowed) -> dtype [C, R_pad] with zero padding (R_pad defaults to R rounded up to 64)."""
    _need_cuda(x2d)
    R, Cc = x2d.shape
    assert x2d.dtype == torch.float32 and x2d.stride(1) == 1
    R_pad = R_pad or ((R + 63) // 64) * 64
    out = torch.empty((Cc, R_pad), dtype=dtype, device=x2d.device)
    _chk(_lib.load().avi_transpose_cast(_ptr(x2d), _ptr(out), C.c_int32(_dt(out)), C.c_int32(R), C.c_int32(Cc), C.c_int64(x2d.stride(0)),
                                        C.c_int32(R_pad), _stream()), "avi_transpose_cast")
    return out


def cast_pad2d(x2d, dtype, R_pad=None, C_pad=None):
    _need_cuda(x2d)
    R, Cc = x2d.shape
    assert x2d.dtype == torch.float32 and x2d.stride(1) == 1
    R_pad, C_pad = R_pad or R, C_pad or Cc
    out = torch.empty((R_pad, C_pad), dtype=dtype, device=x2d.device)
    _chk(_lib.load().avi_cast_pad2d(_ptr(x2d), _ptr(out), C.c_int32(_dt(out)), C.c_int32(R), C.c_int32(Cc), C.c_int64(x2d.stride(0)),
                                    C.c_int32(R_pad), C.c_int32(C_pad), _stream()), "avi_cast_pad2d")
    return out


def tf_input_rows(gt, template, C_pad):
    """gt fp32 [B, T, C] (uniform row stride) -> fp32 [B*T, C_pad]: cat([template, gt[:, :-1]]) - template, zero padded."""
    _need_cuda(gt, template)
    B, T, Cc = gt.shape
    assert gt.stride(2) == 1 and gt.stride(0) == T * gt.stride(1)
    out = torch.empty((B * T, C_pad), dtype=torch.float32, device=gt.device)
    _chk(_lib.load().avi_tf_input_rows(_ptr(gt), C.c_int64(gt.stride(1)), _ptr(template), _ptr(out), C.c_int32(B), C.c_int32(T), C.c_int32(Cc),
                                       C.c_int32(C_pad), _stream()), "avi_tf_input_rows")
    return out


def ff_add_style_pe(x, style, pe, B, T, fd, period):
    _need_cuda(x, style, pe)
    stride = 0 if style.shape[0] == 1 else style.stride(0)
    _chk(_lib.load().avi_ff_add_style_pe(_ptr(x), _ptr(style), C.c_int64(stride), _ptr(pe), C.c_int32(B), C.c_int32(T), C.c_int32(fd),
                                         C.c_int32(period), _stream()), "avi_ff_add_style_pe")
    return x


def colsum(x2d, out=None, accumulate=False):
    _need_cuda(x2d)
    R, N = x2d.shape
    if out is None:
        out = torch.empty((N,), dtype=torch.float32, device=x2d.device)
    _chk(_lib.load().avi_colsum(_ptr(x2d), _ptr(out), C.c_int32(R), C.c_int32(N), C.c_int64(x2d.stride(0)), C.c_int32(1 if accumulate else 0),
                                _stream()), "avi_colsum")
    return out


def act_fwd(pre, act, want_f32=True, want_bf16=False):
    _need_cuda(pre)
    o32 = torch.empty_like(pre) if want_f32 else None
    o16 = torch.empty(pre.shape, dtype=torch.bfloat16, device=pre.device) if want_bf16 else None
    _chk(_lib.load().avi_act_fwd(_ptr(pre), _ptr(o32), _ptr(o16), C.c_int64(pre.numel()), C.c_int32(act), _stream()), "avi_act_fwd")
    return o32, o16


def act_bwd(pre, dout, act):
    _need_cuda(pre, dout)
    dpre = torch.empty_like(pre)
    _chk(_lib.load().avi_act_bwd(_ptr(pre), _ptr(dout.contiguous()), _ptr(dpre), C.c_int64(pre.numel()), C.c_int32(act), _stream()), "avi_act_bwd")
    return dpre


def layernorm_bwd(x, w, dy, dw, db, want_dx=True, eps=1e-5):
    """dw / db are accumulated into (pre-zeroed fp32 buffers)."""
    _need_cuda(x, dy)
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty_like(x) if want_dx else None
    _chk(_lib.load().avi_layernorm_bwd(_ptr(x.contiguous()), _ptr(w), _ptr(dy.contiguous()), _ptr(dx), _ptr(dw), _ptr(db), C.c_int64(rows),
                                       C.c_int32(Cc), C.c_float(eps), _stream()), "avi_layernorm_bwd")
    return dx


def _pmask(pmask, B, H, T):
    if pmask is None:
        return None
    _need_cuda(pmask)
    if tuple(pmask.shape) != (B, H, T, T) or pmask.dtype != torch.float32 or not pmask.is_contiguous():
        raise ValueError(f"attention dropout mask must be a contiguous fp32 [B={B}, H={H}, T={T}, T] tensor, got {tuple(pmask.shape)}")
    return pmask


def attn_train_fwd(qkv, B, T, H, D, bias_mode=0, period=1, want_p=True, pmask=None):
    """pmask: the (pre-scaled) dropout draw on the probabilities, train mode only; P keeps the softmax itself."""
    _need_cuda(qkv)
    out = torch.empty((B * T, H * D), dtype=torch.float32, device=qkv.device)
    P = torch.empty((B, H, T, T), dtype=torch.float32, device=qkv.device) if want_p else None
    _chk(_lib.load().avi_attn_train_fwd_drop(_ptr(qkv), _ptr(out), _ptr(P), _ptr(_pmask(pmask, B, H, T)), C.c_int32(B), C.c_int32(T),
                                             C.c_int32(H), C.c_int32(D), C.c_float(D ** -0.5), C.c_int32(bias_mode), C.c_int32(period),
                                             _stream()), "avi_attn_train_fwd")
    return out, P


def attn_train_bwd(qkv, P, dout, B, T, H, D, pmask=None):
    _need_cuda(qkv, P, dout)
    dqkv = torch.empty_like(qkv)
    dS = torch.empty_like(P)
    _chk(_lib.load().avi_attn_train_bwd_drop(_ptr(qkv), _ptr(P), _ptr(_pmask(pmask, B, H, T)), _ptr(dout.contiguous()), _ptr(dqkv), _ptr(dS),
                                             C.c_int32(B), C.c_int32(T), C.c_int32(H), C.c_int32(D), C.c_float(D ** -0.5), _stream()),
         "avi_attn_train_bwd")
    return dqkv


def mask_mul(x, mask, residual=None):
    """nn.Dropout with the draw as an input: x * mask (+ residual); also its backward (dy * mask). fp32, same number of elements."""
    _need_cuda(x, mask)
    if mask.dtype != torch.float32 or mask.numel() != x.numel() or not mask.is_contiguous():
        raise ValueError(f"dropout mask: contiguous fp32 with {x.numel()} elements expected, got {tuple(mask.shape)} {mask.dtype}")
    x = x.contiguous()
    y = torch.empty_like(x)
    _chk(_lib.load().avi_mask_mul_add(_ptr(x), _ptr(mask), _ptr(None if residual is None else residual.contiguous()), _ptr(y),
                                      C.c_int64(x.numel()), _stream()), "avi_mask_mul_add")
    return y


def spec_augment_fwd_(x, row_mask, embed):
    """In place: rows of x [rows, C] whose row_mask byte (uint8 [rows]) is set become masked_spec_embed (wav2vec.py:120-131)."""
    _need_cuda(x, row_mask, embed)
    assert x.is_contiguous() and row_mask.dtype == torch.uint8 and row_mask.numel() == x.shape[0]
    _chk(_lib.load().avi_spec_augment_fwd(_ptr(x), _ptr(row_mask), _ptr(embed.contiguous()), C.c_int64(x.shape[0]), C.c_int32(x.shape[1]),
                                          _stream()), "avi_spec_augment_fwd")
    return x


def spec_augment_bwd_(dx, row_mask, g_embed):
    """In place: g_embed <- sum of the masked rows of dx; those rows of dx <- 0."""
    _need_cuda(dx, row_mask, g_embed)
    assert dx.is_contiguous() and g_embed.is_contiguous() and row_mask.dtype == torch.uint8 and row_mask.numel() == dx.shape[0]
    _chk(_lib.load().avi_spec_augment_bwd(_ptr(dx), _ptr(row_mask), _ptr(g_embed), C.c_int64(dx.shape[0]), C.c_int32(dx.shape[1]),
                                          _stream()), "avi_spec_augment_bwd")
    return dx


def dropout_masks_(out, p, state, stream_id=0):
    """Fill the flat fp32 buffer `out` with pre-scaled dropout masks (Philox4x32-10 keyed by the device-resident `state`)."""
    _need_cuda(out, state)
    assert out.dtype == torch.float32 and out.is_contiguous() and state.dtype == torch.int32 and state.numel() == 4
    _chk(_lib.load().avi_dropout_masks(_ptr(out), C.c_int64(out.numel()), C.c_float(p), _ptr(state), C.c_uint32(stream_id), _stream()),
         "avi_dropout_masks")
    return out


def layerdrop_spec_draw_(blend, keep_flags, rows, layerdrop, spec, B, T, span_len, span_rate, min_spans, state):
    _need_cuda(keep_flags, state)
    n_layers = keep_flags.numel()
    assert blend is None or (blend.is_contiguous() and blend.numel() == 2 * n_layers * rows)
    assert spec is None or (spec.dtype == torch.uint8 and spec.numel() == B * T)
    _chk(_lib.load().avi_layerdrop_spec_draw(_ptr(blend), _ptr(keep_flags), C.c_int32(n_layers), C.c_int64(rows), C.c_float(layerdrop),
                                             _ptr(spec), C.c_int32(B), C.c_int32(T), C.c_int32(span_len), C.c_float(span_rate),
                                             C.c_int32(min_spans), _ptr(state), _stream()), "avi_layerdrop_spec_draw")


def draw_bump_step_(state):
    _need_cuda(state)
    _chk(_lib.load().avi_draw_bump_step(_ptr(state), _stream()), "avi_draw_bump_step")


def posconv_dw(x, dpc, B, T, groups, k):
    _need_cuda(x, dpc)
    Cc = x.shape[-1]
    dw = torch.empty((Cc, Cc // groups, k), dtype=torch.float32, device=x.device)
    _chk(_lib.load().avi_posconv_dw(_ptr(x.contiguous()), _ptr(dpc.contiguous()), _ptr(dw), C.c_int32(B), C.c_int32(T), C.c_int32(Cc),
                                    C.c_int32(groups), C.c_int32(k), _stream()), "avi_posconv_dw")
    return dw


def posconv_unfold_t(x, B, T, Tq, c0, CW, k, dtype):
    _need_cuda(x)
    Cc = x.shape[-1]
    out = torch.empty((k * CW, B * Tq), dtype=dtype, device=x.device)
    _chk(_lib.load().avi_posconv_unfold_t(_ptr(x.contiguous()), _ptr(out), C.c_int32(_dt(out)), C.c_int32(B), C.c_int32(T), C.c_int32(Tq),
                                          C.c_int32(Cc), C.c_int32(c0), C.c_int32(CW), C.c_int32(k), _stream()), "avi_posconv_unfold_t")
    return out


def weightnorm_bwd(v, g, dw):
    _need_cuda(v, g, dw)
    k = v.shape[-1]
    dv = torch.empty_like(v)
    dg = torch.empty((k,), dtype=torch.float32, device=v.device)
    _chk(_lib.load().avi_weightnorm_bwd(_ptr(v.contiguous()), _ptr(g.contiguous()), _ptr(dw), _ptr(dv), _ptr(dg), C.c_int32(v.numel() // k),
                                        C.c_int32(k), _stream()), "avi_weightnorm_bwd")
    return dv, dg.view(g.shape)


def mse_loss_grad(out2d, gt2d, loss_scale):
    """out2d / gt2d fp32 [rows, C] (row strides allowed) -> (loss fp64 device scalar, dout dense [rows, C])."""
    _need_cuda(out2d, gt2d)
    rows, Cc = out2d.shape
    dout = torch.empty((rows, Cc), dtype=torch.float32, device=out2d.device)
    loss = torch.empty((1,), dtype=torch.float64, device=out2d.device)
    _chk(_lib.load().avi_mse_loss_grad(_ptr(out2d), _ptr(gt2d), _ptr(dout), _ptr(loss), C.c_int64(rows), C.c_int32(Cc), C.c_int64(out2d.stride(0)),
                                       C.c_int64(gt2d.stride(0)), C.c_float(loss_scale), _stream()), "avi_mse_loss_grad")
    return loss, dout


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0, p_bf16=None):
    _need_cuda(p, g, m, v, p_bf16)
    _chk(_lib.load().avi_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p_bf16), C.c_int64(p.numel()), C.c_float(lr), C.c_float(beta1),
                                   C.c_float(beta2), C.c_float(eps), C.c_int32(step), C.c_float(grad_scale), _stream()), "avi_adam_step")


def add_f32(a, b):
    _need_cuda(a, b)
    y = torch.empty_like(a)
    _chk(_lib.load().avi_add_f32(_ptr(a.contiguous()), _ptr(b.contiguous()), _ptr(y), C.c_int64(a.numel()), _stream()), "avi_add_f32")
    return y


def w2v_lerp(x, in_batch_stride, B, T_in, T_out, Cc):
    _need_cuda(x)
    out = torch.empty((B * T_out, Cc), dtype=torch.float32, device=x.device)
    _chk(_lib.load().avi_w2v_lerp(_ptr(x), C.c_int32(_dt(x)), C.c_int64(in_batch_stride), _ptr(out), C.c_int32(B), C.c_int32(T_in),
                                  C.c_int32(T_out), C.c_int32(Cc), _stream()), "avi_w2v_lerp")
    return out


# ------------------------------------------------------------------------------------------------ CLIP text tower pieces
def embed_tokens(ids, tok_emb, pos_emb):
    _need_cuda(ids, tok_emb, pos_emb)
    B, T = ids.shape
    Cc = tok_emb.shape[1]
    out = torch.empty((B * T, Cc), dtype=torch.float32, device=ids.device)
    _chk(_lib.load().avi_embed_tokens(_ptr(ids.contiguous().long()), _ptr(tok_emb), _ptr(pos_emb), _ptr(out), C.c_int32(B), C.c_int32(T),
                                      C.c_int32(Cc), C.c_int32(tok_emb.shape[0]), _stream()), "avi_embed_tokens")
    return out


def token_mean(x2d, B, T):
    _need_cuda(x2d)
    Cc = x2d.shape[-1]
    out = torch.empty((B, Cc), dtype=torch.float32, device=x2d.device)
    _chk(_lib.load().avi_token_mean(_ptr(x2d.contiguous()), _ptr(out), C.c_int32(B), C.c_int32(T), C.c_int32(Cc), _stream()), "avi_token_mean")
    return out


# ------------------------------------------------------------------------------------------------ FanEncoder image branch pieces
def round_tf32(t):
    """fp32 tensor rounded to nearest TF32 (10 explicit significand bits), for weights packed once; data movement level bit math."""
    u = t.detach().float().contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def im2col_affine(x, N, H, W, Cc, k, stride, pad, Kpad, dtype, pre=None, tf32=False, Wp_in=None, Wo_extra=0):
    """x fp32 rows [N*H*Wp_in, C] (row stride allowed) -> cols [N*Ho*(Wo+Wo_extra), Kpad] of `dtype`; pre = (scale, shift) applies
    relu(x*scale+shift); tf32 rounds the fp32 output to TF32."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.stride(1) == 1
    Wp_in = W if Wp_in is None else Wp_in
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1 + Wo_extra
    cols = torch.empty((N * Ho * Wo, Kpad), dtype=dtype, device=x.device)
    sc, sh = pre if pre is not None else (None, None)
    with _timed("im2col", float(cols.numel() * cols.element_size())):
        _chk(_lib.load().avi_im2col_affine(_ptr(x), C.c_int64(x.stride(0)), _ptr(cols), C.c_int32(2 if tf32 else _dt(cols)), C.c_int32(N), C.c_int32(H),
                                           C.c_int32(W), C.c_int32(Cc), C.c_int32(k), C.c_int32(stride), C.c_int32(pad), C.c_int32(Kpad),
                                           _ptr(sc), _ptr(sh), C.c_int32(Wp_in), C.c_int32(Wo_extra), _stream()), "avi_im2col_affine")
    return cols


def pad_act(x, N, H, W, Cc, dtype, pre=None, tf32=False):
    """x fp32 rows [N*H*(W+2), C] (padded-width layout, row stride allowed) -> activated, zero-bordered operand rows
    [N*(H+2)*(W+2) + 2, C] of `dtype` for the implicit 3x3 convolution (include/avi_b200.h: avi_pad_act)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.stride(1) == 1
    rows = N * (H + 2) * (W + 2)
    a = torch.empty((rows + 2, Cc), dtype=dtype, device=x.device)
    a[rows:].zero_()
    sc, sh = pre if pre is not None else (None, None)
    with _timed("pad_act", float(a.numel() * a.element_size())):
        _chk(_lib.load().avi_pad_act(_ptr(x), C.c_int64(x.stride(0)), _ptr(a), C.c_int32(2 if tf32 else _dt(a)), C.c_int32(N), C.c_int32(H),
                                     C.c_int32(W), C.c_int32(Cc), _ptr(sc), _ptr(sh), _stream()), "avi_pad_act")
    return a


def maxpool2x2(x, N, H, W, Cc, Wp_in=None, Wp_out=None):
    _need_cuda(x)
    Wp_in = W if Wp_in is None else Wp_in
    Wp_out = W // 2 if Wp_out is None else Wp_out
    y = torch.empty((N * (H // 2) * Wp_out, Cc), dtype=torch.float32, device=x.device)
    _chk(_lib.load().avi_maxpool2x2(_ptr(x.contiguous()), _ptr(y), C.c_int32(N), C.c_int32(H), C.c_int32(W), C.c_int32(Cc), C.c_int32(Wp_in),
                                    C.c_int32(Wp_out), _stream()), "avi_maxpool2x2")
    return y


def upsample_bilinear_add(low, up1, N, Hi, Wi, Ho, Wo, Cc, Wp_in=None, Wp_out=None):
    _need_cuda(low, up1)
    Wp_in = Wi if Wp_in is None else Wp_in
    Wp_out = Wo if Wp_out is None else Wp_out
    out = torch.empty_like(up1)
    _chk(_lib.load().avi_upsample_bilinear_add(_ptr(low.contiguous()), _ptr(up1.contiguous()), _ptr(out), C.c_int32(N), C.c_int32(Hi),
                                               C.c_int32(Wi), C.c_int32(Ho), C.c_int32(Wo), C.c_int32(Cc), C.c_int32(Wp_in), C.c_int32(Wp_out),
                                               _stream()), "avi_upsample_bilinear_add")
    return out


def affine_act(x, scale, shift, relu):
    _need_cuda(x, scale, shift)
    assert x.is_contiguous() and x.dtype == torch.float32
    _chk(_lib.load().avi_affine_act(_ptr(x), _ptr(scale), _ptr(shift), C.c_int64(x.numel() // x.shape[-1]), C.c_int32(x.shape[-1]),
                                    C.c_int32(1 if relu else 0), _stream()), "avi_affine_act")
    return x


# ------------------------------------------------------------------------------------------------ diffusion-prior training step
ACT_SILU = 4


def prior_tokens_fwd(brain, null_brain, keep_brain, x0, noise, sqrt_ac, sqrt_1mac, times_i32, null_image, keep_image, learned_query, temb):
    _need_cuda(brain, x0, noise, temb)
    B, dim = x0.shape
    tokens = torch.empty((B, 3, dim), dtype=torch.float32, device=x0.device)
    x_noisy = torch.empty((B, dim), dtype=torch.float32, device=x0.device)
    _chk(_lib.load().avi_prior_tokens_fwd(_ptr(brain), _ptr(null_brain), _ptr(keep_brain), _ptr(x0), _ptr(noise), _ptr(sqrt_ac), _ptr(sqrt_1mac),
                                          _ptr(times_i32), _ptr(null_image), _ptr(keep_image), _ptr(learned_query), _ptr(temb), _ptr(tokens),
                                          _ptr(x_noisy), C.c_int32(B), C.c_int32(dim), _stream()), "avi_prior_tokens_fwd")
    return tokens, x_noisy


def prior_tokens_bwd(dtokens, keep_brain, keep_image, dnull_brain, dnull_image, dlearned_query):
    """-> dbrain [B,dim], dtemb [B,dim]; the three parameter gradients are accumulated into the given fp32 buffers."""
    _need_cuda(dtokens)
    B, _, dim = dtokens.shape
    dbrain = torch.empty((B, dim), dtype=torch.float32, device=dtokens.device)
    dtemb = torch.empty((B, dim), dtype=torch.float32, device=dtokens.device)
    _chk(_lib.load().avi_prior_tokens_bwd(_ptr(dtokens.contiguous()), _ptr(keep_brain), _ptr(keep_image), _ptr(dbrain), _ptr(dtemb),
                                          _ptr(dnull_brain), _ptr(dnull_image), _ptr(dlearned_query), C.c_int32(B), C.c_int32(dim), _stream()),
         "avi_prior_tokens_bwd")
    return dbrain, dtemb


def prior_attn_fwd(q, kv, null_kv, rotary, rel_bias, B, heads=8, dim_head=64):
    _need_cuda(q, kv)
    out = torch.empty_like(q)
    P = torch.empty((B, heads, 3, 4), dtype=torch.float32, device=q.device)
    _chk(_lib.load().avi_prior_attn_fwd(_ptr(q), _ptr(kv), _ptr(null_kv), _ptr(rotary), _ptr(rel_bias), _ptr(out), _ptr(P), C.c_int32(B),
                                        C.c_int32(3), C.c_int32(heads), C.c_int32(dim_head), _stream()), "avi_prior_attn_fwd")
    return out, P


def prior_attn_bwd(q, kv, null_kv, rotary, P, dout, B, heads=8, dim_head=64):
    """-> dq, dkv, dnull_kv [2, dim_head], dbias [heads, 3, 4] (the per-sample partials folded by avi_colsum in a fixed order)."""
    _need_cuda(q, kv, P, dout)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    dnull = torch.empty((B, 2 * dim_head), dtype=torch.float32, device=q.device)
    dS = torch.empty((B, heads * 12), dtype=torch.float32, device=q.device)
    _chk(_lib.load().avi_prior_attn_bwd(_ptr(q), _ptr(kv), _ptr(null_kv), _ptr(rotary), _ptr(P), _ptr(dout.contiguous()), _ptr(dq), _ptr(dkv),
                                        _ptr(dnull), _ptr(dS), C.c_int32(B), C.c_int32(3), C.c_int32(heads), C.c_int32(dim_head), _stream()),
         "avi_prior_attn_bwd")
    return dq, dkv, colsum(dnull).view(2, dim_head), colsum(dS).view(heads, 3, 4)


def swiglu_fwd(h):
    _need_cuda(h)
    rows, two_i = h.shape
    y = torch.empty((rows, two_i // 2), dtype=torch.float32, device=h.device)
    _chk(_lib.load().avi_swiglu_fwd(_ptr(h), _ptr(y), C.c_int64(rows), C.c_int32(two_i // 2), _stream()), "avi_swiglu_fwd")
    return y


def swiglu_bwd(h, dy):
    _need_cuda(h, dy)
    dh = torch.empty_like(h)
    _chk(_lib.load().avi_swiglu_bwd(_ptr(h), _ptr(dy.contiguous()), _ptr(dh), C.c_int64(h.shape[0]), C.c_int32(h.shape[1] // 2), _stream()),
         "avi_swiglu_bwd")
    return dh


def rows_stat_div(x, mode):
    """mode 0: x / rowmax (LayerNorm(stable=True)), mode 1: F.normalize. -> (out, stat [rows])"""
    _need_cuda(x)
    x = x.contiguous()
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    out = torch.empty_like(x)
    stat = torch.empty((rows,), dtype=torch.float32, device=x.device)
    _chk(_lib.load().avi_rows_stat_div(_ptr(x), _ptr(out), _ptr(stat), C.c_int64(rows), C.c_int32(Cc), C.c_int32(mode), _stream()),
         "avi_rows_stat_div")
    return out, stat


def rows_stat_div_bwd(y, dy, stat, mode):
    _need_cuda(y, dy, stat)
    rows, Cc = y.numel() // y.shape[-1], y.shape[-1]
    dx = torch.empty_like(y)
    _chk(_lib.load().avi_rows_stat_div_bwd(_ptr(y), _ptr(dy.contiguous()), _ptr(stat), _ptr(dx), C.c_int64(rows), C.c_int32(Cc), C.c_int32(mode),
                                           _stream()), "avi_rows_stat_div_bwd")
    return dx


def mul_f32(a, b):
    _need_cuda(a, b)
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    y = torch.empty_like(a)
    _chk(_lib.load().avi_mul_f32(_ptr(a), _ptr(b), _ptr(y), C.c_int64(a.numel()), _stream()), "avi_mul_f32")
    return y


def scale_f32(a, alpha):
    _need_cuda(a)
    a = a.contiguous().float()
    y = torch.empty_like(a)
    _chk(_lib.load().avi_scale_f32(_ptr(a), C.c_float(alpha), _ptr(y), C.c_int64(a.numel()), _stream()), "avi_scale_f32")
    return y


def soft_clip_loss_grad(pt, tt, temp):
    """pt = preds targs^T, tt = targs targs^T [B,B] fp32 -> (loss fp64 device scalar, d loss / d pt [B,B])."""
    _need_cuda(pt, tt)
    B = pt.shape[0]
    dsim = torch.empty_like(pt)
    loss = torch.empty((1,), dtype=torch.float64, device=pt.device)
    scratch = torch.empty((3 * B,), dtype=torch.float32, device=pt.device)
    _chk(_lib.load().avi_soft_clip_loss_grad(_ptr(pt), _ptr(tt), _ptr(scratch), _ptr(dsim), _ptr(loss), C.c_int32(B), C.c_float(temp), _stream()),
         "avi_soft_clip_loss_grad")
    return loss, dsim


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    _need_cuda(p, g, m, v)
    _chk(_lib.load().avi_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), C.c_int64(p.numel()), C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
                                    C.c_float(eps), C.c_float(weight_decay), C.c_int32(step), C.c_float(grad_scale), _stream()), "avi_adamw_step")


class AviAdamwEntry(C.Structure):
    """Mirror of AviAdamwEntry in include/avi_b200.h."""
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64), ("weight_decay", C.c_float),
                ("reserved", C.c_int32)]


def adamw_table(items):
    """items: (param, grad, m, v, weight_decay) tensors -> device-resident uint8 tensor holding the AviAdamwEntry records."""
    arr = (AviAdamwEntry * len(items))()
    for i, (p, g, m, v, wd) in enumerate(items):
        _need_cuda(p, g, m, v)
        assert p.is_contiguous() and g.is_contiguous() and g.numel() == p.numel()
        arr[i] = AviAdamwEntry(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), wd, 0)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(items[0][0].device)


def adamw_multi(table, n_entries, lr, beta1, beta2, eps, step, grad_scale=1.0):
    _chk(_lib.load().avi_adamw_multi(_ptr(table), C.c_int32(n_entries), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                     C.c_int32(step), C.c_float(grad_scale), _stream()), "avi_adamw_multi")


def pack_disp_f16(verts2d, template):
    """fp32 vertices [rows, C] (row stride allowed) -> dense fp16 [rows, C] displacement from `template` [C] (opt-in compact sink)."""
    _need_cuda(verts2d, template)
    rows, Cc = verts2d.shape
    assert verts2d.dtype == torch.float32 and verts2d.stride(1) == 1 and template.numel() == Cc
    out = torch.empty((rows, Cc), dtype=torch.float16, device=verts2d.device)
    _chk(_lib.load().avi_pack_disp_f16(_ptr(verts2d), _ptr(template.contiguous().float()), _ptr(out), C.c_int64(rows), C.c_int32(Cc),
                                       C.c_int64(verts2d.stride(0)), _stream()), "avi_pack_disp_f16")
    return out
