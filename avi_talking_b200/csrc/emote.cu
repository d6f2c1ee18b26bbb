// Small kernels of the EMOTE talking-head decoder (Path B, third_party/inferno): everything that is not a GEMM, a LayerNorm
// or FLAME.
//   audio z-norm        Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm as called at inferno/models/temporal/AudioEncoders.py:170-178
//   small-head MHA      nn.TransformerEncoderLayer self-attention of BertPriorDecoder (8 x 16, no mask; FaceFormerDecoder.py:996-1002,
//                       1210) and of the L2L motion-prior decoder (8 x 32, additive -slope_h*|i-j| bias built on the fly;
//                       L2lMotionPrior.py:390-398,474-483, TransformerMasking.py:80-98)
//   row staging         zero / replicate padding and zero-insertion, so Conv1d(k=5) and ConvTranspose1d(k=5,s=2) of the L2L
//                       expander run as conv-mode GEMMs (L2lMotionPrior.py:368-389,465-470)
//   LeakyReLU+BN(+x2)   the expander's activation, eval-mode BatchNorm1d and repeat_interleave(2) in one pass
#include "common.cuh"

namespace avi {

// ------------------------------------------------------------------------------------------------ audio z-norm
__global__ void __launch_bounds__(1024) znorm_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float eps) {
  __shared__ double red[64];
  const float* xb = x + (int64_t)blockIdx.x * n;
  float* yb = y + (int64_t)blockIdx.x * n;
  double s = 0.0, ss = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = xb[i];
    s += v;
    ss += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    red[w] = s;
    red[32 + w] = ss;
  }
  __syncthreads();
  if (w == 0) {
    s = l < (blockDim.x >> 5) ? red[l] : 0.0;
    ss = l < (blockDim.x >> 5) ? red[32 + l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (l == 0) {
      red[0] = s;
      red[1] = ss;
    }
  }
  __syncthreads();
  const double mean = red[0] / (double)n;
  const double var = fmax(red[1] / (double)n - mean * mean, 0.0);
  const float fm = (float)mean, rs = (float)(1.0 / sqrt(var + (double)eps));
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) yb[i] = (xb[i] - fm) * rs;
}

// ------------------------------------------------------------------------------------------------ small-head attention
// qkv fp32 [B, T, 3*H*D] (q | k | v); one thread per query row, 128 rows per block, key tiles of 32 through shared memory.
constexpr int SA_ROWS = 128, SA_KT = 32;

template <int D>
__global__ void __launch_bounds__(SA_ROWS) mha_small_kernel(const float* __restrict__ qkv, float* __restrict__ out_f32,
                                                             __nv_bfloat16* __restrict__ out_bf16, int T, int H, float scale,
                                                             const float* __restrict__ slopes) {
  __shared__ float Ks[SA_KT][D], Vs[SA_KT][D];
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * SA_ROWS + threadIdx.x;
  const int E = H * D;
  const float* base = qkv + (int64_t)b * T * 3 * E;
  const bool valid = i < T;
  float q[D], acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    q[d] = valid ? base[(int64_t)i * 3 * E + h * D + d] * scale : 0.f;
    acc[d] = 0.f;
  }
  const float slope = slopes ? slopes[h] : 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < T; j0 += SA_KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < SA_KT * D; idx += SA_ROWS) {
      const int jj = idx / D, d = idx % D;
      const int j = j0 + jj;
      Ks[jj][d] = j < T ? base[(int64_t)j * 3 * E + E + h * D + d] : 0.f;
      Vs[jj][d] = j < T ? base[(int64_t)j * 3 * E + 2 * E + h * D + d] : 0.f;
    }
    __syncthreads();
    const int nk = min(SA_KT, T - j0);
    float s[SA_KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < SA_KT; ++jj) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(q[d], Ks[jj][d], a);
      a -= slope * fabsf((float)(i - (j0 + jj)));
      s[jj] = jj < nk ? a : -INFINITY;
      tmax = fmaxf(tmax, s[jj]);
    }
    const float mn = fmaxf(m, tmax);
    const float corr = __expf(m - mn);  // exp(-inf) = 0 on the first tile
    l *= corr;
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] *= corr;
#pragma unroll
    for (int jj = 0; jj < SA_KT; ++jj) {
      const float p = expf(s[jj] - mn);  // masked tail: exp(-inf) = 0
      l += p;
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = fmaf(p, Vs[jj][d], acc[d]);
    }
    m = mn;
  }
  if (valid) {
    const float inv = 1.f / l;
    const int64_t o = ((int64_t)b * T + i) * E + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float y = acc[d] * inv;
      if (out_f32) out_f32[o + d] = y;
      if (out_bf16) out_bf16[o + d] = __float2bfloat16_rn(y);
    }
  }
}

// ------------------------------------------------------------------------------------------------ row staging
// mode 0: dst[b, front + t] = src[b, t], zeros elsewhere            (zero padding)
// mode 1: dst[b, r] = src[b, clamp(r - front, 0, L-1)]              (replicate padding)
// mode 2: dst[b, front + 2 t] = src[b, t], zeros elsewhere          (zero insertion for ConvTranspose1d stride 2)
template <typename OutT>
__global__ void stage_rows_kernel(const float* __restrict__ src, OutT* __restrict__ dst, int L, int Lp, int C, int front, int mode,
                                  int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int r = (int)((idx / C) % Lp);
  const int64_t b = idx / ((int64_t)C * Lp);
  float v = 0.f;
  if (mode == 1) {
    int t = r - front;
    t = t < 0 ? 0 : (t > L - 1 ? L - 1 : t);
    v = src[(b * L + t) * C + c];
  } else if (mode == 0) {
    const int t = r - front;
    if (t >= 0 && t < L) v = src[(b * L + t) * C + c];
  } else {
    const int t2 = r - front;
    if (t2 >= 0 && (t2 & 1) == 0 && (t2 >> 1) < L) v = src[(b * L + (t2 >> 1)) * C + c];
  }
  if constexpr (sizeof(OutT) == 2) dst[idx] = __float2bfloat16_rn(v);
  else dst[idx] = v;
}

// y[b, rep*t + u, c] = bn_scale[c] * leaky_relu(x[b, t, c], 0.2) + bn_shift[c],  u < rep   (rep = 1 or 2)
__global__ void lrelu_bn_repeat_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                       float* __restrict__ y, int L, int C, int rep, float slope, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int t = (int)((idx / C) % L);
  const int64_t b = idx / ((int64_t)C * L);
  float v = x[idx];
  v = v > 0.f ? v : v * slope;
  v = fmaf(v, scale[c], shift[c]);
  for (int u = 0; u < rep; ++u) y[((b * L + t) * rep + u) * C + c] = v;
}

}  // namespace avi

using namespace avi;

extern "C" int avi_audio_znorm(const float* x, float* y, int32_t B, int64_t n, float eps, void* stream) {
  AVI_REQUIRE(B > 0 && n > 0, "avi_audio_znorm: bad sizes");
  znorm_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(x, y, n, eps);
  return check_launch("audio_znorm");
}

extern "C" int avi_mha_small_fwd(const float* qkv, float* out_f32, void* out_bf16, int32_t B, int32_t T, int32_t H, int32_t D,
                                 float scale, const float* slopes, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && H > 0, "avi_mha_small_fwd: bad sizes");
  AVI_REQUIRE(D == 16 || D == 32, "avi_mha_small_fwd: head dim must be 16 or 32 (got %d)", D);
  dim3 grid((T + SA_ROWS - 1) / SA_ROWS, H, B);
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (D == 16) mha_small_kernel<16><<<grid, SA_ROWS, 0, (cudaStream_t)stream>>>(qkv, out_f32, o16, T, H, scale, slopes);
  else mha_small_kernel<32><<<grid, SA_ROWS, 0, (cudaStream_t)stream>>>(qkv, out_f32, o16, T, H, scale, slopes);
  return check_launch("mha_small");
}

extern "C" int avi_stage_rows(const float* src, void* dst, int32_t dst_dtype, int32_t B, int32_t L, int32_t Lp, int32_t C, int32_t front,
                              int32_t mode, void* stream) {
  AVI_REQUIRE(B > 0 && L > 0 && Lp > 0 && C > 0 && mode >= 0 && mode <= 2, "avi_stage_rows: bad arguments");
  const int64_t total = (int64_t)B * Lp * C;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (dst_dtype == AVI_DT_BF16)
    stage_rows_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), L, Lp, C, front, mode, total);
  else
    stage_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<float*>(dst), L, Lp, C, front, mode, total);
  return check_launch("stage_rows");
}

extern "C" int avi_lrelu_bn_repeat(const float* x, const float* bn_scale, const float* bn_shift, float* y, int32_t B, int32_t L, int32_t C,
                                   int32_t repeat, float slope, void* stream) {
  AVI_REQUIRE(B > 0 && L > 0 && C > 0 && (repeat == 1 || repeat == 2), "avi_lrelu_bn_repeat: bad arguments");
  const int64_t total = (int64_t)B * L * C;
  lrelu_bn_repeat_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, bn_scale, bn_shift, y, L, C, repeat, slope, total);
  return check_launch("lrelu_bn_repeat");
}

namespace avi {
// out[b, t, c] = (a[b, t, c] - n[b, c]) + tpl[b, c]   (offsets from the neutral shape re-attached to the template,
// FaceFormerDecoder.py:1173-1175 and :690-694), in place allowed
__global__ void sub_add_rows_kernel(const float* a, const float* __restrict__ n, const float* __restrict__ tpl, float* out, int T, int C,
                                    int64_t ld, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int64_t row = idx / C;
  const int64_t b = row / T;
  out[row * ld + c] = (a[row * ld + c] - n[b * C + c]) + tpl[b * C + c];
}
}  // namespace avi

extern "C" int avi_sub_add_rows(const float* a, const float* neutral, const float* tpl, float* out, int32_t B, int32_t T, int32_t C,
                                int64_t row_stride, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && C > 0 && row_stride >= C, "avi_sub_add_rows: bad sizes");
  const int64_t total = (int64_t)B * T * C;
  avi::sub_add_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, neutral, tpl, out, T, C, row_stride, total);
  return avi::check_launch("sub_add_rows");
}
