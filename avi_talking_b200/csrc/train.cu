// Backward / optimizer kernels of the teacher-forced FaceFormer-vert training step (BASELINE configs[4];
// models/faceformer_vert.py:360-482: wav2vec2 encoder (feature extractor frozen, :154) -> audio_feature_map -> teacher-forced
// nn.TransformerDecoderLayer -> vertice_map_r -> MSE x 10). The dense contractions of the backward pass (dX = dY W, dW = dY^T X) run
// on the tcgen05 GEMM of gemm_tc2.cu after avi_transpose_cast_bf16 has put the contraction dimension innermost; this file holds
// everything else: LayerNorm / GELU / ReLU / softmax-attention backward, bias column sums, the positional-conv weight gradient
// with its weight-norm chain rule, the loss and Adam. Clips are short (T ~ 120 frames), so these are small fp32 kernels.
#include "common.cuh"

namespace avi {

// ------------------------------------------------------------------------------------------------ layout helpers
// dst[c][r] = bf16(src[r][c]) for r < R, zero for R <= r < R_pad   (src fp32 [R, C] with row stride src_ld)
template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename OutT>
__global__ void transpose_cast_kernel(const float* __restrict__ src, OutT* __restrict__ dst, int R, int C, int64_t src_ld, int R_pad) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[(int64_t)r * src_ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R_pad) dst[(int64_t)c * R_pad + r] = to_out<OutT>(tile[threadIdx.x][i]);
  }
}

// dst[r][c] = src[r][c] for r < R, c < C, zero elsewhere in [R_pad, C_pad]
template <typename OutT>
__global__ void cast_pad2d_kernel(const float* __restrict__ src, OutT* __restrict__ dst, int R, int C, int64_t src_ld, int R_pad, int C_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R_pad * C_pad) return;
  const int r = (int)(i / C_pad), c = (int)(i % C_pad);
  dst[i] = to_out<OutT>((r < R && c < C) ? src[(int64_t)r * src_ld + c] : 0.f);
}

// teacher-forcing input rows (faceformer_vert.py:443-444): row (b, t) = gt[b, t-1] - template for t >= 1, zero for t = 0
__global__ void tf_input_rows_kernel(const float* __restrict__ gt, int64_t gt_ld, const float* __restrict__ tpl, float* __restrict__ out,
                                     int B, int T, int C, int C_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T * C_pad) return;
  const int c = (int)(i % C_pad);
  const int64_t row = i / C_pad;
  const int t = (int)(row % T);
  // cat([template, gt[:, :-1]]) - template : the first row is template - template
  out[i] = (c < C && t > 0) ? gt[(row - 1) * gt_ld + c] - tpl[c] : 0.f;
}

// x[b, t, :] += style[b or 0, :] + pe[t mod period, :]   (faceformer_vert.py:446-447, PeriodicPositionalEncoding :92-107)
__global__ void add_style_pe_kernel(float* __restrict__ x, const float* __restrict__ style, int64_t style_stride, const float* __restrict__ pe,
                                    int B, int T, int fd, int period) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T * fd) return;
  const int c = (int)(i % fd);
  const int t = (int)((i / fd) % T);
  const int b = (int)(i / ((int64_t)fd * T));
  x[i] += style[b * style_stride + c] + pe[(t % period) * fd + c];
}

// out[n] (+)= sum_r x[r][n]. Block (32 columns x 8 row lanes); gridDim.y splits the rows (atomics into a zeroed / accumulating out)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int R, int N, int64_t ld,
                                                     int rows_per_block, int use_atomic, int accumulate) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += x[(int64_t)r * ld + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
#pragma unroll
    for (int y = 1; y < 8; ++y) s += red[y][threadIdx.x];
    if (use_atomic) atomicAdd(out + n, s);
    else out[n] = accumulate ? out[n] + s : s;
  }
}

// ------------------------------------------------------------------------------------------------ activations
// mode 1 GELU (exact erf), 2 ReLU. fwd: out = act(pre) (fp32 and/or bf16); bwd: dpre = dout * act'(pre)
__global__ void act_fwd_kernel(const float* __restrict__ pre, float* __restrict__ o32, __nv_bfloat16* __restrict__ o16, int64_t n, int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pre[i];
  const float y = mode == AVI_ACT_GELU ? gelu_erf(x) : mode == AVI_ACT_SILU ? x / (1.f + expf(-x)) : fmaxf(x, 0.f);
  if (o32) o32[i] = y;
  if (o16) o16[i] = __float2bfloat16_rn(y);
}
__global__ void act_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dout, float* __restrict__ dpre, int64_t n, int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pre[i];
  float g;
  if (mode == 1) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    g = cdf + x * 0.3989422804014327f * expf(-0.5f * x * x);
  } else if (mode == AVI_ACT_SILU) {
    const float sg = 1.f / (1.f + expf(-x));
    g = sg * (1.f + x * (1.f - sg));
  } else {
    g = x > 0.f ? 1.f : 0.f;
  }
  dpre[i] = dout[i] * g;
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// y = (x - mean) * rstd * w + b over the last dim C (<= 1024). One warp per row:
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w ;  dw += sum_rows dy * xhat ; db += sum_rows dy  (atomics)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dw,
                                                            float* __restrict__ db, int64_t rows, int C, float eps) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[32], g[32];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = lane + 32 * u;
    v[u] = c < C ? x[row * C + c] : 0.f;
    s += v[u];
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = lane + 32 * u;
    if (c < C) q += (v[u] - mean) * (v[u] - mean);
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = lane + 32 * u;
    g[u] = 0.f;
    if (c < C) {
      const float xh = (v[u] - mean) * rstd;
      const float d = dy[row * C + c];
      g[u] = d * w[c];
      sg += g[u];
      sgx += g[u] * xh;
      if (dw) atomicAdd(dw + c, d * xh);
      if (db) atomicAdd(db + c, d);
      v[u] = xh;
    }
  }
  sg = warp_sum(sg) / C;
  sgx = warp_sum(sgx) / C;
  if (dx) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const int c = lane + 32 * u;
      if (c < C) dx[row * C + c] = rstd * (g[u] - sg - v[u] * sgx);
    }
  }
}

// Wide rows (1024 < C <= 8192): one 256-thread block per row, same formulas.
__global__ void __launch_bounds__(256) layernorm_bwd_wide_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ dy, float* __restrict__ dx,
                                                                 float* __restrict__ dw, float* __restrict__ db, int C, float eps) {
  __shared__ float red[64];
  const int64_t row = blockIdx.x;
  float v[32], g[32];
  float s = 0.f, unused = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = threadIdx.x + 256 * u;
    v[u] = c < C ? x[row * C + c] : 0.f;
    s += v[u];
  }
  block_sum2(s, unused, red);
  const float mean = s / C;
  float q = 0.f;
  unused = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = threadIdx.x + 256 * u;
    if (c < C) q += (v[u] - mean) * (v[u] - mean);
  }
  block_sum2(q, unused, red);
  const float rstd = rsqrtf(q / C + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = threadIdx.x + 256 * u;
    g[u] = 0.f;
    if (c < C) {
      const float xh = (v[u] - mean) * rstd;
      const float d = dy[row * C + c];
      g[u] = d * w[c];
      sg += g[u];
      sgx += g[u] * xh;
      if (dw) atomicAdd(dw + c, d * xh);
      if (db) atomicAdd(db + c, d);
      v[u] = xh;
    }
  }
  block_sum2(sg, sgx, red);
  sg /= C;
  sgx /= C;
  if (dx) {
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      const int c = threadIdx.x + 256 * u;
      if (c < C) dx[row * C + c] = rstd * (g[u] - sg - v[u] * sgx);
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention (training, T <= 128)
// qkv fp32 [B, T, 3*H*D] (q | k | v); bias modes: 0 none, 1 FaceFormer biased causal mask (-slope_h * floor((i-j)/period), j <= i).
// forward keeps the probabilities P [B, H, T, T] for the backward pass. Clips are short (T ~ 120) and there are only B*H
// (head, clip) problems, so every kernel is tiled over ATT_ROWS query (or key) rows as well: grid = (H, B, ceil(T / ATT_ROWS)),
// K / V (or Q / dO) of the head staged in shared memory with a D+1 pitch (conflict-free both along keys and along channels).
__device__ __forceinline__ float ff_slope(int h) { return exp2f(-2.f * (float)(h + 1)); }   // 4 heads: 2^-2, 2^-4, 2^-6, 2^-8
constexpr int ATT_ROWS = 16;      // rows per CTA: 8 warps x 2 rows, processed together so that every K / V (Q / dO) load feeds two rows
constexpr int ATT_MAXT = 128;
constexpr int ATT_PAD = 4;        // row pitch HD + 4 floats: 16-byte aligned rows, and lanes = consecutive rows hit distinct bank groups

template <int HD>
__device__ __forceinline__ void att_stage(float* dst, const float* __restrict__ src, int T, int row_stride) {
  // dst[t][d] (pitch HD + ATT_PAD) = src[t * row_stride + d]
  for (int i = threadIdx.x; i < T * (HD / 4); i += blockDim.x) {
    const int t = i / (HD / 4), d4 = i - t * (HD / 4);
    *reinterpret_cast<float4*>(dst + t * (HD + ATT_PAD) + 4 * d4) = *reinterpret_cast<const float4*>(src + (int64_t)t * row_stride + 4 * d4);
  }
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// scores of two query rows (vectors xa, xb in shared memory) against the staged rows `mat`, keys lane, lane+32, ...
template <int HD>
__device__ __forceinline__ void att_two_row_scores(const float* mat, const float* xa, const float* xb, int T, int lane, float (&sa)[4],
                                                   float (&sb)[4]) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = lane + 32 * u;
    sa[u] = 0.f;
    sb[u] = 0.f;
    if (j < T) {
      const float* kr = mat + j * (HD + ATT_PAD);
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
        a = dot4(k4, *reinterpret_cast<const float4*>(xa + d), a);
        b = dot4(k4, *reinterpret_cast<const float4*>(xb + d), b);
      }
      sa[u] = a;
      sb[u] = b;
    }
  }
}

// ya[d] = sum_j pa[j] mat[j][d], yb likewise, for d = lane, lane + 32 (< HD); pa / pb in shared memory, zero padded to a multiple of 4
template <int HD>
__device__ __forceinline__ void att_two_row_mix(const float* mat, const float* pa, const float* pb, int T, int lane, float (&ya)[2],
                                                float (&yb)[2]) {
  ya[0] = ya[1] = yb[0] = yb[1] = 0.f;
  const int T4 = (T + 3) & ~3;
  for (int j = 0; j < T4; j += 4) {
    const float4 a4 = *reinterpret_cast<const float4*>(pa + j), b4 = *reinterpret_cast<const float4*>(pb + j);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      if (j + jj < T) {
#pragma unroll
        for (int h = 0; h < (HD + 31) / 32; ++h) {
          const int d = lane + 32 * h;
          if (d < HD) {
            const float m = mat[(j + jj) * (HD + ATT_PAD) + d];
            ya[h] = fmaf(av[jj], m, ya[h]);
            yb[h] = fmaf(bv[jj], m, yb[h]);
          }
        }
      }
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(256) attn_train_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ P,
                                                             int T, int H, float scale, int bias_mode, int period,
                                                             const float* __restrict__ pmask) {
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;                              // [T][HD+4]
  float* vs = ks + T * (HD + ATT_PAD);         // [T][HD+4]
  float* qs = vs + T * (HD + ATT_PAD);         // [8 warps][2][HD]
  float* ps = qs + 8 * 2 * HD;                 // [8 warps][2][ATT_MAXT]
  const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * ATT_ROWS;
  const int E = H * HD;
  const float* base = qkv + (int64_t)b * T * 3 * E + h * HD;
  att_stage<HD>(ks, base + E, T, 3 * E);
  att_stage<HD>(vs, base + 2 * E, T, 3 * E);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Pb = P ? P + ((int64_t)b * H + h) * T * T : nullptr;   // inference callers (CLIP text tower) do not keep the probabilities
  // train mode: the (pre-scaled) dropout mask on the probabilities, nn.MultiheadAttention / Wav2Vec2Attention dropout. P keeps the
  // softmax itself; the mask multiplies what meets V
  const float* Mb = pmask ? pmask + ((int64_t)b * H + h) * T * T : nullptr;
  float* qa = qs + warp * 2 * HD, *qb = qa + HD;
  float* pa = ps + warp * 2 * ATT_MAXT, *pb = pa + ATT_MAXT;
  const float slope = ff_slope(h);
  const int ia = i0 + warp, ib = i0 + warp + 8;
  if (ia >= T) return;                         // warp-uniform; no block-wide barrier follows
  const bool b_ok = ib < T;
  for (int d = lane; d < HD; d += 32) {
    qa[d] = base[(int64_t)ia * 3 * E + d] * scale;
    qb[d] = b_ok ? base[(int64_t)ib * 3 * E + d] * scale : 0.f;
  }
  __syncwarp();
  float sa[4], sb[4];
  att_two_row_scores<HD>(ks, qa, qb, T, lane, sa, sb);
  float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = lane + 32 * u;
    const bool va = j < T && !(bias_mode != 0 && j > ia), vb = j < T && !(bias_mode != 0 && j > ib);
    if (bias_mode == 1) {
      sa[u] -= slope * (float)((ia - j) / period);
      sb[u] -= slope * (float)((ib - j) / period);
    }
    sa[u] = va ? sa[u] : -INFINITY;
    sb[u] = vb ? sb[u] : -INFINITY;
    ma = fmaxf(ma, sa[u]);
    mb = fmaxf(mb, sb[u]);
  }
  ma = warp_max(ma);
  mb = warp_max(mb);
  float da = 0.f, db = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    sa[u] = (sa[u] == -INFINITY) ? 0.f : expf(sa[u] - ma);
    sb[u] = (sb[u] == -INFINITY) ? 0.f : expf(sb[u] - mb);
    da += sa[u];
    db += sb[u];
  }
  da = 1.f / warp_sum(da);
  db = 1.f / fmaxf(warp_sum(db), 1e-30f);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = lane + 32 * u;        // j < ATT_MAXT always: the padding up to a multiple of 4 is written as zeros
    const float ka = (Mb != nullptr && j < T) ? Mb[(int64_t)ia * T + j] : 1.f;
    const float kb = (Mb != nullptr && j < T && b_ok) ? Mb[(int64_t)ib * T + j] : 1.f;
    pa[j] = sa[u] * da * ka;
    pb[j] = sb[u] * db * kb;
    if (Pb != nullptr && j < T) {
      Pb[(int64_t)ia * T + j] = sa[u] * da;
      if (b_ok) Pb[(int64_t)ib * T + j] = sb[u] * db;
    }
  }
  __syncwarp();
  float ya[2], yb[2];
  att_two_row_mix<HD>(vs, pa, pb, T, lane, ya, yb);
#pragma unroll
  for (int hh = 0; hh < (HD + 31) / 32; ++hh) {
    const int d = lane + 32 * hh;
    if (d < HD) {
      out[((int64_t)b * T + ia) * E + h * HD + d] = ya[hh];
      if (b_ok) out[((int64_t)b * T + ib) * E + h * HD + d] = yb[hh];
    }
  }
}

// backward, query-tiled half: dP = dO V^T ; dS = P * (dP - rowsum(dP * P)) -> scratch [B,H,T,T] ; dQ = scale dS K
template <int HD>
__global__ void __launch_bounds__(256) attn_train_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                               const float* __restrict__ dout, float* __restrict__ dqkv,
                                                               float* __restrict__ dS, int T, int H, float scale,
                                                               const float* __restrict__ pmask) {
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;
  float* vs = ks + T * (HD + ATT_PAD);
  float* gs = vs + T * (HD + ATT_PAD);         // dO rows of the two queries per warp [8][2][HD]
  float* ds = gs + 8 * 2 * HD;                 // dS rows [8][2][ATT_MAXT]
  const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * ATT_ROWS;
  const int E = H * HD;
  const float* base = qkv + (int64_t)b * T * 3 * E + h * HD;
  att_stage<HD>(ks, base + E, T, 3 * E);
  att_stage<HD>(vs, base + 2 * E, T, 3 * E);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* Pb = P + ((int64_t)b * H + h) * T * T;
  const float* Mb = pmask ? pmask + ((int64_t)b * H + h) * T * T : nullptr;
  float* dSb = dS + ((int64_t)b * H + h) * T * T;
  float* ga = gs + warp * 2 * HD, *gb = ga + HD;
  float* ra = ds + warp * 2 * ATT_MAXT, *rb = ra + ATT_MAXT;
  const int ia = i0 + warp, ib = i0 + warp + 8;
  if (ia >= T) return;
  const bool b_ok = ib < T;
  for (int d = lane; d < HD; d += 32) {
    ga[d] = dout[((int64_t)b * T + ia) * E + h * HD + d];
    gb[d] = b_ok ? dout[((int64_t)b * T + ib) * E + h * HD + d] : 0.f;
  }
  __syncwarp();
  float pa[4], pb[4], qa[4], qb[4];
  att_two_row_scores<HD>(vs, ga, gb, T, lane, pa, pb);      // dP rows
  float dota = 0.f, dotb = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = lane + 32 * u;
    qa[u] = j < T ? Pb[(int64_t)ia * T + j] : 0.f;
    qb[u] = (j < T && b_ok) ? Pb[(int64_t)ib * T + j] : 0.f;
    if (Mb != nullptr) {       // dropout on the probabilities: the gradient reaching the softmax is dP' * mask
      pa[u] *= j < T ? Mb[(int64_t)ia * T + j] : 0.f;
      pb[u] *= (j < T && b_ok) ? Mb[(int64_t)ib * T + j] : 0.f;
    }
    dota = fmaf(pa[u], qa[u], dota);
    dotb = fmaf(pb[u], qb[u], dotb);
  }
  dota = warp_sum(dota);
  dotb = warp_sum(dotb);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = lane + 32 * u;
    const float va = qa[u] * (pa[u] - dota), vb = qb[u] * (pb[u] - dotb);
    ra[j] = va;
    rb[j] = vb;
    if (j < T) {
      dSb[(int64_t)ia * T + j] = va;
      if (b_ok) dSb[(int64_t)ib * T + j] = vb;
    }
  }
  __syncwarp();
  float ya[2], yb[2];
  att_two_row_mix<HD>(ks, ra, rb, T, lane, ya, yb);
#pragma unroll
  for (int hh = 0; hh < (HD + 31) / 32; ++hh) {
    const int d = lane + 32 * hh;
    if (d < HD) {
      dqkv[((int64_t)b * T + ia) * 3 * E + h * HD + d] = ya[hh] * scale;
      if (b_ok) dqkv[((int64_t)b * T + ib) * 3 * E + h * HD + d] = yb[hh] * scale;
    }
  }
}

// backward, key-tiled half: dK[j] = scale sum_i dS[i,j] Q[i] ; dV[j] = sum_i P[i,j] dO[i]   (two keys per warp)
template <int HD>
__global__ void __launch_bounds__(256) attn_train_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                                const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                const float* __restrict__ dS, int T, int H, float scale,
                                                                const float* __restrict__ pmask) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                               // Q  [T][HD+4]
  float* gs = qs + T * (HD + ATT_PAD);          // dO [T][HD+4]
  float* dst = gs + T * (HD + ATT_PAD);         // dS^T tile [ATT_ROWS][ATT_MAXT]
  float* pst = dst + ATT_ROWS * ATT_MAXT;       // P^T tile  [ATT_ROWS][ATT_MAXT]
  const int h = blockIdx.x, b = blockIdx.y, j0 = blockIdx.z * ATT_ROWS;
  const int E = H * HD;
  const float* base = qkv + (int64_t)b * T * 3 * E + h * HD;
  att_stage<HD>(qs, base, T, 3 * E);
  att_stage<HD>(gs, dout + (int64_t)b * T * E + h * HD, T, E);
  const float* Pb = P + ((int64_t)b * H + h) * T * T;
  const float* dSb = dS + ((int64_t)b * H + h) * T * T;
  const float* Mb = pmask ? pmask + ((int64_t)b * H + h) * T * T : nullptr;
  for (int idx = threadIdx.x; idx < ATT_MAXT * ATT_ROWS; idx += blockDim.x) {
    const int i = idx / ATT_ROWS, jj = idx - i * ATT_ROWS;
    const int j = j0 + jj;
    const bool ok = i < T && j < T;
    dst[jj * ATT_MAXT + i] = ok ? dSb[(int64_t)i * T + j] : 0.f;
    pst[jj * ATT_MAXT + i] = ok ? Pb[(int64_t)i * T + j] * (Mb != nullptr ? Mb[(int64_t)i * T + j] : 1.f) : 0.f;   // dV sees P * mask
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ja = j0 + warp, jb = j0 + warp + 8;
  if (ja >= T) return;
  const bool b_ok = jb < T;
  float ka[2], kb[2], va[2], vb[2];
  att_two_row_mix<HD>(qs, dst + warp * ATT_MAXT, dst + (warp + 8) * ATT_MAXT, T, lane, ka, kb);
  att_two_row_mix<HD>(gs, pst + warp * ATT_MAXT, pst + (warp + 8) * ATT_MAXT, T, lane, va, vb);
#pragma unroll
  for (int hh = 0; hh < (HD + 31) / 32; ++hh) {
    const int d = lane + 32 * hh;
    if (d < HD) {
      dqkv[((int64_t)b * T + ja) * 3 * E + E + h * HD + d] = ka[hh] * scale;
      dqkv[((int64_t)b * T + ja) * 3 * E + 2 * E + h * HD + d] = va[hh];
      if (b_ok) {
        dqkv[((int64_t)b * T + jb) * 3 * E + E + h * HD + d] = kb[hh] * scale;
        dqkv[((int64_t)b * T + jb) * 3 * E + 2 * E + h * HD + d] = vb[hh];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ positional conv (weight-normed, grouped)
// dW[co][ci][j] = sum_t dpc[b, t, co] * xpad[b, t + j, g*CG + ci],  xpad row r = x row r - pad (zero outside), co in group g
__global__ void __launch_bounds__(256) posconv_dw_kernel(const float* __restrict__ x, const float* __restrict__ dpc, float* __restrict__ dw,
                                                         int B, int T, int C, int CG, int k) {
  const int j = blockIdx.x, g = blockIdx.y;
  const int pad = k / 2;
  for (int o = threadIdx.x; o < CG * CG; o += blockDim.x) {
    const int co = o / CG, ci = o % CG;
    float a = 0.f;
    for (int b = 0; b < B; ++b)
      for (int t = 0; t < T; ++t) {
        const int r = t + j - pad;
        if (r >= 0 && r < T) a = fmaf(dpc[((int64_t)b * T + t) * C + g * CG + co], x[((int64_t)b * T + r) * C + g * CG + ci], a);
      }
    dw[((int64_t)(g * CG + co) * CG + ci) * k + j] = a;
  }
}

// Transposed unfold for the tensor-core weight gradient of the positional conv: for a block of CW channels starting at c0,
//   out[(j * CW + ci), b * Tq + t] = xpad[b, t + j, c0 + ci]   (t < T; zero for T <= t < Tq),   xpad row r = x row r - pad (zero outside)
// so that dW_block[co, (j, ci)] = sum_{b,t} dpc[b, t, co] * out[(j, ci), (b, t)] is ONE GEMM with the (clip, frame) index as the
// contraction. Tiles of 32 channels x 32 frames go through shared memory: reads coalesced over channels, writes over frames.
template <typename OutT>
__global__ void __launch_bounds__(256) posconv_unfold_t_kernel(const float* __restrict__ x, OutT* __restrict__ out, int B, int T, int Tq,
                                                               int C, int c0, int CW, int k) {
  __shared__ float tile[32][33];
  const int j = blockIdx.z % k, b = blockIdx.z / k;
  const int ci0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  const int pad = k / 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, r = t + j - pad, ci = ci0 + tx;
    tile[i][tx] = (t < T && r >= 0 && r < T && ci < CW) ? x[((int64_t)b * T + r) * C + c0 + ci] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int ci = ci0 + i, t = t0 + tx;
    if (ci < CW && t < Tq) out[((int64_t)j * CW + ci) * ((int64_t)B * Tq) + (int64_t)b * Tq + t] = to_out<OutT>(tile[tx][i]);
  }
}

// w = g[j] * v / ||v[:, :, j]||  (norm over the first two dims): dg[j] = sum dw v / n ; dv = (g/n) dw - (g/n^3) (sum dw v) v
__global__ void __launch_bounds__(256) weightnorm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ dw,
                                                             float* __restrict__ dv, float* __restrict__ dg, int n_rows, int k) {
  __shared__ float red[64];
  const int j = blockIdx.x;
  float nn = 0.f, dot = 0.f;
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
    const float vv = v[(int64_t)r * k + j];
    nn += vv * vv;
    dot += dw[(int64_t)r * k + j] * vv;
  }
  block_sum2(nn, dot, red);
  const float n = sqrtf(nn);
  const float gj = g[j];
  if (threadIdx.x == 0) dg[j] = dot / n;
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
    const float vv = v[(int64_t)r * k + j];
    dv[(int64_t)r * k + j] = gj / n * dw[(int64_t)r * k + j] - gj * dot / (n * n * n) * vv;
  }
}

// ------------------------------------------------------------------------------------------------ loss, optimizer, misc
// loss = mean((out - gt)^2) * loss_scale ; dout = 2 * loss_scale / n * (out - gt)
__global__ void __launch_bounds__(256) mse_loss_grad_kernel(const float* __restrict__ out, const float* __restrict__ gt, float* __restrict__ dout,
                                                            double* __restrict__ loss, int64_t rows, int C, int64_t out_ld, int64_t gt_ld,
                                                            float loss_scale) {
  __shared__ float red[64];
  const int64_t n = rows * C;
  float acc = 0.f, dummy = 0.f;
  const float k = 2.f * loss_scale / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i % C);
    const float d = out[r * out_ld + c] - gt[r * gt_ld + c];
    acc = fmaf(d, d, acc);
    dout[i] = k * d;
  }
  block_sum2(acc, dummy, red);
  if (threadIdx.x == 0) atomicAdd(loss, (double)acc * (double)loss_scale / (double)n);
}

// 4 elements per thread (16-byte loads / stores: the kernel is pure HBM streaming, 28 B per parameter + 2 B for the bf16 shadow)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, __nv_bfloat16* __restrict__ p16, int64_t n, float lr, float b1,
                                                   float b2, float eps, float bc1, float bc2, float grad_scale) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  const float step = lr / bc1, rs2 = rsqrtf(bc2);
  if (i4 + 4 <= n) {
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<const float4*>(m + i4), vv = *reinterpret_cast<const float4*>(v + i4);
    float4 pp = *reinterpret_cast<const float4*>(p + i4);
    const float ga[4] = {gg.x * grad_scale, gg.y * grad_scale, gg.z * grad_scale, gg.w * grad_scale};
    float ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w}, pa[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = b1 * ma[k] + (1.f - b1) * ga[k];
      va[k] = b2 * va[k] + (1.f - b2) * ga[k] * ga[k];
      // torch.optim.Adam: p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
      pa[k] -= step * ma[k] / (sqrtf(va[k]) * rs2 + eps);
    }
    *reinterpret_cast<float4*>(m + i4) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + i4) = make_float4(va[0], va[1], va[2], va[3]);
    *reinterpret_cast<float4*>(p + i4) = make_float4(pa[0], pa[1], pa[2], pa[3]);
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pa[0], pa[1]), hi = __floats2bfloat162_rn(pa[2], pa[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p16 + i4) = pk;
    }
  } else {
    for (int64_t i = i4; i < n; ++i) {
      const float gi = g[i] * grad_scale;
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      const float pn = p[i] - step * mi / (sqrtf(vi) * rs2 + eps);
      p[i] = pn;
      if (p16) p16[i] = __float2bfloat16_rn(pn);
    }
  }
}

// y[i] = a[i] + b[i]
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

// align_corners linear resample over time (models/lib/wav2vec.py:67-73), fp32 output, no LayerNorm (the training path keeps the
// pre-norm activations for the LayerNorm backward). in [B, T_in, C] (in_dtype) -> out fp32 [B*T_out, C]
__global__ void lerp_kernel(const void* __restrict__ in, int in_dtype, int64_t in_batch_stride, float* __restrict__ out, int B, int T_in,
                            int T_out, int C) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * T_out * C) return;
  const int c = (int)(idx % C);
  const int t = (int)((idx / C) % T_out);
  const int b = (int)(idx / ((int64_t)C * T_out));
  const float scale = T_out > 1 ? (float)(T_in - 1) / (float)(T_out - 1) : 0.f;
  const float src = scale * (float)t;
  const int i0 = (int)src;
  const int i1 = i0 + (i0 < T_in - 1 ? 1 : 0);
  const float l1 = src - (float)i0, l0 = 1.f - l1;
  const int64_t base = (int64_t)b * in_batch_stride;
  out[idx] = l0 * load_as_float(in, in_dtype, base + (int64_t)i0 * C + c) + l1 * load_as_float(in, in_dtype, base + (int64_t)i1 * C + c);
}

// CLIPTextEmbeddings: out[b*T + t, :] = tok_emb[ids[b, t], :] + pos_emb[t, :]
__global__ void embed_tokens_kernel(const int64_t* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                                    float* __restrict__ out, int T, int C, int vocab, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c4 = (int)(i % (C / 4));
  const int64_t row = i / (C / 4);
  const int t = (int)(row % T);
  int64_t id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4 a = reinterpret_cast<const float4*>(tok + id * C)[c4];
  const float4 b = reinterpret_cast<const float4*>(pos + (int64_t)t * C)[c4];
  reinterpret_cast<float4*>(out + row * C)[c4] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// out[b, c] = mean_t x[b*T + t, c]
__global__ void token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int T, int C) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += x[((int64_t)b * T + t) * C + c];
  out[(int64_t)b * C + c] = s / (float)T;
}

}  // namespace avi

using namespace avi;

extern "C" int avi_embed_tokens(const int64_t* ids, const float* tok_emb, const float* pos_emb, float* out, int32_t B, int32_t T,
                                int32_t C, int32_t vocab, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && C > 0 && C % 4 == 0 && vocab > 0, "avi_embed_tokens: bad shape");
  const int64_t n4 = (int64_t)B * T * (C / 4);
  embed_tokens_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ids, tok_emb, pos_emb, out, T, C, vocab, n4);
  return check_launch("embed_tokens");
}

extern "C" int avi_token_mean(const float* x, float* out, int32_t B, int32_t T, int32_t C, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && C > 0 && B <= 65535, "avi_token_mean: bad shape");
  token_mean_kernel<<<dim3((C + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(x, out, T, C);
  return check_launch("token_mean");
}

extern "C" int avi_transpose_cast(const float* src, void* dst, int32_t dst_dtype, int32_t R, int32_t C, int64_t src_ld, int32_t R_pad,
                                  void* stream) {
  AVI_REQUIRE(R > 0 && C > 0 && R_pad >= R && src_ld >= C, "avi_transpose_cast: bad shape");
  AVI_REQUIRE((R_pad + 31) / 32 <= 65535, "avi_transpose_cast: too many rows");
  dim3 grid((C + 31) / 32, (R_pad + 31) / 32), block(32, 8);
  if (dst_dtype == AVI_DT_BF16)
    transpose_cast_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), R, C, src_ld, R_pad);
  else
    transpose_cast_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<float*>(dst), R, C, src_ld, R_pad);
  return check_launch("transpose_cast");
}

extern "C" int avi_cast_pad2d(const float* src, void* dst, int32_t dst_dtype, int32_t R, int32_t C, int64_t src_ld, int32_t R_pad,
                              int32_t C_pad, void* stream) {
  AVI_REQUIRE(R > 0 && C > 0 && R_pad >= R && C_pad >= C && src_ld >= C, "avi_cast_pad2d: bad shape");
  const int64_t n = (int64_t)R_pad * C_pad;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dst_dtype == AVI_DT_BF16)
    cast_pad2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), R, C, src_ld, R_pad, C_pad);
  else
    cast_pad2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<float*>(dst), R, C, src_ld, R_pad, C_pad);
  return check_launch("cast_pad2d");
}

extern "C" int avi_tf_input_rows(const float* gt, int64_t gt_ld, const float* tpl, float* out, int32_t B, int32_t T, int32_t C,
                                 int32_t C_pad, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && C > 0 && C_pad >= C && gt_ld >= C, "avi_tf_input_rows: bad shape");
  const int64_t n = (int64_t)B * T * C_pad;
  tf_input_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gt, gt_ld, tpl, out, B, T, C, C_pad);
  return check_launch("tf_input_rows");
}

extern "C" int avi_ff_add_style_pe(float* x, const float* style, int64_t style_stride, const float* pe, int32_t B, int32_t T, int32_t fd,
                                   int32_t period, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && fd > 0 && period > 0, "avi_ff_add_style_pe: bad shape");
  const int64_t n = (int64_t)B * T * fd;
  add_style_pe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, style, style_stride, pe, B, T, fd, period);
  return check_launch("add_style_pe");
}

extern "C" int avi_colsum(const float* x, float* out, int32_t R, int32_t N, int64_t ld, int32_t accumulate, void* stream) {
  AVI_REQUIRE(R > 0 && N > 0 && ld >= N, "avi_colsum: bad shape");
  int splits = (R + 127) / 128;
  if (splits > 64) splits = 64;
  const int rows_per_block = (R + splits - 1) / splits;
  if (splits > 1 && !accumulate && cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, (cudaStream_t)stream) != cudaSuccess) {
    set_error("avi_colsum: memset failed");
    return 1;
  }
  colsum_kernel<<<dim3((N + 31) / 32, splits), dim3(32, 8), 0, (cudaStream_t)stream>>>(x, out, R, N, ld, rows_per_block, splits > 1 ? 1 : 0,
                                                                                       accumulate);
  return check_launch("colsum");
}

extern "C" int avi_act_fwd(const float* pre, float* out_f32, void* out_bf16, int64_t n, int32_t act, void* stream) {
  AVI_REQUIRE(n > 0 && (act == AVI_ACT_GELU || act == AVI_ACT_RELU || act == AVI_ACT_SILU), "avi_act_fwd: bad arguments");
  act_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pre, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), n, act);
  return check_launch("act_fwd");
}

extern "C" int avi_act_bwd(const float* pre, const float* dout, float* dpre, int64_t n, int32_t act, void* stream) {
  AVI_REQUIRE(n > 0 && (act == AVI_ACT_GELU || act == AVI_ACT_RELU || act == AVI_ACT_SILU), "avi_act_bwd: bad arguments");
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pre, dout, dpre, n, act);
  return check_launch("act_bwd");
}

extern "C" int avi_layernorm_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int64_t rows, int32_t C,
                                 float eps, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0 && C <= 8192, "avi_layernorm_bwd: bad shape");
  if (C > 1024) {
    layernorm_bwd_wide_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, w, dy, dx, dw, db, C, eps);
    return check_launch("layernorm_bwd_wide");
  }
  layernorm_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, w, dy, dx, dw, db, rows, C, eps);
  return check_launch("layernorm_bwd");
}

static size_t att_smem_fwd(int T, int D) { return ((size_t)2 * T * (D + ATT_PAD) + 8 * 2 * D + 8 * 2 * ATT_MAXT) * sizeof(float); }
static size_t att_smem_kv(int T, int D) { return ((size_t)2 * T * (D + ATT_PAD) + 2 * ATT_ROWS * ATT_MAXT) * sizeof(float); }

template <int HD>
static int attn_fwd_launch(const float* qkv, float* out, float* P, int B, int T, int H, float scale, int bias_mode, int period,
                           const float* pmask, cudaStream_t st) {
  static SmemOptIn optin;
  const cudaError_t attr_err = smem_optin(attn_train_fwd_kernel<HD>, (int)att_smem_fwd(ATT_MAXT, HD), optin);
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_attn_train_fwd: cudaFuncSetAttribute failed");
  attn_train_fwd_kernel<HD><<<dim3(H, B, (T + ATT_ROWS - 1) / ATT_ROWS), 256, att_smem_fwd(T, HD), st>>>(qkv, out, P, T, H, scale, bias_mode,
                                                                                                        period > 0 ? period : 1, pmask);
  return check_launch("attn_train_fwd");
}

template <int HD>
static int attn_bwd_launch(const float* qkv, const float* P, const float* dout, float* dqkv, float* dS, int B, int T, int H, float scale,
                           const float* pmask, cudaStream_t st) {
  static SmemOptIn optin_q, optin_kv;
  const cudaError_t e1 = smem_optin(attn_train_bwd_q_kernel<HD>, (int)att_smem_fwd(ATT_MAXT, HD), optin_q);
  const cudaError_t e2 = smem_optin(attn_train_bwd_kv_kernel<HD>, (int)att_smem_kv(ATT_MAXT, HD), optin_kv);
  AVI_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, "avi_attn_train_bwd: cudaFuncSetAttribute failed");
  const dim3 grid(H, B, (T + ATT_ROWS - 1) / ATT_ROWS);
  attn_train_bwd_q_kernel<HD><<<grid, 256, att_smem_fwd(T, HD), st>>>(qkv, P, dout, dqkv, dS, T, H, scale, pmask);
  if (check_launch("attn_train_bwd_q")) return 1;
  attn_train_bwd_kv_kernel<HD><<<grid, 256, att_smem_kv(T, HD), st>>>(qkv, P, dout, dqkv, dS, T, H, scale, pmask);
  return check_launch("attn_train_bwd_kv");
}

extern "C" int avi_attn_train_fwd_drop(const float* qkv, float* out, float* P, const float* pmask, int32_t B, int32_t T, int32_t H,
                                       int32_t D, float scale, int32_t bias_mode, int32_t period, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && T <= ATT_MAXT && H > 0 && (D == 16 || D == 32 || D == 64),
              "avi_attn_train_fwd: T <= 128 and head dim 16 / 32 / 64 (T=%d D=%d)", T, D);
  AVI_REQUIRE(bias_mode >= 0 && bias_mode <= 2, "avi_attn_train_fwd: bias_mode 0 (none), 1 (FaceFormer biased causal), 2 (causal)");
  AVI_REQUIRE(bias_mode != 1 || H == 4, "avi_attn_train_fwd: the FaceFormer bias mask is defined for 4 heads");
  AVI_REQUIRE(((uintptr_t)qkv % 16) == 0, "avi_attn_train_fwd: qkv must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 64) return attn_fwd_launch<64>(qkv, out, P, B, T, H, scale, bias_mode, period, pmask, st);
  if (D == 32) return attn_fwd_launch<32>(qkv, out, P, B, T, H, scale, bias_mode, period, pmask, st);
  return attn_fwd_launch<16>(qkv, out, P, B, T, H, scale, bias_mode, period, pmask, st);
}

extern "C" int avi_attn_train_fwd(const float* qkv, float* out, float* P, int32_t B, int32_t T, int32_t H, int32_t D, float scale,
                                  int32_t bias_mode, int32_t period, void* stream) {
  return avi_attn_train_fwd_drop(qkv, out, P, nullptr, B, T, H, D, scale, bias_mode, period, stream);
}

extern "C" int avi_attn_train_bwd_drop(const float* qkv, const float* P, const float* pmask, const float* dout, float* dqkv,
                                       float* dS_scratch, int32_t B, int32_t T, int32_t H, int32_t D, float scale, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && T <= ATT_MAXT && H > 0 && (D == 16 || D == 32 || D == 64),
              "avi_attn_train_bwd: T <= 128 and head dim 16 / 32 / 64 (T=%d D=%d)", T, D);
  AVI_REQUIRE((((uintptr_t)qkv | (uintptr_t)dout) % 16) == 0, "avi_attn_train_bwd: qkv / dout must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 64) return attn_bwd_launch<64>(qkv, P, dout, dqkv, dS_scratch, B, T, H, scale, pmask, st);
  if (D == 32) return attn_bwd_launch<32>(qkv, P, dout, dqkv, dS_scratch, B, T, H, scale, pmask, st);
  return attn_bwd_launch<16>(qkv, P, dout, dqkv, dS_scratch, B, T, H, scale, pmask, st);
}

extern "C" int avi_attn_train_bwd(const float* qkv, const float* P, const float* dout, float* dqkv, float* dS_scratch, int32_t B, int32_t T,
                                  int32_t H, int32_t D, float scale, void* stream) {
  return avi_attn_train_bwd_drop(qkv, P, nullptr, dout, dqkv, dS_scratch, B, T, H, D, scale, stream);
}

// ------------------------------------------------------------------------------------------------ train-mode regularisers
// nn.Dropout with the draw as an input: y = x * mask (+ residual); mask is Bernoulli(1-p) / (1-p), i.e. already scaled. The same kernel is
// its own backward (dx = dy * mask). float4 over n (n % 4 == 0 is required by the caller-side shapes: 64 / 768 / 3072 wide rows).
__global__ void __launch_bounds__(256) mask_mul_add_kernel(const float4* __restrict__ x, const float4* __restrict__ mask,
                                                           const float4* __restrict__ res, float4* __restrict__ y, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = x[i], k = mask[i];
    float4 o = make_float4(a.x * k.x, a.y * k.y, a.z * k.z, a.w * k.w);
    if (res != nullptr) {
      const float4 r = res[i];
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    y[i] = o;
  }
}

extern "C" int avi_mask_mul_add(const float* x, const float* mask, const float* residual, float* y, int64_t n, void* stream) {
  AVI_REQUIRE(n > 0 && n % 4 == 0, "avi_mask_mul_add: n must be a positive multiple of 4 (n=%lld)", (long long)n);
  AVI_REQUIRE((((uintptr_t)x | (uintptr_t)mask | (uintptr_t)residual | (uintptr_t)y) % 16) == 0, "avi_mask_mul_add: 16-byte aligned buffers");
  const int64_t n4 = n / 4;
  const unsigned grid = (unsigned)std::min<int64_t>((n4 + 255) / 256, 148 * 8);
  mask_mul_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)mask, (const float4*)residual, (float4*)y, n4);
  return check_launch("mask_mul_add");
}

// SpecAugment along time (models/lib/wav2vec.py:120-131): rows whose mask byte is set are REPLACED by masked_spec_embed (in place).
__global__ void __launch_bounds__(256) spec_augment_fwd_kernel(float* __restrict__ x, const uint8_t* __restrict__ row_mask,
                                                               const float* __restrict__ embed, int C) {
  const int64_t r = blockIdx.x;
  if (!row_mask[r]) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) x[r * C + c] = embed[c];
}

// backward: g_embed[c] = sum over the masked rows of dx[r][c] (fixed order: deterministic), then those rows of dx become zero.
__global__ void __launch_bounds__(256) spec_augment_bwd_kernel(float* __restrict__ dx, const uint8_t* __restrict__ row_mask,
                                                               float* __restrict__ g_embed, int64_t rows, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int64_t r = 0; r < rows; ++r) {
    if (row_mask[r]) {          // warp-uniform
      acc += dx[r * C + c];
      dx[r * C + c] = 0.f;
    }
  }
  g_embed[c] = acc;
}

extern "C" int avi_spec_augment_fwd(float* x, const uint8_t* row_mask, const float* embed, int64_t rows, int32_t C, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0, "avi_spec_augment_fwd: bad shape");
  spec_augment_fwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, row_mask, embed, C);
  return check_launch("spec_augment_fwd");
}

extern "C" int avi_spec_augment_bwd(float* dx, const uint8_t* row_mask, float* g_embed, int64_t rows, int32_t C, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0, "avi_spec_augment_bwd: bad shape");
  spec_augment_bwd_kernel<<<(unsigned)((C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dx, row_mask, g_embed, rows, C);
  return check_launch("spec_augment_bwd");
}

extern "C" int avi_posconv_dw(const float* x, const float* dpc, float* dw, int32_t B, int32_t T, int32_t C, int32_t groups, int32_t k,
                              void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && C > 0 && groups > 0 && C % groups == 0 && k > 0, "avi_posconv_dw: bad shape");
  posconv_dw_kernel<<<dim3(k, groups), 256, 0, (cudaStream_t)stream>>>(x, dpc, dw, B, T, C, C / groups, k);
  return check_launch("posconv_dw");
}

extern "C" int avi_posconv_unfold_t(const float* x, void* out, int32_t out_dtype, int32_t B, int32_t T, int32_t Tq, int32_t C, int32_t c0,
                                    int32_t CW, int32_t k, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && Tq >= T && C > 0 && c0 >= 0 && CW > 0 && c0 + CW <= C && k > 0 && (int64_t)B * k <= 65535,
              "avi_posconv_unfold_t: bad shape");
  const dim3 grid((CW + 31) / 32, (Tq + 31) / 32, B * k);
  if (out_dtype == AVI_DT_BF16)
    posconv_unfold_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), B, T, Tq, C, c0, CW, k);
  else
    posconv_unfold_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<float*>(out), B, T, Tq, C, c0, CW, k);
  return check_launch("posconv_unfold_t");
}

extern "C" int avi_weightnorm_bwd(const float* v, const float* g, const float* dw, float* dv, float* dg, int32_t n_rows, int32_t k,
                                  void* stream) {
  AVI_REQUIRE(n_rows > 0 && k > 0, "avi_weightnorm_bwd: bad shape");
  weightnorm_bwd_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(v, g, dw, dv, dg, n_rows, k);
  return check_launch("weightnorm_bwd");
}

extern "C" int avi_mse_loss_grad(const float* out, const float* gt, float* dout, double* loss, int64_t rows, int32_t C, int64_t out_ld,
                                 int64_t gt_ld, float loss_scale, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0 && out_ld >= C && gt_ld >= C, "avi_mse_loss_grad: bad shape");
  if (cudaMemsetAsync(loss, 0, sizeof(double), (cudaStream_t)stream) != cudaSuccess) {
    set_error("avi_mse_loss_grad: memset failed");
    return 1;
  }
  mse_loss_grad_kernel<<<device_sms() * 4, 256, 0, (cudaStream_t)stream>>>(out, gt, dout, loss, rows, C, out_ld, gt_ld, loss_scale);
  return check_launch("mse_loss_grad");
}

extern "C" int avi_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                             float eps, int32_t step, float grad_scale, void* stream) {
  AVI_REQUIRE(n > 0 && step >= 1, "avi_adam_step: bad arguments");
  AVI_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0 && ((uintptr_t)p_bf16 % 8) == 0,
              "avi_adam_step: buffers must be 16-byte aligned");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  const int64_t n4 = (n + 3) / 4;
  adam_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, lr, beta1,
                                                                            beta2, eps, bc1, bc2, grad_scale);
  return check_launch("adam");
}

extern "C" int avi_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream) {
  AVI_REQUIRE(n > 0, "avi_add_f32: bad size");
  add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, y, n);
  return check_launch("add");
}

extern "C" int avi_w2v_lerp(const void* in, int32_t in_dtype, int64_t in_batch_stride, float* out, int32_t B, int32_t T_in, int32_t T_out,
                            int32_t C, void* stream) {
  AVI_REQUIRE(B > 0 && T_in > 0 && T_out > 0 && C > 0, "avi_w2v_lerp: bad shape");
  const int64_t n = (int64_t)B * T_out * C;
  lerp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, in_dtype, in_batch_stride, out, B, T_in, T_out, C);
  return check_launch("lerp");
}
