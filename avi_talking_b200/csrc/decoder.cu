// FaceFormer decoder kernels (Path A), models/faceformer_disentangle.py:
//   * avi_ff_decoder_ar  - the autoregressive branch of forward_ff (:461-476) as ONE persistent kernel per clip:
//     KV-cached (causal => O(T) instead of the reference's O(T^2) prefix recomputation), temporal bias
//     -slope_h*floor((i-j)/period) (init_biased_mask :56-77) built on the fly, degenerate alignment-masked
//     cross-attention (enc_dec_mask :80-88 leaves one key => out_proj(v_proj(mem_t))) precomputed by GEMMs,
//     feedback through the composed vertice_map o vertice_map_r map.
//   * avi_ff_biased_attn - biased causal self-attention over a whole sequence (teacher-forced branch :445-458).
// nn.TransformerDecoderLayer semantics: post-LN, ReLU FFN, eps 1e-5 (torch defaults, :195-196).
#include "common.cuh"

namespace avi {

constexpr int DEC_THREADS = 256;
constexpr int DEC_NH = 4;

// y[o] = b[o] + sum_k Wt[k*OUT + o] * x[k]   (Wt is the nn.Linear weight transposed: [IN, OUT], coalesced over o)
// All DEC_THREADS threads must call. x, y in shared memory; red = scratch of DEC_THREADS floats.
__device__ __forceinline__ void matvec(const float* __restrict__ Wt, const float* __restrict__ b, const float* x, float* y,
                                       float* red, int IN, int OUT, bool relu) {
  const int tid = threadIdx.x;
  if (OUT * 2 <= DEC_THREADS) {
    const int S = DEC_THREADS / OUT;  // k-split
    const int g = tid / OUT, o = tid % OUT;
    float acc = 0.f;
    if (g < S) {
      const int k0 = (IN * g) / S, k1 = (IN * (g + 1)) / S;
#pragma unroll 4
      for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(Wt + (int64_t)k * OUT + o), x[k], acc);
    }
    red[tid] = acc;
    __syncthreads();
    if (tid < OUT) {
      float s = b ? b[tid] : 0.f;
      for (int q = 0; q < S; ++q) s += red[q * OUT + tid];
      y[tid] = relu ? fmaxf(s, 0.f) : s;
    }
  } else {
    for (int o = tid; o < OUT; o += DEC_THREADS) {
      float acc = b ? b[o] : 0.f;
#pragma unroll 4
      for (int k = 0; k < IN; ++k) acc = fmaf(__ldg(Wt + (int64_t)k * OUT + o), x[k], acc);
      y[o] = relu ? fmaxf(acc, 0.f) : acc;
    }
  }
  __syncthreads();
}

// y = LayerNorm(a + r) * w + b over fd elements (fd <= 256), computed by warp 0; all threads must call.
__device__ __forceinline__ void add_layernorm(const float* a, const float* r, const float* __restrict__ w,
                                              const float* __restrict__ b, float* y, int fd) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = lane + 32 * u;
      v[u] = (c < fd) ? a[c] + r[c] : 0.f;
      s += v[u];
    }
    const float mean = warp_sum(s) / fd;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = lane + 32 * u;
      if (c < fd) q += (v[u] - mean) * (v[u] - mean);
    }
    const float rstd = rsqrtf(warp_sum(q) / fd + 1e-5f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = lane + 32 * u;
      if (c < fd) y[c] = (v[u] - mean) * rstd * w[c] + b[c];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float head_slope(int h) { return exp2f(-2.0f * (float)(h + 1)); }  // 4 heads: 2^-2,2^-4,2^-6,2^-8

__global__ void __launch_bounds__(DEC_THREADS, 1)
ff_decoder_ar_kernel(const AviDecoderWeights w, const float* __restrict__ cross, const float* __restrict__ style,
                     float* __restrict__ hidden_out, float* __restrict__ kv_scratch, int T, int fd, int period, int kv_in_smem) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int dff = 2 * fd, hd = fd / DEC_NH;
  const int kvs = fd + 1;  // padded row stride (bank-conflict-free column walks)
  float* xs = sm;               // [fd]    layer input
  float* qkv = xs + fd;         // [3fd]
  float* att = qkv + 3 * fd;    // [fd]    attention output (pre out_proj)
  float* t0 = att + fd;         // [fd]
  float* x1 = t0 + fd;          // [fd]
  float* hbuf = x1 + fd;        // [dff]
  float* emb = hbuf + dff;      // [fd]    next input embedding
  float* sty = emb + fd;        // [fd]
  float* red = sty + fd;        // [DEC_THREADS]
  float* prob = red + DEC_THREADS;         // [NH][T]
  float* Kc = prob + DEC_NH * T;            // [T][kvs] (smem) or global
  float* Vc = Kc + (size_t)T * kvs;
  if (!kv_in_smem) {
    Kc = kv_scratch + (size_t)b * 2 * T * kvs;
    Vc = Kc + (size_t)T * kvs;
  }
  const float inv_sqrt_hd = rsqrtf((float)hd);
  for (int c = tid; c < fd; c += DEC_THREADS) {
    sty[c] = style[(int64_t)b * fd + c];
    emb[c] = sty[c];
  }
  __syncthreads();

  for (int i = 0; i < T; ++i) {
    // ---- PPE: x = emb + pe[i mod period]  (:466-468; pe repeats with period, PeriodicPositionalEncoding :92-107)
    for (int c = tid; c < fd; c += DEC_THREADS) xs[c] = emb[c] + w.pe[(i % period) * fd + c];
    __syncthreads();
    // ---- self-attention in_proj -> q | k | v
    matvec(w.sa_in_w, w.sa_in_b, xs, qkv, red, fd, 3 * fd, false);
    for (int c = tid; c < fd; c += DEC_THREADS) {
      Kc[(size_t)i * kvs + c] = qkv[fd + c];
      Vc[(size_t)i * kvs + c] = qkv[2 * fd + c];
    }
    __syncthreads();
    // ---- scores: head h handled by 64 threads, keys strided by 64
    {
      const int h = tid >> 6, jl = tid & 63;
      const float slope = head_slope(h);
      const float* q = qkv + h * hd;
      float mx = -INFINITY;
      for (int j = jl; j <= i; j += 64) {
        const float* kr = Kc + (size_t)j * kvs + h * hd;
        float s = 0.f;
        for (int d = 0; d < hd; ++d) s = fmaf(q[d], kr[d], s);
        s = s * inv_sqrt_hd - slope * (float)((i - j) / period);
        prob[h * T + j] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      if ((tid & 31) == 0) red[tid >> 5] = mx;
      __syncthreads();
      mx = fmaxf(red[(tid >> 6) * 2], red[(tid >> 6) * 2 + 1]);
      float sum = 0.f;
      for (int j = jl; j <= i; j += 64) {
        const float e = expf(prob[h * T + j] - mx);
        prob[h * T + j] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      __syncthreads();
      if ((tid & 31) == 0) red[8 + (tid >> 5)] = sum;
      __syncthreads();
      const float inv = 1.f / (red[8 + (tid >> 6) * 2] + red[8 + (tid >> 6) * 2 + 1]);
      for (int j = jl; j <= i; j += 64) prob[h * T + j] *= inv;
      __syncthreads();
    }
    // ---- att[c] = sum_j p[h(c)][j] * V[j][c] : DEC_THREADS/fd key groups, then reduce
    {
      const int G = DEC_THREADS / fd >= 1 ? DEC_THREADS / fd : 1;
      if (fd <= DEC_THREADS) {
        const int g = tid / fd, c = tid % fd;
        float acc = 0.f;
        if (g < G) {
          const float* pr = prob + (c / hd) * T;
          for (int j = g; j <= i; j += G) acc = fmaf(pr[j], Vc[(size_t)j * kvs + c], acc);
        }
        red[tid] = acc;
        __syncthreads();
        if (tid < fd) {
          float s = 0.f;
          for (int q = 0; q < G; ++q) s += red[q * fd + tid];
          att[tid] = s;
        }
        __syncthreads();
      }
    }
    // ---- out_proj, residual, LN1
    matvec(w.sa_out_w, w.sa_out_b, att, t0, red, fd, fd, false);
    add_layernorm(xs, t0, w.ln1_w, w.ln1_b, x1, fd);
    // ---- cross-attention term (precomputed), LN2
    for (int c = tid; c < fd; c += DEC_THREADS) t0[c] = cross[((int64_t)b * T + i) * fd + c];
    __syncthreads();
    add_layernorm(x1, t0, w.ln2_w, w.ln2_b, xs, fd);
    // ---- FFN, LN3
    matvec(w.ff1_w, w.ff1_b, xs, hbuf, red, fd, dff, true);
    matvec(w.ff2_w, w.ff2_b, hbuf, t0, red, dff, fd, false);
    add_layernorm(xs, t0, w.ln3_w, w.ln3_b, x1, fd);
    for (int c = tid; c < fd; c += DEC_THREADS) hidden_out[((int64_t)b * T + i) * fd + c] = x1[c];
    // ---- feedback: emb_{i+1} = vertice_map(vertice_map_r(y_i)) + style   (:473-476), composed map
    if (i + 1 < T) {
      matvec(w.fb_w, w.fb_b, x1, t0, red, fd, fd, false);
      for (int c = tid; c < fd; c += DEC_THREADS) emb[c] = t0[c] + sty[c];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// feature_dim == 64 fast path: 512 threads, ALL decoder weights (36.9k floats) live in registers (80 per thread), so a
// step is a handful of register-resident mat-vecs with shuffle reductions instead of L2-latency-bound weight streaming.
// K/V cache, probabilities and the activation vectors live in shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int A64_THREADS = 512, A64_FD = 64, A64_DFF = 128, A64_HD = 16, A64_KVS = 65;

__device__ __forceinline__ float dot_seg(const float* w, const float* x, int n) {  // n multiple of 4, x 16-byte aligned smem
  float acc = 0.f;
  for (int q = 0; q < n; q += 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + q);
    acc = fmaf(w[q], xv.x, acc);
    acc = fmaf(w[q + 1], xv.y, acc);
    acc = fmaf(w[q + 2], xv.z, acc);
    acc = fmaf(w[q + 3], xv.w, acc);
  }
  return acc;
}

__device__ __forceinline__ void ln64(const float* a, const float* r, const float* __restrict__ w, const float* __restrict__ b, float* y) {
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    const float v0 = a[l] + r[l], v1 = a[l + 32] + r[l + 32];
    const float mean = warp_sum(v0 + v1) * (1.f / 64.f);
    const float d0 = v0 - mean, d1 = v1 - mean;
    const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-5f);
    y[l] = d0 * rstd * w[l] + b[l];
    y[l + 32] = d1 * rstd * w[l + 32] + b[l + 32];
  }
  __syncthreads();
}

// y = LN2( LN1(a + r) + c ): the two post-LN residual steps around the (degenerate) cross-attention in one warp-level phase, the
// cross row c coming straight from global memory (its loads are issued before the first reduction)
__device__ __forceinline__ void ln64x2(const float* a, const float* r, const float* __restrict__ c, const float* __restrict__ w1,
                                       const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, float* y) {
  if (threadIdx.x < 32) {
    const int l = threadIdx.x;
    const float c0 = c[l], c1 = c[l + 32];
    float v0 = a[l] + r[l], v1 = a[l + 32] + r[l + 32];
    float mean = warp_sum(v0 + v1) * (1.f / 64.f);
    float d0 = v0 - mean, d1 = v1 - mean;
    float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-5f);
    v0 = d0 * rstd * w1[l] + b1[l] + c0;
    v1 = d1 * rstd * w1[l + 32] + b1[l + 32] + c1;
    mean = warp_sum(v0 + v1) * (1.f / 64.f);
    d0 = v0 - mean;
    d1 = v1 - mean;
    rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-5f);
    y[l] = d0 * rstd * w2[l] + b2[l];
    y[l + 32] = d1 * rstd * w2[l + 32] + b2[l + 32];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(A64_THREADS, 1)
ff_decoder_ar64_kernel(const AviDecoderWeights w, const float* __restrict__ cross, const float* __restrict__ style,
                       float* __restrict__ hidden_out, int T, int period) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* xs = sm;                 // [64]
  float* qkv = xs + 64;           // [192]
  float* att = qkv + 192;         // [64]
  float* t0 = att + 64;           // [64]
  float* x1 = t0 + 64;            // [64]
  float* hb = x1 + 64;            // [128]
  float* emb = hb + 128;          // [64]
  float* sty = emb + 64;          // [64]
  float* red = sty + 64;          // [32 + 512] warp partials + PV partials
  float* prob = red + 576;        // [4][T]
  float* Kc = prob + 4 * T + ((4 - (4 * T) % 4) % 4);  // [T][65]
  float* Vc = Kc + (size_t)T * A64_KVS;

  // ---- weights -> registers (transposed global layout Wt[k*OUT + o])
  float w_in[32], w_o[8], w_1[16], w_2[16], w_f[8];
  const int in_o = tid >> 1, in_s = tid & 1;
#pragma unroll
  for (int kk = 0; kk < 32; ++kk) w_in[kk] = (in_o < 192) ? w.sa_in_w[(in_s * 32 + kk) * 192 + in_o] : 0.f;
  const int o8 = tid >> 3, s8 = tid & 7;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    w_o[kk] = w.sa_out_w[(s8 * 8 + kk) * 64 + o8];
    w_f[kk] = w.fb_w[(s8 * 8 + kk) * 64 + o8];
  }
  const int o4 = tid >> 2, s4 = tid & 3;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    w_1[kk] = w.ff1_w[(s4 * 16 + kk) * 128 + o4];
    w_2[kk] = w.ff2_w[(s8 * 16 + kk) * 64 + o8];
  }
  const float b_in = (in_o < 192) ? w.sa_in_b[in_o] : 0.f;
  const float b_o = w.sa_out_b[o8], b_f = w.fb_b[o8], b_1 = w.ff1_b[o4], b_2 = w.ff2_b[o8];

  if (tid < 64) {
    sty[tid] = style[(int64_t)b * 64 + tid];
    xs[tid] = sty[tid] + w.pe[tid];                     // x_0 = style + pe[0]
  }
  __syncthreads();
  const int h = tid >> 7, jl = tid & 127;         // scores: 4 heads x 128 key lanes
  const float slope = head_slope(h);
  const int pg = tid >> 6, pc = tid & 63;         // PV: 8 key groups x 64 channels

  for (int i = 0; i < T; ++i) {
    {  // in_proj (xs = emb_i + pe[i mod period] was written by the feedback phase of the previous frame)
      float acc = dot_seg(w_in, xs + in_s * 32, 32);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (in_s == 0 && in_o < 192) {
        acc += b_in;
        qkv[in_o] = acc;
        if (in_o >= 128) Vc[(size_t)i * A64_KVS + in_o - 128] = acc;
        else if (in_o >= 64) Kc[(size_t)i * A64_KVS + in_o - 64] = acc;
      }
    }
    __syncthreads();
    // scores + softmax statistics
    float q[A64_HD];
#pragma unroll
    for (int d = 0; d < A64_HD; d += 4) {
      const float4 qv = *reinterpret_cast<const float4*>(qkv + h * A64_HD + d);
      q[d] = qv.x; q[d + 1] = qv.y; q[d + 2] = qv.z; q[d + 3] = qv.w;
    }
    float sc[2];  // T <= 256 -> at most two keys per thread (guarded on the host)
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = jl + 128 * u;
      sc[u] = -INFINITY;
      if (j <= i) {
        const float* kr = Kc + (size_t)j * A64_KVS + h * A64_HD;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < A64_HD; ++d) s = fmaf(q[d], kr[d], s);
        sc[u] = s * 0.25f - slope * (float)((i - j) / period);
        mx = fmaxf(mx, sc[u]);
      }
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[h * 4], red[h * 4 + 1]), fmaxf(red[h * 4 + 2], red[h * 4 + 3]));
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = jl + 128 * u;
      if (j <= i) {
        const float e = expf(sc[u] - mx);
        prob[h * T + j] = e;
        sum += e;
      }
    }
    sum = warp_sum(sum);
    if (lane == 0) red[16 + warp] = sum;
    __syncthreads();
    {  // PV (unnormalised), 8 key groups
      const float* pr = prob + (pc >> 4) * T;
      float acc = 0.f;
      for (int j = pg; j <= i; j += 8) acc = fmaf(pr[j], Vc[(size_t)j * A64_KVS + pc], acc);
      red[32 + tid] = acc;
    }
    __syncthreads();
    if (tid < 64) {
      const int hh = tid >> 4;
      const float inv = 1.f / (red[16 + hh * 4] + red[16 + hh * 4 + 1] + red[16 + hh * 4 + 2] + red[16 + hh * 4 + 3]);
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) s += red[32 + g * 64 + tid];
      att[tid] = s * inv;
    }
    __syncthreads();
    {  // out_proj
      float acc = dot_seg(w_o, att + s8 * 8, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (s8 == 0) t0[o8] = acc + b_o;
    }
    __syncthreads();
    ln64x2(xs, t0, cross + ((int64_t)b * T + i) * 64, w.ln1_w, w.ln1_b, w.ln2_w, w.ln2_b, xs);
    {  // ff1 + ReLU
      float acc = dot_seg(w_1, xs + s4 * 16, 16);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (s4 == 0) hb[o4] = fmaxf(acc + b_1, 0.f);
    }
    __syncthreads();
    {  // ff2
      float acc = dot_seg(w_2, hb + s8 * 16, 16);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (s8 == 0) t0[o8] = acc + b_2;
    }
    __syncthreads();
    ln64(xs, t0, w.ln3_w, w.ln3_b, x1);
    if (tid < 64) hidden_out[((int64_t)b * T + i) * 64 + tid] = x1[tid];
    {  // feedback map: emb_{i+1} = fb(y_i) + style
      float acc = dot_seg(w_f, x1 + s8 * 8, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      // x_{i+1} = emb_{i+1} + pe[(i+1) mod period]: the LN3 phase above was the last reader of xs (its residual input)
      if (s8 == 0) xs[o8] = acc + b_f + sty[o8] + w.pe[((i + 1) % period) * 64 + o8];
    }
    __syncthreads();
  }
}

// Biased causal self-attention for a whole sequence (teacher forcing): one warp per (clip, head, query row).
// qkv [B,T,3fd] fp32 -> out [B,T,fd].
__global__ void __launch_bounds__(128) ff_biased_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int B, int T,
                                                             int fd, int period) {
  extern __shared__ float sm[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * 4 + wib;
  if (gw >= (int64_t)B * DEC_NH * T) return;
  const int i = (int)(gw % T);
  const int h = (int)((gw / T) % DEC_NH);
  const int b = (int)(gw / ((int64_t)T * DEC_NH));
  const int hd = fd / DEC_NH;
  float* p = sm + wib * T;
  const float* base = qkv + (int64_t)b * T * 3 * fd;
  const float* q = base + (int64_t)i * 3 * fd + h * hd;
  const float slope = head_slope(h);
  const float sc = rsqrtf((float)hd);
  float mx = -INFINITY;
  for (int j = lane; j <= i; j += 32) {
    const float* kr = base + (int64_t)j * 3 * fd + fd + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(q[d], kr[d], s);
    s = s * sc - slope * (float)((i - j) / period);
    p[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j <= i; j += 32) {
    const float e = expf(p[j] - mx);
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int d = lane; d < hd; d += 32) {
    float acc = 0.f;
    for (int j = 0; j <= i; ++j) acc = fmaf(p[j], base[(int64_t)j * 3 * fd + 2 * fd + h * hd + d], acc);
    out[((int64_t)b * T + i) * fd + h * hd + d] = acc * inv;
  }
}

// convert_coeff2verts prologue (faceformer_disentangle.py:426-430): exp_out[f, :n_exp] = coeff[f, :n_exp] * std + mean, and the global
// rotation pose[f, :3] is zeroed IN PLACE (as upstream does to the caller's tensor).
__global__ void ff_denorm_coeff_kernel(const float* __restrict__ coeff, const float* __restrict__ mean, const float* __restrict__ stdv,
                                       float* pose, float* __restrict__ exp_out, int F, int n_coeff, int n_exp, int n_pose) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < F * n_exp) {
    const int f = idx / n_exp, c = idx - f * n_exp;
    exp_out[idx] = coeff[(int64_t)f * n_coeff + c] * stdv[c] + mean[c];
  }
  if (idx < F * 3) pose[(int64_t)(idx / 3) * n_pose + idx % 3] = 0.f;
}

// Conditioning columns of the decoder input (faceformer_disentangle.py:808): out[r, 0:6] = eye (one learnable row, or a row per
// frame), out[r, 6:36] = emotion embedding of frame r; the audio columns [36, 36 + fd) are written by the audio_feature_map GEMM.
__global__ void ff_fill_cond_kernel(const float* __restrict__ eye, int eye_per_row, const float* __restrict__ emo, int64_t emo_clip_stride,
                                    int T, float* __restrict__ out, int64_t rows, int ld) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 36) return;
  const int64_t r = idx / 36;
  const int c = (int)(idx - r * 36);
  float v;
  if (c < 6) v = eye[eye_per_row ? r * 6 + c : c];
  else v = emo[(r / T) * emo_clip_stride + (r % T) * 30 + (c - 6)];
  out[r * ld + c] = v;
}

}  // namespace avi

using namespace avi;

extern "C" int avi_ff_denorm_coeff(const float* coeff, const float* mean, const float* stdv, float* pose, float* exp_out, int32_t F,
                                   int32_t n_coeff, int32_t n_exp, int32_t n_pose, void* stream) {
  AVI_REQUIRE(coeff && mean && stdv && pose && exp_out && F > 0 && n_exp > 0 && n_exp <= n_coeff && n_pose >= 3, "avi_ff_denorm_coeff: bad arguments");
  const int n = F * (n_exp > 3 ? n_exp : 3);
  ff_denorm_coeff_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(coeff, mean, stdv, pose, exp_out, F, n_coeff, n_exp, n_pose);
  return check_launch("ff_denorm_coeff");
}

extern "C" int avi_ff_fill_cond(const float* eye, int32_t eye_per_row, const float* emo, int64_t emo_clip_stride, float* out, int32_t B,
                                int32_t T, int32_t ld, void* stream) {
  AVI_REQUIRE(eye && emo && out && B > 0 && T > 0 && ld >= 36, "avi_ff_fill_cond: bad arguments");
  const int64_t rows = (int64_t)B * T;
  ff_fill_cond_kernel<<<(unsigned)((rows * 36 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(eye, eye_per_row, emo, emo_clip_stride, T, out, rows, ld);
  return check_launch("ff_fill_cond");
}

extern "C" int avi_ff_decoder_ar(const AviDecoderWeights* w, const float* cross, const float* style, float* hidden_out,
                                 float* kv_scratch, int32_t B, int32_t T, int32_t fd, int32_t period, void* stream) {
  AVI_REQUIRE(w != nullptr && B > 0 && T > 0 && period > 0, "avi_ff_decoder_ar: bad arguments");
  AVI_REQUIRE(fd % 32 == 0 && fd >= 32 && fd <= 256 && DEC_THREADS % fd == 0, "avi_ff_decoder_ar: feature_dim %d unsupported (32/64/128/256)", fd);
  if (fd == A64_FD && T <= 256) {
    const size_t sm64 = sizeof(float) * (64 + 192 + 64 * 3 + 128 + 64 * 2 + 576 + 4 * (size_t)T + 4 + 2 * (size_t)T * A64_KVS);
    if (sm64 <= 227 * 1024) {
      static SmemOptIn optin64;
      cudaError_t e64 = smem_optin(ff_decoder_ar64_kernel, 227 * 1024, optin64);
      AVI_REQUIRE(e64 == cudaSuccess, "avi_ff_decoder_ar: cudaFuncSetAttribute: %s", cudaGetErrorString(e64));
      ff_decoder_ar64_kernel<<<B, A64_THREADS, sm64, (cudaStream_t)stream>>>(*w, cross, style, hidden_out, T, period);
      return check_launch("ff_decoder_ar64");
    }
  }
  const size_t fixed = sizeof(float) * ((size_t)fd * 9 + 2 * fd + DEC_THREADS + (size_t)DEC_NH * T);
  const size_t kv = sizeof(float) * 2 * (size_t)T * (fd + 1);
  int kv_in_smem = (fixed + kv <= 200 * 1024) ? 1 : 0;
  AVI_REQUIRE(kv_in_smem || kv_scratch != nullptr, "avi_ff_decoder_ar: kv_scratch required for T=%d fd=%d", T, fd);
  const size_t smem = fixed + (kv_in_smem ? kv : 0);
  AVI_REQUIRE(smem <= 227 * 1024, "avi_ff_decoder_ar: sequence too long (T=%d)", T);
  static SmemOptIn optin;
  cudaError_t e = smem_optin(ff_decoder_ar_kernel, 227 * 1024, optin);
  AVI_REQUIRE(e == cudaSuccess, "avi_ff_decoder_ar: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  ff_decoder_ar_kernel<<<B, DEC_THREADS, smem, (cudaStream_t)stream>>>(*w, cross, style, hidden_out, kv_scratch, T, fd, period,
                                                                       kv_in_smem);
  return check_launch("ff_decoder_ar");
}

extern "C" int avi_ff_biased_attn(const float* qkv, float* out, int32_t B, int32_t T, int32_t fd, int32_t period, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && fd % DEC_NH == 0 && period > 0, "avi_ff_biased_attn: bad arguments");
  const int64_t warps = (int64_t)B * DEC_NH * T;
  const size_t smem = sizeof(float) * 4 * (size_t)T;
  AVI_REQUIRE(smem <= 48 * 1024, "avi_ff_biased_attn: T=%d too long", T);
  ff_biased_attn_kernel<<<(unsigned)((warps + 3) / 4), 128, smem, (cudaStream_t)stream>>>(qkv, out, B, T, fd, period);
  return check_launch("ff_biased_attn");
}
