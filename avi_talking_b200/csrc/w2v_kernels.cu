// wav2vec2 non-GEMM kernels: conv0+GroupNorm+GELU, time resample + LayerNorm, LayerNorm,
// positional grouped conv (+GELU+residual+LayerNorm), multi-head attention.
// Numerics follow HF Wav2Vec2 as wrapped by models/lib/wav2vec.py (reference file:line in include/avi_b200.h).
#include "common.cuh"

namespace avi {

// ------------------------------------------------------------------------------------------------
// conv0: y[b,t,c] = sum_j w[c,j] * x[b, 5t+j];  GroupNorm(groups == channels) over t;  GELU.
// Pass 1 accumulates per-(b,c) sum / sum of squares (fp32 inside a 256-step chunk, fp64 across chunks);
// pass 2 recomputes the 10-tap conv (cheaper than storing 65 MB/clip of fp32) and writes the normalised,
// activated result time-major.
// ------------------------------------------------------------------------------------------------
constexpr int C0_K = 10, C0_S = 5, C0_TT = 256;

__global__ void __launch_bounds__(512) conv0_stats_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                                          double* __restrict__ stats, int n_samples, int L0, int C) {
  __shared__ float xs[C0_TT * C0_S + C0_K];
  const int b = blockIdx.y, t0 = blockIdx.x * C0_TT;
  const int nt = min(C0_TT, L0 - t0);
  const float* x = audio + (int64_t)b * n_samples + (int64_t)t0 * C0_S;
  const int need = (nt - 1) * C0_S + C0_K;
  for (int i = threadIdx.x; i < need; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float wr[C0_K];
#pragma unroll
    for (int j = 0; j < C0_K; ++j) wr[j] = w[c * C0_K + j];
    float s = 0.f, ss = 0.f;
    for (int t = 0; t < nt; ++t) {
      float y = 0.f;
#pragma unroll
      for (int j = 0; j < C0_K; ++j) y = fmaf(wr[j], xs[t * C0_S + j], y);
      s += y;
      ss = fmaf(y, y, ss);
    }
    atomicAdd(&stats[((int64_t)b * C + c) * 2 + 0], (double)s);
    atomicAdd(&stats[((int64_t)b * C + c) * 2 + 1], (double)ss);
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256) conv0_apply_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                                          const float* __restrict__ gn_w, const float* __restrict__ gn_b,
                                                          const double* __restrict__ stats, OutT* __restrict__ out,
                                                          int64_t out_batch_stride, int n_samples, int L0, int C, float eps) {
  constexpr int TT = 64;
  __shared__ float xs[TT * C0_S + C0_K];
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  const int nt = min(TT, L0 - t0);
  const float* x = audio + (int64_t)b * n_samples + (int64_t)t0 * C0_S;
  const int need = (nt - 1) * C0_S + C0_K;
  for (int i = threadIdx.x; i < need; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  // each thread owns a pair of adjacent channels -> 8-byte (fp32) / 4-byte (bf16) coalesced stores
  for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) {
    float wr[2][C0_K], scale[2], shift[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int j = 0; j < C0_K; ++j) wr[u][j] = w[(c + u) * C0_K + j];
      const double s = stats[((int64_t)b * C + c + u) * 2], ss = stats[((int64_t)b * C + c + u) * 2 + 1];
      const double mean = s / L0;
      const double var = fmax(ss / L0 - mean * mean, 0.0);
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      scale[u] = rstd * gn_w[c + u];
      shift[u] = gn_b[c + u] - (float)mean * scale[u];
    }
    OutT* o = out + (int64_t)b * out_batch_stride + (int64_t)t0 * C + c;
    for (int t = 0; t < nt; ++t) {
      float y0 = 0.f, y1 = 0.f;
#pragma unroll
      for (int j = 0; j < C0_K; ++j) {
        const float xv = xs[t * C0_S + j];
        y0 = fmaf(wr[0][j], xv, y0);
        y1 = fmaf(wr[1][j], xv, y1);
      }
      if constexpr (sizeof(OutT) == 2) {  // bf16 output: the 1.5e-7-accurate branch-free GELU is far below the output rounding
        y0 = gelu_fast(fmaf(y0, scale[0], shift[0]));
        y1 = gelu_fast(fmaf(y1, scale[1], shift[1]));
      } else {
        y0 = gelu_erf(fmaf(y0, scale[0], shift[0]));
        y1 = gelu_erf(fmaf(y1, scale[1], shift[1]));
      }
      if constexpr (sizeof(OutT) == 2) {
        *reinterpret_cast<__nv_bfloat162*>(o + (int64_t)t * C) = __floats2bfloat162_rn(y0, y1);
      } else {
        *reinterpret_cast<float2*>(o + (int64_t)t * C) = make_float2(y0, y1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Linear resample (align_corners=True) over time fused with LayerNorm(C). One warp per output row.
// ------------------------------------------------------------------------------------------------
template <int MAXPER>
__global__ void __launch_bounds__(256) lerp_ln_kernel(const void* __restrict__ in, int in_dtype, int64_t in_batch_stride,
                                                      const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                      float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                                                      int B, int T_in, int T_out, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * T_out) return;
  const int b = row / T_out, t = row % T_out;
  // ATen upsample_linear1d, align_corners: scale = (in-1)/(out-1) in float, src = scale*t
  const float scale = T_out > 1 ? (float)(T_in - 1) / (float)(T_out - 1) : 0.f;
  const float src = scale * (float)t;
  const int i0 = (int)src;
  const int i1 = i0 + (i0 < T_in - 1 ? 1 : 0);
  const float l1 = src - (float)i0, l0 = 1.f - l1;
  const int64_t base0 = (int64_t)b * in_batch_stride + (int64_t)i0 * C;
  const int64_t base1 = (int64_t)b * in_batch_stride + (int64_t)i1 * C;
  float v[MAXPER];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < MAXPER; ++u) {
    const int c = lane + u * 32;
    v[u] = 0.f;
    if (c < C) {
      v[u] = l0 * load_as_float(in, in_dtype, base0 + c) + l1 * load_as_float(in, in_dtype, base1 + c);
      s += v[u];
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int u = 0; u < MAXPER; ++u) {
    const int c = lane + u * 32;
    if (c < C) q += (v[u] - mean) * (v[u] - mean);
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int u = 0; u < MAXPER; ++u) {
    const int c = lane + u * 32;
    if (c < C) {
      const float y = (v[u] - mean) * rstd * ln_w[c] + ln_b[c];
      if (out_f32) out_f32[(int64_t)row * C + c] = y;
      if (out_bf16) out_bf16[(int64_t)row * C + c] = __float2bfloat16_rn(y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim, optionally of (x + gelu(p)) (the positional-conv merge). One warp per row.
// ------------------------------------------------------------------------------------------------
template <int MAXPER, int MODE>  // MODE 0: LN(x); 1: LN(x + gelu(p)); 2: LN(x + p)
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* p,  // p may alias out_f32
                                                        const float* __restrict__ w, const float* __restrict__ bsh,
                                                        float* out_f32, __nv_bfloat16* __restrict__ out_bf16,
                                                        int64_t rows, int C, float eps) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[MAXPER];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < MAXPER; ++u) {
    const int c = lane + u * 32;
    v[u] = 0.f;
    if (c < C) {
      v[u] = x[row * C + c];
      if constexpr (MODE == 1) v[u] += gelu_erf(p[row * C + c]);
      if constexpr (MODE == 2) v[u] += p[row * C + c];
      s += v[u];
    }
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int u = 0; u < MAXPER; ++u) {
    const int c = lane + u * 32;
    if (c < C) q += (v[u] - mean) * (v[u] - mean);
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int u = 0; u < MAXPER; ++u) {
    const int c = lane + u * 32;
    if (c < C) {
      const float y = (v[u] - mean) * rstd * w[c] + bsh[c];
      if (out_f32) out_f32[row * C + c] = y;
      if (out_bf16) out_bf16[row * C + c] = __float2bfloat16_rn(y);
    }
  }
}

// Vectorised variant for C % 128 == 0 (512 / 768): NV float4 per lane, 16-byte loads/stores, 8-byte bf16 stores.
template <int NV, int MODE>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* __restrict__ x, const float* p,  // p may alias out_f32
                                                            const float* __restrict__ w, const float* __restrict__ bsh,
                                                            float* out_f32, __nv_bfloat16* __restrict__ out_bf16, int64_t rows,
                                                            float eps) {
  constexpr int C = NV * 128;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_wait();   // launched as a programmatic dependent of the GEMM that produced x
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    v[u] = xr[lane + 32 * u];
    if constexpr (MODE != 0) {
      const float4 q = reinterpret_cast<const float4*>(p + row * C)[lane + 32 * u];
      if constexpr (MODE == 1) {
        v[u].x += gelu_erf(q.x); v[u].y += gelu_erf(q.y); v[u].z += gelu_erf(q.z); v[u].w += gelu_erf(q.w);
      } else {
        v[u].x += q.x; v[u].y += q.y; v[u].z += q.z; v[u].w += q.w;
      }
    }
    s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q2 = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
    q2 += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q2) * (1.f / C) + eps);
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const float4 g = reinterpret_cast<const float4*>(w)[lane + 32 * u];
    const float4 bb = reinterpret_cast<const float4*>(bsh)[lane + 32 * u];
    float4 y;
    y.x = (v[u].x - mean) * rstd * g.x + bb.x;
    y.y = (v[u].y - mean) * rstd * g.y + bb.y;
    y.z = (v[u].z - mean) * rstd * g.z + bb.z;
    y.w = (v[u].w - mean) * rstd * g.w + bb.w;
    if (out_f32) reinterpret_cast<float4*>(out_f32 + row * C)[lane + 32 * u] = y;
    if (out_bf16) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(out_bf16 + row * C)[lane + 32 * u] = pk;
    }
  }
}

// Wide rows (1024 < C <= 8192: the 4096- and 2048-wide LayerNorms of BrainNetwork's training step): one 256-thread block per row,
// up to 32 values per thread, two block reductions (mean, then the centred second moment - the same two-pass arithmetic).
__global__ void __launch_bounds__(256) layernorm_wide_kernel(const float* __restrict__ x, const float* __restrict__ p,
                                                             const float* __restrict__ w, const float* __restrict__ bsh,
                                                             float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int C,
                                                             float eps) {
  __shared__ float red[64];
  const int64_t row = blockIdx.x;
  float v[32];
  float s = 0.f, unused = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = threadIdx.x + u * 256;
    v[u] = 0.f;
    if (c < C) {
      v[u] = x[row * C + c] + (p ? p[row * C + c] : 0.f);
      s += v[u];
    }
  }
  block_sum2(s, unused, red);
  const float mean = s / C;
  float q = 0.f;
  unused = 0.f;
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = threadIdx.x + u * 256;
    if (c < C) q += (v[u] - mean) * (v[u] - mean);
  }
  block_sum2(q, unused, red);
  const float rstd = rsqrtf(q / C + eps);
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int c = threadIdx.x + u * 256;
    if (c < C) {
      const float y = (v[u] - mean) * rstd * w[c] + bsh[c];
      if (out_f32) out_f32[row * C + c] = y;
      if (out_bf16) out_bf16[row * C + c] = __float2bfloat16_rn(y);
    }
  }
}

template <int MODE>
static void launch_layernorm(const float* x, const float* p, const float* w, const float* b, float* o32, __nv_bfloat16* o16,
                             int64_t rows, int C, float eps, cudaStream_t st) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const bool al = (((uintptr_t)x | (uintptr_t)p | (uintptr_t)w | (uintptr_t)b | (uintptr_t)o32) % 16 == 0) && ((uintptr_t)o16 % 8 == 0);
  if (al && C == 768) launch_pdl(layernorm_vec_kernel<6, MODE>, dim3(grid), dim3(256), 0, st, x, p, w, b, o32, o16, rows, eps);
  else if (al && C == 512) launch_pdl(layernorm_vec_kernel<4, MODE>, dim3(grid), dim3(256), 0, st, x, p, w, b, o32, o16, rows, eps);
  else layernorm_kernel<32, MODE><<<grid, 256, 0, st>>>(x, p, w, b, o32, o16, rows, C, eps);
}

// ------------------------------------------------------------------------------------------------
// Positional grouped conv, CUDA-core fp32: pc[b,t,co] = bias[co] + sum_{j,ci} xpad[b, t+j-pad, g*CG+ci] * wk[g][j][ci][co']
// Block = (t-tile of 32, group, clip); 192 threads = CG(48) outputs x 4 time sub-tiles of 8 (sliding register window).
// ------------------------------------------------------------------------------------------------
constexpr int PC_TT = 32, PC_JC = 8;

template <int CG>
__global__ void __launch_bounds__(CG * 4) posconv_kernel(const float* __restrict__ x, const float* __restrict__ wk,
                                                          const float* __restrict__ bias, float* __restrict__ pc, int T, int C,
                                                          int k) {
  extern __shared__ float smem[];
  const int pad = k / 2;
  const int rows = PC_TT + k - 1;
  float* xs = smem;                 // [rows][CG]
  float* ws = smem + rows * CG;     // [PC_JC][CG ci][CG co]
  const int t0 = blockIdx.x * PC_TT, g = blockIdx.y, b = blockIdx.z;
  const int co = threadIdx.x % CG, tg = threadIdx.x / CG;  // tg in 0..3 -> 8 time steps each
  for (int i = threadIdx.x; i < rows * CG; i += blockDim.x) {
    const int r = i / CG, ci = i % CG;
    const int t = t0 + r - pad;
    xs[i] = (t >= 0 && t < T) ? x[((int64_t)b * T + t) * C + g * CG + ci] : 0.f;
  }
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  const float* wg = wk + (int64_t)g * k * CG * CG;
  for (int j0 = 0; j0 < k; j0 += PC_JC) {
    __syncthreads();
    for (int i = threadIdx.x; i < PC_JC * CG * CG; i += blockDim.x) ws[i] = wg[(int64_t)j0 * CG * CG + i];
    __syncthreads();
    for (int ci = 0; ci < CG; ++ci) {
      // window of inputs for this thread's 8 outputs and PC_JC taps: rows tg*8 + j0 .. + 8 + PC_JC - 2
      float win[8 + PC_JC - 1];
#pragma unroll
      for (int u = 0; u < 8 + PC_JC - 1; ++u) win[u] = xs[(tg * 8 + j0 + u) * CG + ci];
#pragma unroll
      for (int jj = 0; jj < PC_JC; ++jj) {
        const float wv = ws[(jj * CG + ci) * CG + co];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaf(wv, win[u + jj], acc[u]);
      }
    }
  }
  const float bv = bias[g * CG + co];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int t = t0 + tg * 8 + u;
    if (t < T) pc[((int64_t)b * T + t) * C + g * CG + co] = acc[u] + bv;
  }
}

// ------------------------------------------------------------------------------------------------
// Multi-head attention, fp32 CUDA-core flash-style: one warp per query row, key tiles of 128 in smem.
// qkv [B,T,3*H*D] (q|k|v), D == 64. Block = 16 warps x 4 rows = 64 query rows of one (clip, head).
// ------------------------------------------------------------------------------------------------
constexpr int MHA_KT = 128, MHA_WARPS = 16, MHA_RPW = 4, MHA_D = 64;

template <typename T>
__global__ void __launch_bounds__(MHA_WARPS * 32) mha_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tlen, int H,
                                                              float scale) {
  extern __shared__ float smem[];
  float* Ks = smem;                               // [KT][D+1]
  float* Vs = smem + MHA_KT * (MHA_D + 1);        // [KT][D]
  float* Qs = Vs + MHA_KT * MHA_D;                // [WARPS*RPW][D]
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (MHA_WARPS * MHA_RPW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = H * MHA_D;
  const T* base = qkv + (int64_t)b * Tlen * 3 * E;
  for (int i = threadIdx.x; i < MHA_WARPS * MHA_RPW * MHA_D; i += blockDim.x) {
    const int r = i / MHA_D, d = i % MHA_D;
    const int t = q0 + r;
    Qs[i] = (t < Tlen) ? (float)base[(int64_t)t * 3 * E + h * MHA_D + d] * scale : 0.f;
  }
  float m[MHA_RPW], l[MHA_RPW], o0[MHA_RPW], o1[MHA_RPW];
#pragma unroll
  for (int r = 0; r < MHA_RPW; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
    o0[r] = 0.f;
    o1[r] = 0.f;
  }
  for (int k0 = 0; k0 < Tlen; k0 += MHA_KT) {
    const int nk = min(MHA_KT, Tlen - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < MHA_KT * MHA_D; i += blockDim.x) {
      const int r = i / MHA_D, d = i % MHA_D;
      float kv = 0.f, vv = 0.f;
      if (r < nk) {
        const T* p = base + (int64_t)(k0 + r) * 3 * E + h * MHA_D + d;
        kv = (float)p[E];
        vv = (float)p[2 * E];
      }
      Ks[r * (MHA_D + 1) + d] = kv;
      Vs[r * MHA_D + d] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MHA_RPW; ++r) {
      const float* q = Qs + (warp * MHA_RPW + r) * MHA_D;
      float s[MHA_KT / 32];
#pragma unroll
      for (int u = 0; u < MHA_KT / 32; ++u) s[u] = 0.f;
#pragma unroll 8
      for (int d = 0; d < MHA_D; ++d) {
        const float qv = q[d];
#pragma unroll
        for (int u = 0; u < MHA_KT / 32; ++u) s[u] = fmaf(qv, Ks[(lane + 32 * u) * (MHA_D + 1) + d], s[u]);
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < MHA_KT / 32; ++u) {
        if (lane + 32 * u >= nk) s[u] = -INFINITY;
        tmax = fmaxf(tmax, s[u]);
      }
      tmax = warp_max(tmax);
      const float mnew = fmaxf(m[r], tmax);
      const float corr = __expf(m[r] - mnew);  // m[r] = -inf on the first tile -> 0
      float psum = 0.f;
#pragma unroll
      for (int u = 0; u < MHA_KT / 32; ++u) {
        s[u] = __expf(s[u] - mnew);
        psum += s[u];
      }
      l[r] = l[r] * corr + warp_sum(psum);
      float a0 = o0[r] * corr, a1 = o1[r] * corr;
#pragma unroll
      for (int u = 0; u < MHA_KT / 32; ++u) {
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
          const float p = __shfl_sync(0xffffffffu, s[u], kk);
          const float* vr = Vs + (u * 32 + kk) * MHA_D;
          a0 = fmaf(p, vr[lane], a0);
          a1 = fmaf(p, vr[lane + 32], a1);
        }
      }
      o0[r] = a0;
      o1[r] = a1;
      m[r] = mnew;
    }
  }
#pragma unroll
  for (int r = 0; r < MHA_RPW; ++r) {
    const int t = q0 + warp * MHA_RPW + r;
    if (t < Tlen) {
      T* o = out + ((int64_t)b * Tlen + t) * E + h * MHA_D;
      const float inv = 1.f / l[r];
      o[lane] = (T)(o0[r] * inv);
      o[lane + 32] = (T)(o1[r] * inv);
    }
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<__nv_bfloat162*>(dst + i) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(dst + i + 2) = __floats2bfloat162_rn(v.z, v.w);
  } else {
    for (; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// dst[b, t, :] (bf16, rows_out rows per clip) = t in [front, front+T) ? src[b, t-front, :] : 0
__global__ void pad_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int T, int C, int front,
                                int rows_out, int64_t total4) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const int64_t i = i4 * 4;
  const int c = (int)(i % C);
  const int64_t r = i / C;
  const int t = (int)(r % rows_out) - front;
  const int64_t b = r / rows_out;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t >= 0 && t < T) v = *reinterpret_cast<const float4*>(src + ((int64_t)b * T + t) * C + c);
  *reinterpret_cast<__nv_bfloat162*>(dst + i) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162*>(dst + i + 2) = __floats2bfloat162_rn(v.z, v.w);
}

// split-bf16 operand for near-fp32 GEMMs on the bf16 tensor path: dst[r] = [hi(x) | lo(x) | hi(x)], lo = bf16(x - hi)
__global__ void split_bf16x3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows, int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  const float x = src[i];
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  __nv_bfloat16* d = dst + r * 3 * K;
  d[k] = hi;
  d[K + k] = lo;
  d[2 * K + k] = hi;
}

// General form for fp32-accurate GEMMs on the bf16 tensor path: x = s0 + s1 + s2 (+ 2^-25 |x|) with s0 = bf16(x), s1 = bf16(x - s0),
// s2 = bf16(x - s0 - s1); dst[r] = [s_{p0}(x) | s_{p1}(x) | ...] over `nterms` K-wide blocks, p_t = bits 2t..2t+1 of `pattern`.
// An operand pair (A pattern, W pattern) whose blocks line up as the products s_i * t_j to keep gives
//   3 terms  A = [0,1,0], W = [0,0,1]             : s0 t0 + s1 t0 + s0 t1                (dropped: 2^-16 relative)
//   6 terms  A = [0,0,0,1,1,2], W = [0,1,2,0,1,0] : every product down to 2^-24 relative (fp32 accuracy)
__global__ void __launch_bounds__(256) split_bf16_terms_kernel(const float4* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n4,
                                                               int K4, int nterms, uint32_t pattern) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int64_t r = i / K4;
  const int k = (int)(i - r * K4) * 4;
  const float4 x4 = src[i];
  const float x[4] = {x4.x, x4.y, x4.z, x4.w};
  __nv_bfloat16 s[3][4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    s[0][e] = __float2bfloat16_rn(x[e]);
    const float r1 = x[e] - __bfloat162float(s[0][e]);
    s[1][e] = __float2bfloat16_rn(r1);
    s[2][e] = __float2bfloat16_rn(r1 - __bfloat162float(s[1][e]));
  }
  const int K = K4 * 4;
  __nv_bfloat16* d = dst + r * (int64_t)nterms * K + k;
  for (int t = 0; t < nterms; ++t) {
    const int which = (pattern >> (2 * t)) & 3u;
    const __nv_bfloat16* v = which == 0 ? s[0] : (which == 1 ? s[1] : s[2]);
    __nv_bfloat162 lo2, hi2;
    lo2.x = v[0]; lo2.y = v[1]; hi2.x = v[2]; hi2.y = v[3];
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&lo2);
    w.y = *reinterpret_cast<uint32_t*>(&hi2);
    *reinterpret_cast<uint2*>(d + (int64_t)t * K) = w;
  }
}

}  // namespace avi

using namespace avi;

extern "C" int avi_split_bf16_terms(const float* src, void* dst, int64_t rows, int32_t K, int32_t nterms, uint32_t pattern, void* stream) {
  AVI_REQUIRE(rows > 0 && K > 0 && K % 4 == 0 && nterms >= 1 && nterms <= 8, "avi_split_bf16_terms: K %% 4 == 0 and 1..8 terms (K=%d)", K);
  AVI_REQUIRE((((uintptr_t)src % 16) | ((uintptr_t)dst % 8)) == 0, "avi_split_bf16_terms: unaligned buffers");
  for (int t = 0; t < nterms; ++t) AVI_REQUIRE(((pattern >> (2 * t)) & 3u) < 3u, "avi_split_bf16_terms: pattern entries are 0, 1 or 2");
  const int64_t n4 = rows * (K / 4);
  split_bf16_terms_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)src, (__nv_bfloat16*)dst, n4, K / 4,
                                                                                         nterms, pattern);
  return check_launch("split_bf16_terms");
}

extern "C" int avi_pad_cast_bf16(const float* src, void* dst, int32_t B, int32_t T, int32_t C, int32_t front, int32_t rows_out,
                                 void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && C > 0 && C % 4 == 0 && front >= 0 && rows_out >= front + T, "avi_pad_cast_bf16: bad shape");
  const int64_t total4 = (int64_t)B * rows_out * C / 4;
  pad_cast_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, T, C, front,
                                                                                    rows_out, total4);
  return check_launch("pad_cast");
}

extern "C" int avi_split_bf16x3(const float* src, void* dst, int64_t rows, int32_t K, void* stream) {
  AVI_REQUIRE(rows > 0 && K > 0, "avi_split_bf16x3: bad shape");
  const int64_t n = rows * K;
  split_bf16x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, rows, K);
  return check_launch("split_bf16x3");
}

extern "C" int avi_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n <= 0) return 0;
  AVI_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0), "avi_cast_f32_to_bf16: unaligned pointers");
  const int64_t nth = (n + 3) / 4;
  cast_f32_bf16_kernel<<<(unsigned)((nth + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  return check_launch("cast_f32_bf16");
}

extern "C" int avi_w2v_conv0_gn_gelu(const float* audio, const float* w, const float* gn_w, const float* gn_b, void* stats,
                                     void* out, int32_t out_dtype, int64_t out_batch_stride, int32_t B, int32_t n_samples,
                                     int32_t C, float eps, void* stream) {
  AVI_REQUIRE(B > 0 && n_samples >= C0_K && C > 0 && C % 2 == 0, "avi_w2v_conv0_gn_gelu: bad shape B=%d n=%d C=%d", B, n_samples, C);
  const int L0 = (n_samples - C0_K) / C0_S + 1;
  AVI_REQUIRE(out_batch_stride >= (int64_t)L0 * C, "avi_w2v_conv0_gn_gelu: out_batch_stride too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)B * C, st) != cudaSuccess) {
    set_error("avi_w2v_conv0_gn_gelu: memset failed");
    return 1;
  }
  conv0_stats_kernel<<<dim3((L0 + C0_TT - 1) / C0_TT, B), 512, 0, st>>>(audio, w, (double*)stats, n_samples, L0, C);
  if (check_launch("conv0_stats")) return 1;
  dim3 grid((L0 + 63) / 64, B);
  if (out_dtype == AVI_DT_BF16)
    conv0_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(audio, w, gn_w, gn_b, (const double*)stats, (__nv_bfloat16*)out,
                                                           out_batch_stride, n_samples, L0, C, eps);
  else
    conv0_apply_kernel<float><<<grid, 256, 0, st>>>(audio, w, gn_w, gn_b, (const double*)stats, (float*)out, out_batch_stride,
                                                   n_samples, L0, C, eps);
  return check_launch("conv0_apply");
}

extern "C" int avi_w2v_lerp_layernorm(const void* in, int32_t in_dtype, int64_t in_batch_stride, const float* ln_w,
                                      const float* ln_b, float* out_f32, void* out_bf16, int32_t B, int32_t T_in, int32_t T_out,
                                      int32_t C, float eps, void* stream) {
  AVI_REQUIRE(B > 0 && T_in > 0 && T_out > 0 && C > 0 && C <= 1024, "avi_w2v_lerp_layernorm: bad shape");
  const int rows = B * T_out;
  lerp_ln_kernel<32><<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(in, in_dtype, in_batch_stride, ln_w, ln_b, out_f32,
                                                                     (__nv_bfloat16*)out_bf16, B, T_in, T_out, C, eps);
  return check_launch("lerp_ln");
}

extern "C" int avi_layernorm(const float* x, const float* res, const float* w, const float* b, float* out_f32, void* out_bf16,
                             int64_t rows, int32_t C, float eps, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0 && C <= 8192, "avi_layernorm: bad shape rows=%lld C=%d", (long long)rows, C);
  if (C > 1024) {
    layernorm_wide_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, res, w, b, out_f32, (__nv_bfloat16*)out_bf16, C, eps);
    return check_launch("layernorm_wide");
  }
  if (res)
    launch_layernorm<2>(x, res, w, b, out_f32, (__nv_bfloat16*)out_bf16, rows, C, eps, (cudaStream_t)stream);
  else
    launch_layernorm<0>(x, nullptr, w, b, out_f32, (__nv_bfloat16*)out_bf16, rows, C, eps, (cudaStream_t)stream);
  return check_launch("layernorm");
}

// pc scratch is carried inside out_f32: the conv writes its raw output there, then the merge kernel
// overwrites it in place with LayerNorm(x + gelu(pc)) (each warp reads its row fully before writing it).
extern "C" int avi_w2v_posconv_ln(const float* x, const float* w_packed, const float* conv_bias, const float* ln_w,
                                  const float* ln_b, float* out_f32, void* out_bf16, int32_t B, int32_t T, int32_t C,
                                  int32_t groups, int32_t k, float eps, void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && groups > 0 && C % groups == 0 && C / groups == 48 && k % PC_JC == 0 && C <= 1024,
              "avi_w2v_posconv_ln: unsupported shape C=%d groups=%d k=%d", C, groups, k);
  AVI_REQUIRE(out_f32 != nullptr, "avi_w2v_posconv_ln: out_f32 is required (it doubles as the conv scratch)");
  constexpr int CG = 48;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(float) * ((PC_TT + k - 1) * CG + PC_JC * CG * CG);
  static SmemOptIn optin;
  AVI_REQUIRE(smem_optin(posconv_kernel<CG>, (int)smem, optin) == cudaSuccess, "avi_w2v_posconv_ln: cannot opt in to %zu bytes of shared memory", smem);
  posconv_kernel<CG><<<dim3((T + PC_TT - 1) / PC_TT, groups, B), CG * 4, smem, st>>>(x, w_packed, conv_bias, out_f32, T, C, k);
  if (check_launch("posconv")) return 1;
  const int64_t rows = (int64_t)B * T;
  layernorm_kernel<32, 1><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, out_f32, ln_w, ln_b, out_f32,
                                                                        (__nv_bfloat16*)out_bf16, rows, C, eps);
  return check_launch("posconv_merge_ln");
}

extern "C" int avi_w2v_posconv_merge_ln(const float* x, const float* pc, const float* ln_w, const float* ln_b, float* out_f32,
                                        void* out_bf16, int64_t rows, int32_t C, float eps, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0 && C <= 1024, "avi_w2v_posconv_merge_ln: bad shape");
  launch_layernorm<1>(x, pc, ln_w, ln_b, out_f32, (__nv_bfloat16*)out_bf16, rows, C, eps, (cudaStream_t)stream);
  return check_launch("posconv_merge_ln");
}

extern "C" int avi_mha_fwd(const void* qkv, void* out, int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t D, float scale,
                           void* stream) {
  AVI_REQUIRE(B > 0 && T > 0 && H > 0 && D == MHA_D, "avi_mha_fwd: head dim must be 64 (got %d)", D);
  const size_t smem = sizeof(float) * (MHA_KT * (MHA_D + 1) + MHA_KT * MHA_D + MHA_WARPS * MHA_RPW * MHA_D);
  static SmemOptIn optin_f, optin_h;
  AVI_REQUIRE(smem_optin(mha_kernel<float>, (int)smem, optin_f) == cudaSuccess &&
                  smem_optin(mha_kernel<__nv_bfloat16>, (int)smem, optin_h) == cudaSuccess,
              "avi_mha_fwd: cannot opt in to %zu bytes of shared memory", smem);
  dim3 grid((T + MHA_WARPS * MHA_RPW - 1) / (MHA_WARPS * MHA_RPW), H, B);
  if (dtype == AVI_DT_BF16)
    mha_kernel<__nv_bfloat16><<<grid, MHA_WARPS * 32, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv,
                                                                                   (__nv_bfloat16*)out, T, H, scale);
  else
    mha_kernel<float><<<grid, MHA_WARPS * 32, smem, (cudaStream_t)stream>>>((const float*)qkv, (float*)out, T, H, scale);
  return check_launch("mha");
}
