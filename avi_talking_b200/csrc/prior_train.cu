// Training-side kernels of the text->style diffusion prior: everything of one optimisation step of
// train_diffusion_prior.py:422-499 that is not a dense contraction (those go through avi_gemm_*):
//   * token assembly of VersatileDiffusionPriorNetwork.forward (models/diffusion_prior.py:258-306) fused with q_sample
//     (p_losses :369-373) and its backward (null-embedding / learned-query gradients),
//   * the 3-query x 4-key cosine-similarity attention of dalle2_pytorch.Attention (scale 16, rotary on the first 32 features,
//     one shared key/value head, null key/value, T5 relative-position bias), forward with saved probabilities and backward,
//   * SwiGLU, SiLU, row l2-normalisation, the `stable` LayerNorm pre-division by the (detached) row maximum,
//   * soft_clip_loss (train_diffusion_prior.py:125-133) forward + gradient,
//   * AdamW with decoupled weight decay (torch.optim.AdamW as built at train_diffusion_prior.py:996-1004).
// All fp32; the shapes are tiny (3 tokens x 128 features per sample), so the kernels are organised for few launches and
// deterministic reductions (per-sample partials that avi_colsum folds), not for bandwidth.
#include "common.cuh"

namespace avi {

constexpr int PT_DIM = 128, PT_H = 8, PT_D = 64, PT_N = 3, PT_J = 4, PT_INNER = PT_H * PT_D, PT_ROT_PAIRS = 16;

// ------------------------------------------------------------------------------------------------ tokens (+ q_sample)
// tokens[b] = [ keep_b ? brain[b] : null_brain ; temb[b] ; (keep_i ? sa[t_b] * x0[b] + s1[t_b] * noise[b] : null_image) + learned_query ]
__global__ void __launch_bounds__(PT_DIM) prior_tokens_fwd_kernel(const float* __restrict__ brain, const float* __restrict__ null_brain,
                                                                  const float* __restrict__ keep_b, const float* __restrict__ x0,
                                                                  const float* __restrict__ noise, const float* __restrict__ sqrt_ac,
                                                                  const float* __restrict__ sqrt_1mac, const int32_t* __restrict__ times,
                                                                  const float* __restrict__ null_image, const float* __restrict__ keep_i,
                                                                  const float* __restrict__ lq, const float* __restrict__ temb,
                                                                  float* __restrict__ tokens, float* __restrict__ x_noisy) {
  const int b = blockIdx.x, c = threadIdx.x;
  const int t = times[b];
  const float xt = sqrt_ac[t] * x0[b * PT_DIM + c] + sqrt_1mac[t] * noise[b * PT_DIM + c];
  if (x_noisy) x_noisy[b * PT_DIM + c] = xt;
  float* o = tokens + (int64_t)b * PT_N * PT_DIM;
  o[c] = keep_b[b] != 0.f ? brain[b * PT_DIM + c] : null_brain[c];
  o[PT_DIM + c] = temb[b * PT_DIM + c];
  o[2 * PT_DIM + c] = (keep_i[b] != 0.f ? xt : null_image[c]) + lq[c];
}

// blocks 0..B-1: per-sample gradients; block B: the three parameter gradients, summed over the batch in a fixed order
__global__ void __launch_bounds__(PT_DIM) prior_tokens_bwd_kernel(const float* __restrict__ dtok, const float* __restrict__ keep_b,
                                                                  const float* __restrict__ keep_i, float* __restrict__ dbrain,
                                                                  float* __restrict__ dtemb, float* __restrict__ dnull_brain,
                                                                  float* __restrict__ dnull_image, float* __restrict__ dlq, int B) {
  const int c = threadIdx.x;
  if ((int)blockIdx.x < B) {
    const int b = blockIdx.x;
    const float* d = dtok + (int64_t)b * PT_N * PT_DIM;
    dbrain[b * PT_DIM + c] = keep_b[b] != 0.f ? d[c] : 0.f;
    dtemb[b * PT_DIM + c] = d[PT_DIM + c];
    return;
  }
  float nb = 0.f, ni = 0.f, q = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* d = dtok + (int64_t)b * PT_N * PT_DIM;
    if (keep_b[b] == 0.f) nb += d[c];
    if (keep_i[b] == 0.f) ni += d[2 * PT_DIM + c];
    q += d[2 * PT_DIM + c];
  }
  dnull_brain[c] += nb;
  dnull_image[c] += ni;
  dlq[c] += q;
}

// ------------------------------------------------------------------------------------------------ attention (3 x 4, 8 heads x 64)
// lane l of a warp owns features (2l, 2l+1): exactly the interleaved pairs rotary_embedding_torch rotates (pairs 0..15).
__device__ __forceinline__ float2 rot_fwd(float2 x, const float* __restrict__ rot, int pos, int lane) {
  if (lane >= PT_ROT_PAIRS) return x;
  const float c = rot[(pos * PT_ROT_PAIRS + lane) * 2], s = rot[(pos * PT_ROT_PAIRS + lane) * 2 + 1];
  return make_float2(x.x * c - x.y * s, x.y * c + x.x * s);
}
__device__ __forceinline__ float2 rot_bwd(float2 d, const float* __restrict__ rot, int pos, int lane) {   // transpose of rot_fwd
  if (lane >= PT_ROT_PAIRS) return d;
  const float c = rot[(pos * PT_ROT_PAIRS + lane) * 2], s = rot[(pos * PT_ROT_PAIRS + lane) * 2 + 1];
  return make_float2(d.x * c + d.y * s, d.y * c - d.x * s);
}

// stage the four keys (null + 3 rotated, l2-normalised, x sqrt(16)) and values of sample b; knorm[j] = max(||k_j||, 1e-12)
__device__ __forceinline__ void stage_kv(const float* __restrict__ kv, const float* __restrict__ null_kv, const float* __restrict__ rot,
                                         int b, int warp, int lane, float (*k4)[PT_D], float (*v)[PT_D], float* knorm) {
  if (warp < PT_J) {
    float2 k, vv;
    if (warp == 0) {
      k = *reinterpret_cast<const float2*>(null_kv + 2 * lane);
      vv = *reinterpret_cast<const float2*>(null_kv + PT_D + 2 * lane);
    } else {
      const float* row = kv + ((int64_t)b * PT_N + warp - 1) * (2 * PT_D);
      k = rot_fwd(*reinterpret_cast<const float2*>(row + 2 * lane), rot, warp - 1, lane);
      vv = *reinterpret_cast<const float2*>(row + PT_D + 2 * lane);
    }
    const float nrm = fmaxf(sqrtf(warp_sum(k.x * k.x + k.y * k.y)), 1e-12f);
    k4[warp][2 * lane] = 4.f * k.x / nrm;
    k4[warp][2 * lane + 1] = 4.f * k.y / nrm;
    v[warp][2 * lane] = vv.x;
    v[warp][2 * lane + 1] = vv.y;
    if (lane == 0) knorm[warp] = nrm;
  }
}

// q [3B, 512] raw projections, kv [3B, 128] raw, null_kv [2, 64], rot [3, 16, 2] (cos, sin), bias [8, 3, 4] -> out [3B, 512], P [B, 8, 3, 4]
__global__ void __launch_bounds__(256) prior_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                             const float* __restrict__ null_kv, const float* __restrict__ rot,
                                                             const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ P) {
  __shared__ float k4[PT_J][PT_D], v[PT_J][PT_D], knorm[PT_J];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = warp;
  stage_kv(kv, null_kv, rot, b, warp, lane, k4, v, knorm);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < PT_N; ++i) {
    const int64_t row = (int64_t)b * PT_N + i;
    float2 x = *reinterpret_cast<const float2*>(q + row * PT_INNER + h * PT_D + 2 * lane);
    x = rot_fwd(make_float2(16.f * x.x, 16.f * x.y), rot, i, lane);
    const float nrm = fmaxf(sqrtf(warp_sum(x.x * x.x + x.y * x.y)), 1e-12f);
    const float2 q4 = make_float2(4.f * x.x / nrm, 4.f * x.y / nrm);
    float s[PT_J], m = -INFINITY;
#pragma unroll
    for (int j = 0; j < PT_J; ++j) {
      s[j] = warp_sum(q4.x * k4[j][2 * lane] + q4.y * k4[j][2 * lane + 1]) + bias[(h * PT_N + i) * PT_J + j];
      m = fmaxf(m, s[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PT_J; ++j) {
      s[j] = expf(s[j] - m);
      sum += s[j];
    }
    float2 o = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < PT_J; ++j) {
      s[j] /= sum;
      o.x += s[j] * v[j][2 * lane];
      o.y += s[j] * v[j][2 * lane + 1];
    }
    *reinterpret_cast<float2*>(out + row * PT_INNER + h * PT_D + 2 * lane) = o;
    if (lane < PT_J) P[(((int64_t)b * PT_H + h) * PT_N + i) * PT_J + lane] = s[lane];
  }
}

// -> dq [3B, 512], dkv [3B, 128], dnull [B, 128] (per-sample null key | null value gradients), dS [B, 96] (per-sample bias gradients)
__global__ void __launch_bounds__(256) prior_attn_bwd_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                             const float* __restrict__ null_kv, const float* __restrict__ rot,
                                                             const float* __restrict__ P, const float* __restrict__ dout,
                                                             float* __restrict__ dq, float* __restrict__ dkv, float* __restrict__ dnull,
                                                             float* __restrict__ dS) {
  __shared__ float k4[PT_J][PT_D], v[PT_J][PT_D], knorm[PT_J];
  __shared__ float red_k[PT_H][PT_J][PT_D], red_v[PT_H][PT_J][PT_D];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = warp;
  stage_kv(kv, null_kv, rot, b, warp, lane, k4, v, knorm);
  __syncthreads();
  float2 dk4[PT_J], dv[PT_J];
#pragma unroll
  for (int j = 0; j < PT_J; ++j) dk4[j] = dv[j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < PT_N; ++i) {
    const int64_t row = (int64_t)b * PT_N + i;
    float2 x = *reinterpret_cast<const float2*>(q + row * PT_INNER + h * PT_D + 2 * lane);
    x = rot_fwd(make_float2(16.f * x.x, 16.f * x.y), rot, i, lane);
    const float nrm = fmaxf(sqrtf(warp_sum(x.x * x.x + x.y * x.y)), 1e-12f);
    const float2 qh = make_float2(x.x / nrm, x.y / nrm);     // unit query; q4 = 4 qh
    const float2 d_o = *reinterpret_cast<const float2*>(dout + row * PT_INNER + h * PT_D + 2 * lane);
    float p[PT_J], dp[PT_J], dot = 0.f;
#pragma unroll
    for (int j = 0; j < PT_J; ++j) {
      p[j] = P[(((int64_t)b * PT_H + h) * PT_N + i) * PT_J + j];
      dp[j] = warp_sum(d_o.x * v[j][2 * lane] + d_o.y * v[j][2 * lane + 1]);
      dot += p[j] * dp[j];
    }
    float2 dq4 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < PT_J; ++j) {
      const float ds = p[j] * (dp[j] - dot);
      if (lane == j) dS[(int64_t)b * (PT_H * PT_N * PT_J) + (h * PT_N + i) * PT_J + j] = ds;
      dq4.x += ds * k4[j][2 * lane];
      dq4.y += ds * k4[j][2 * lane + 1];
      dk4[j].x += ds * 4.f * qh.x;
      dk4[j].y += ds * 4.f * qh.y;
      dv[j].x += p[j] * d_o.x;
      dv[j].y += p[j] * d_o.y;
    }
    // q4 = 4 x / ||x||  ->  dx = 4 (dq4 - qh (qh . dq4)) / ||x|| ; then the transposed rotation and the factor 16
    const float proj = warp_sum(qh.x * dq4.x + qh.y * dq4.y);
    float2 dx = make_float2(4.f * (dq4.x - qh.x * proj) / nrm, 4.f * (dq4.y - qh.y * proj) / nrm);
    dx = rot_bwd(dx, rot, i, lane);
    *reinterpret_cast<float2*>(dq + row * PT_INNER + h * PT_D + 2 * lane) = make_float2(16.f * dx.x, 16.f * dx.y);
  }
#pragma unroll
  for (int j = 0; j < PT_J; ++j) {
    red_k[h][j][2 * lane] = dk4[j].x;
    red_k[h][j][2 * lane + 1] = dk4[j].y;
    red_v[h][j][2 * lane] = dv[j].x;
    red_v[h][j][2 * lane + 1] = dv[j].y;
  }
  __syncthreads();
  // warps 0..3: key j = warp (heads summed in a fixed order), through the normalisation and the rotation; warps 4..7: value j = warp - 4
  const int j = warp & 3;
  float2 g = make_float2(0.f, 0.f);
  if (warp < PT_J) {
#pragma unroll
    for (int hh = 0; hh < PT_H; ++hh) {
      g.x += red_k[hh][j][2 * lane];
      g.y += red_k[hh][j][2 * lane + 1];
    }
    const float2 kh = make_float2(k4[j][2 * lane] * 0.25f, k4[j][2 * lane + 1] * 0.25f);
    const float proj = warp_sum(kh.x * g.x + kh.y * g.y);
    g = make_float2(4.f * (g.x - kh.x * proj) / knorm[j], 4.f * (g.y - kh.y * proj) / knorm[j]);
    if (j == 0) {
      *reinterpret_cast<float2*>(dnull + (int64_t)b * (2 * PT_D) + 2 * lane) = g;
    } else {
      g = rot_bwd(g, rot, j - 1, lane);
      *reinterpret_cast<float2*>(dkv + ((int64_t)b * PT_N + j - 1) * (2 * PT_D) + 2 * lane) = g;
    }
  } else {
#pragma unroll
    for (int hh = 0; hh < PT_H; ++hh) {
      g.x += red_v[hh][j][2 * lane];
      g.y += red_v[hh][j][2 * lane + 1];
    }
    if (j == 0) *reinterpret_cast<float2*>(dnull + (int64_t)b * (2 * PT_D) + PT_D + 2 * lane) = g;
    else *reinterpret_cast<float2*>(dkv + ((int64_t)b * PT_N + j - 1) * (2 * PT_D) + PT_D + 2 * lane) = g;
  }
}

// ------------------------------------------------------------------------------------------------ SwiGLU
// h [R, 2I] = (a | gate) -> y = a * silu(gate)
__global__ void swiglu_fwd_kernel(const float* __restrict__ h, float* __restrict__ y, int64_t R, int I) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * I) return;
  const int64_t r = idx / I;
  const int c = (int)(idx - r * I);
  const float a = h[r * 2 * I + c], g = h[r * 2 * I + I + c];
  y[idx] = a * (g / (1.f + expf(-g)));
}
__global__ void swiglu_bwd_kernel(const float* __restrict__ h, const float* __restrict__ dy, float* __restrict__ dh, int64_t R, int I) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * I) return;
  const int64_t r = idx / I;
  const int c = (int)(idx - r * I);
  const float a = h[r * 2 * I + c], g = h[r * 2 * I + I + c], d = dy[idx];
  const float sg = 1.f / (1.f + expf(-g));
  dh[r * 2 * I + c] = d * g * sg;
  dh[r * 2 * I + I + c] = d * a * sg * (1.f + g * (1.f - sg));
}

// ------------------------------------------------------------------------------------------------ row scalings
// mode 0: amax   -> stat[r] = max_c x[r, c]           ; out = x / stat          (dalle2 LayerNorm(stable=True) pre-division)
// mode 1: l2norm -> stat[r] = max(||x[r]||, 1e-12)     ; out = x / stat          (F.normalize)
__global__ void __launch_bounds__(256) rows_stat_div_kernel(const float* __restrict__ x, float* __restrict__ out, float* __restrict__ stat,
                                                            int64_t R, int C, int mode) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* xr = x + r * C;
  float s = mode == 0 ? -INFINITY : 0.f;
  for (int c = lane; c < C; c += 32) s = mode == 0 ? fmaxf(s, xr[c]) : s + xr[c] * xr[c];
  if (mode == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
  } else {
    s = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  }
  for (int c = lane; c < C; c += 32) out[r * C + c] = xr[c] / s;
  if (lane == 0) stat[r] = s;
}
// mode 0: dx = dy / stat (the maximum is detached upstream) ; mode 1: dx = (dy - y (y . dy)) / stat with y = the normalised row
__global__ void __launch_bounds__(256) rows_stat_div_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                                const float* __restrict__ stat, float* __restrict__ dx, int64_t R, int C,
                                                                int mode) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float proj = 0.f;
  if (mode == 1) {
    for (int c = lane; c < C; c += 32) proj += y[r * C + c] * dy[r * C + c];
    proj = warp_sum(proj);
  }
  const float s = stat[r];
  for (int c = lane; c < C; c += 32) dx[r * C + c] = (dy[r * C + c] - (mode == 1 ? y[r * C + c] * proj : 0.f)) / s;
}

__global__ void scale_kernel(const float* __restrict__ a, float alpha, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = alpha * a[i];
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] * b[i];
}

// ------------------------------------------------------------------------------------------------ soft_clip_loss
// log-sum-exp of every row (by_col = 0) or column (1) of M [B, B]; one warp per row / column
__global__ void __launch_bounds__(256) lse_kernel(const float* __restrict__ M, float* __restrict__ lse, int B, int by_col, float scale) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= B) return;
  float m = -INFINITY;
  for (int c = lane; c < B; c += 32) m = fmaxf(m, scale * (by_col ? M[(int64_t)c * B + r] : M[(int64_t)r * B + c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < B; c += 32) s += expf(scale * (by_col ? M[(int64_t)c * B + r] : M[(int64_t)r * B + c]) - m);
  s = warp_sum(s);
  if (lane == 0) lse[r] = m + logf(s);
}
// bc = pt / temp, cc = tt / temp with pt = preds targs^T and tt = targs targs^T the raw products handed in.
//   loss = 1/2 [ -mean_i sum_j log_softmax(bc)_ij softmax(cc)_ij  -  mean_i sum_j log_softmax(bc^T)_ij softmax(cc)_ij ]
//   dbc_ab = (1 / 2B) [ softmax_row(bc)_ab - softmax_row(cc)_ab + softmax_col(bc)_ab - softmax_row(cc)_ba ]      (x 1/temp: d(p t^T))
__global__ void __launch_bounds__(256) soft_clip_grad_kernel(const float* __restrict__ pt, const float* __restrict__ tt,
                                                             const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                                                             const float* __restrict__ lse_cc, float* __restrict__ dsim,
                                                             double* __restrict__ loss, int B, float inv_temp) {
  __shared__ float red[32];
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (idx < (int64_t)B * B) {
    const int a = (int)(idx / B), b = (int)(idx - (int64_t)a * B);
    const float x = inv_temp * pt[idx];
    const float scc_ab = expf(inv_temp * tt[idx] - lse_cc[a]), scc_ba = expf(inv_temp * tt[(int64_t)b * B + a] - lse_cc[b]);
    const float lr = x - lse_r[a], lc = x - lse_c[b];
    l = -(lr * scc_ab + lc * scc_ba);
    dsim[idx] = (0.5f / B) * inv_temp * ((expf(lr) - scc_ab) + (expf(lc) - scc_ba));
  }
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x < 32) {
    l = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    l = warp_sum(l);
    if (threadIdx.x == 0) atomicAdd(loss, (double)l * (0.5 / B));
  }
}

// ------------------------------------------------------------------------------------------------ AdamW
// torch.optim.AdamW: p *= 1 - lr * wd ; m, v updates ; p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                                                    float bc2, float wd, float grad_scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gg = g[i] * grad_scale;
  const float mm = b1 * m[i] + (1.f - b1) * gg;
  const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
  m[i] = mm;
  v[i] = vv;
  p[i] = p[i] * (1.f - lr * wd) - (lr / bc1) * mm / (sqrtf(vv) * rsqrtf(bc2) + eps);
}

// one launch for a whole parameter list: blockIdx.y walks the table, blockIdx.x grid-strides over that tensor
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AviAdamwEntry* __restrict__ tab, float lr, float b1, float b2, float eps,
                                                          float bc1, float bc2, float grad_scale) {
  const AviAdamwEntry e = tab[blockIdx.y];
  const float decay = 1.f - lr * e.weight_decay, step = lr / bc1, rs2 = rsqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gg = e.g[i] * grad_scale;
    const float mm = b1 * e.m[i] + (1.f - b1) * gg;
    const float vv = b2 * e.v[i] + (1.f - b2) * gg * gg;
    e.m[i] = mm;
    e.v[i] = vv;
    e.p[i] = e.p[i] * decay - step * mm / (sqrtf(vv) * rs2 + eps);
  }
}

}  // namespace avi

using namespace avi;

extern "C" int avi_prior_tokens_fwd(const float* brain, const float* null_brain, const float* keep_brain, const float* x0, const float* noise,
                                    const float* sqrt_ac, const float* sqrt_1mac, const int32_t* times, const float* null_image,
                                    const float* keep_image, const float* learned_query, const float* temb, float* tokens, float* x_noisy,
                                    int32_t B, int32_t dim, void* stream) {
  AVI_REQUIRE(B > 0 && dim == PT_DIM, "avi_prior_tokens_fwd: built for dim 128 (dim=%d)", dim);
  prior_tokens_fwd_kernel<<<B, PT_DIM, 0, (cudaStream_t)stream>>>(brain, null_brain, keep_brain, x0, noise, sqrt_ac, sqrt_1mac, times, null_image,
                                                                  keep_image, learned_query, temb, tokens, x_noisy);
  return check_launch("prior_tokens_fwd");
}

extern "C" int avi_prior_tokens_bwd(const float* dtokens, const float* keep_brain, const float* keep_image, float* dbrain, float* dtemb,
                                    float* dnull_brain, float* dnull_image, float* dlearned_query, int32_t B, int32_t dim, void* stream) {
  AVI_REQUIRE(B > 0 && dim == PT_DIM, "avi_prior_tokens_bwd: built for dim 128 (dim=%d)", dim);
  prior_tokens_bwd_kernel<<<B + 1, PT_DIM, 0, (cudaStream_t)stream>>>(dtokens, keep_brain, keep_image, dbrain, dtemb, dnull_brain, dnull_image,
                                                                      dlearned_query, B);
  return check_launch("prior_tokens_bwd");
}

static bool prior_attn_shape_ok(int32_t n_tokens, int32_t heads, int32_t dim_head) { return n_tokens == PT_N && heads == PT_H && dim_head == PT_D; }

extern "C" int avi_prior_attn_fwd(const float* q, const float* kv, const float* null_kv, const float* rotary, const float* rel_bias, float* out,
                                  float* P, int32_t B, int32_t n_tokens, int32_t heads, int32_t dim_head, void* stream) {
  AVI_REQUIRE(B > 0 && prior_attn_shape_ok(n_tokens, heads, dim_head), "avi_prior_attn_fwd: built for 3 tokens, 8 heads x 64");
  prior_attn_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(q, kv, null_kv, rotary, rel_bias, out, P);
  return check_launch("prior_attn_fwd");
}

extern "C" int avi_prior_attn_bwd(const float* q, const float* kv, const float* null_kv, const float* rotary, const float* P, const float* dout,
                                  float* dq, float* dkv, float* dnull_per_sample, float* dS_per_sample, int32_t B, int32_t n_tokens,
                                  int32_t heads, int32_t dim_head, void* stream) {
  AVI_REQUIRE(B > 0 && prior_attn_shape_ok(n_tokens, heads, dim_head), "avi_prior_attn_bwd: built for 3 tokens, 8 heads x 64");
  prior_attn_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(q, kv, null_kv, rotary, P, dout, dq, dkv, dnull_per_sample, dS_per_sample);
  return check_launch("prior_attn_bwd");
}

extern "C" int avi_swiglu_fwd(const float* h, float* y, int64_t rows, int32_t inner, void* stream) {
  AVI_REQUIRE(rows > 0 && inner > 0, "avi_swiglu_fwd: bad shape");
  swiglu_fwd_kernel<<<(unsigned)((rows * inner + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h, y, rows, inner);
  return check_launch("swiglu_fwd");
}

extern "C" int avi_swiglu_bwd(const float* h, const float* dy, float* dh, int64_t rows, int32_t inner, void* stream) {
  AVI_REQUIRE(rows > 0 && inner > 0, "avi_swiglu_bwd: bad shape");
  swiglu_bwd_kernel<<<(unsigned)((rows * inner + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h, dy, dh, rows, inner);
  return check_launch("swiglu_bwd");
}

extern "C" int avi_rows_stat_div(const float* x, float* out, float* stat, int64_t rows, int32_t C, int32_t mode, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0 && (mode == 0 || mode == 1), "avi_rows_stat_div: bad arguments");
  rows_stat_div_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, out, stat, rows, C, mode);
  return check_launch("rows_stat_div");
}

extern "C" int avi_rows_stat_div_bwd(const float* y, const float* dy, const float* stat, float* dx, int64_t rows, int32_t C, int32_t mode,
                                     void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0 && (mode == 0 || mode == 1), "avi_rows_stat_div_bwd: bad arguments");
  rows_stat_div_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(y, dy, stat, dx, rows, C, mode);
  return check_launch("rows_stat_div_bwd");
}

extern "C" int avi_mul_f32(const float* a, const float* b, float* y, int64_t n, void* stream) {
  AVI_REQUIRE(n > 0, "avi_mul_f32: bad size");
  mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, y, n);
  return check_launch("mul");
}

extern "C" int avi_scale_f32(const float* a, float alpha, float* y, int64_t n, void* stream) {
  AVI_REQUIRE(n > 0, "avi_scale_f32: bad size");
  scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, alpha, y, n);
  return check_launch("scale");
}

extern "C" int avi_soft_clip_loss_grad(const float* pt, const float* tt, float* lse_scratch, float* dsim, double* loss, int32_t B, float temp,
                                       void* stream) {
  AVI_REQUIRE(B > 0 && temp > 0.f, "avi_soft_clip_loss_grad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(loss, 0, sizeof(double), st) != cudaSuccess) {
    set_error("avi_soft_clip_loss_grad: memset failed");
    return 1;
  }
  const unsigned blocks = (unsigned)((B + 7) / 8);
  lse_kernel<<<blocks, 256, 0, st>>>(pt, lse_scratch, B, 0, 1.f / temp);
  lse_kernel<<<blocks, 256, 0, st>>>(pt, lse_scratch + B, B, 1, 1.f / temp);
  lse_kernel<<<blocks, 256, 0, st>>>(tt, lse_scratch + 2 * (size_t)B, B, 0, 1.f / temp);
  soft_clip_grad_kernel<<<(unsigned)(((int64_t)B * B + 255) / 256), 256, 0, st>>>(pt, tt, lse_scratch, lse_scratch + B,
                                                                                 lse_scratch + 2 * (size_t)B, dsim, loss, B, 1.f / temp);
  return check_launch("soft_clip_loss_grad");
}

extern "C" int avi_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int32_t step, float grad_scale, void* stream) {
  AVI_REQUIRE(n > 0 && step >= 1, "avi_adamw_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adamw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2, weight_decay,
                                                                            grad_scale);
  return check_launch("adamw");
}

extern "C" int avi_adamw_multi(const AviAdamwEntry* table_dev, int32_t n_entries, float lr, float beta1, float beta2, float eps, int32_t step,
                               float grad_scale, void* stream) {
  AVI_REQUIRE(table_dev != nullptr && n_entries > 0 && n_entries <= 65535 && step >= 1, "avi_adamw_multi: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adamw_multi_kernel<<<dim3(2 * device_sms(), n_entries), 256, 0, (cudaStream_t)stream>>>(table_dev, lr, beta1, beta2, eps, bc1, bc2, grad_scale);
  return check_launch("adamw_multi");
}
