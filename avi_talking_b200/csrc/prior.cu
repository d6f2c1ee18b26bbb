// Diffusion-prior sampling (text embedding -> style embedding) as ONE persistent kernel for the whole DDPM / DDIM loop.
//
// Reference: models/diffusion_prior.py  VersatileDiffusionPriorNetwork.forward :223-313, FlaggedCausalTransformer.forward
// :154-166, InstructDiffusionPrior.p_sample :329-341 / p_sample_loop_ddpm :344-367, and the un-vendored dalle2_pytorch
// Attention / FeedForward / LayerNorm / RelPosBias / DiffusionPrior.p_sample_loop_ddim they import (:12-18).
// Configuration built at train_diffusion_prior.py:963-991: dim 128, depth 6, 8 heads x 64, one shared K/V head, null
// key/value, cosine-sim attention (scale 16), rotary(32) on q/k, T5 relative-position bias, SwiGLU FF (inner 512), stable
// final LayerNorm, project_out; sequence = [text token, time token, noisy-embedding token + learned query].
//
// The reference issues ~150 tiny kernels per denoising step (launch-latency bound). Here every CTA owns S samples for the
// WHOLE loop: activations never leave shared memory, the 2.06 M fp32 weights of the six layers are streamed from L2 each
// step (all CTAs read the same 8.3 MB, L2-resident), schedule constants and time embeddings are precomputed per step, and the
// caller supplies the noise draws (so the reference's torch.Generator stream can be reproduced exactly).
// All arithmetic is fp32 (the <= 1e-5 mode); one launch replaces steps x ~150 launches.
#include "common.cuh"

namespace avi {

constexpr int PR_DIM = 128, PR_HEADS = 8, PR_DH = 64, PR_INNER = PR_HEADS * PR_DH, PR_FF = 512, PR_NTOK = 3, PR_NKEY = 4;
constexpr int PR_THREADS = 512;

// per-layer packed weights (floats), every matrix TRANSPOSED to [in][out]
constexpr int PR_OFF_ATT_G = 0;                                   // [128]
constexpr int PR_OFF_NULLK = PR_OFF_ATT_G + PR_DIM;               // [64]
constexpr int PR_OFF_NULLV = PR_OFF_NULLK + PR_DH;                // [64]
constexpr int PR_OFF_WQ = PR_OFF_NULLV + PR_DH;                   // [128][512]
constexpr int PR_OFF_WKV = PR_OFF_WQ + PR_DIM * PR_INNER;         // [128][128]
constexpr int PR_OFF_WO = PR_OFF_WKV + PR_DIM * 2 * PR_DH;        // [512][128]
constexpr int PR_OFF_OUT_G = PR_OFF_WO + PR_INNER * PR_DIM;       // [128]
constexpr int PR_OFF_FF_G = PR_OFF_OUT_G + PR_DIM;                // [128]
constexpr int PR_OFF_W1 = PR_OFF_FF_G + PR_DIM;                   // [128][1024]
constexpr int PR_OFF_W2 = PR_OFF_W1 + PR_DIM * 2 * PR_FF;         // [512][128]
constexpr int PR_LAYER_FLOATS = PR_OFF_W2 + PR_FF * PR_DIM;

struct PriorParams {
  const float* layers;      // [depth][PR_LAYER_FLOATS]
  const float* learned_q;   // [128]
  const float* rel_bias;    // [heads][3][4]
  const float* rot;         // [3][16][2] cos, sin of position * freq
  const float* norm_g;      // [128] final (stable) LayerNorm gain
  const float* proj_t;      // [128][128] project_out transposed
  const float* temb;        // [steps][128] time-token embeddings
  const float* sched;       // [steps][6]: mode, p0..p4
  const float* text;        // [B][128]
  const float* x_init;      // [B][128]
  const float* noise;       // [steps][B][128]
  float* out;               // [B][128]
  int B, steps, depth;
  float out_scale;          // result multiplied by this (1 / image_embed_scale)
  const float* null_pred;   // [steps][128] or nullptr: the denoiser's output with BOTH conditions replaced by the null embeddings
  float cond_scale;         // classifier-free guidance: x0 = null + (x0 - null) * cond_scale (forward_with_cond_scale :209-221)
};

// out[r][n] = sum_k inT[k][r] * Wt[k][n]   (r < R rows, all in shared memory except Wt)
// V columns per thread; when N / V < PR_THREADS the K range is split over thread groups and reduced through `partial`.
template <int R, int V>
__device__ __forceinline__ void matmul_rows(const float* __restrict__ Wt, int K, int N, const float* inT, float* out, float* partial) {
  const int tpk = N / V;                 // threads per k-row
  const int ksplit = PR_THREADS / tpk;   // >= 1
  const int cg = threadIdx.x % tpk, ks = threadIdx.x / tpk;
  const int kchunk = K / ksplit;
  const int k0 = ks * kchunk;
  float acc[R][V];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[r][v] = 0.f;
  const float* wp = Wt + (int64_t)k0 * N + cg * V;
  // The weights come from L2 (every CTA streams the same 8 MB per denoising step) and each thread walks its own column: with a
  // handful of loads in flight per thread the loop ran at L2 LATENCY (~51 GB/s per SM). Issue PR_WB independent loads first, then
  // consume them: 16 x 512 threads x 4-8 bytes in flight per SM.
  constexpr int PR_WB = 16;
  for (int kb = 0; kb < kchunk; kb += PR_WB) {
    float w[PR_WB][V];
#pragma unroll
    for (int u = 0; u < PR_WB; ++u) {
      if (kb + u < kchunk) {
        if constexpr (V == 2) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(wp + (int64_t)(kb + u) * N));
          w[u][0] = t.x;
          w[u][1] = t.y;
        } else {
          w[u][0] = __ldg(wp + (int64_t)(kb + u) * N);
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) w[u][v] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < PR_WB; ++u) {
      if (kb + u < kchunk) {
        const float* xr = inT + (k0 + kb + u) * R;
        float x[R];
        if constexpr (R % 4 == 0) {
#pragma unroll
          for (int q = 0; q < R / 4; ++q) {
            const float4 t = reinterpret_cast<const float4*>(xr)[q];
            x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) x[r] = xr[r];
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int v = 0; v < V; ++v) acc[r][v] = fmaf(x[r], w[u][v], acc[r][v]);
      }
    }
  }
  if (ksplit == 1) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) out[r * N + cg * V + v] = acc[r][v];
    __syncthreads();
    return;
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int v = 0; v < V; ++v) partial[(ks * R + r) * N + cg * V + v] = acc[r][v];
  __syncthreads();
  for (int i = threadIdx.x; i < R * N; i += PR_THREADS) {
    float s = 0.f;
    for (int q = 0; q < ksplit; ++q) s += partial[q * R * N + i];
    out[i] = s;
  }
  __syncthreads();
}

// dalle2_pytorch.LayerNorm over rows of x [R][128] (gain only, biased variance, eps 1e-5; `stable` divides by the row max first)
// result written TRANSPOSED to outT[k][R] (+ optionally row-major to out_rm). One warp per row.
template <int R>
__device__ __forceinline__ void ln_rows_T(const float* x, const float* __restrict__ g, float* outT, float* out_rm, bool stable) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += PR_THREADS / 32) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = x[r * PR_DIM + lane + 32 * u];
    if (stable) {
      const float m = warp_max(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])));
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = v[u] / m;
    }
    const float mean = warp_sum((v[0] + v[1]) + (v[2] + v[3])) * (1.f / PR_DIM);
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) q += (v[u] - mean) * (v[u] - mean);
    const float rstd = rsqrtf(warp_sum(q) * (1.f / PR_DIM) + 1e-5f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = lane + 32 * u;
      const float y = (v[u] - mean) * rstd * __ldg(g + c);
      outT[c * R + r] = y;
      if (out_rm) out_rm[r * PR_DIM + c] = y;
    }
  }
  __syncthreads();
}

template <int S>
__global__ void __launch_bounds__(PR_THREADS, 1) prior_sample_kernel(const PriorParams p) {
  constexpr int R = PR_NTOK * S;
  extern __shared__ float sm[];
  float* xtok = sm;                          // [R][128]      residual stream
  float* inT = xtok + R * PR_DIM;            // [512][R]      transposed input of the current matmul
  float* buf = inT + PR_FF * R;              // [R][1024]     matmul output (q | ff hidden | ...)
  float* kv = buf + R * 2 * PR_FF;           // [R][128]      k | v of the shared head
  float* khat = kv + R * 2 * PR_DH;          // [S][4][64]    rotated, normalised, scaled keys (incl. null key)
  float* vall = khat + S * PR_NKEY * PR_DH;  // [S][4][64]
  float* xcur = vall + S * PR_NKEY * PR_DH;  // [S][128]      current noisy embedding
  float* partial = xcur + S * PR_DIM;        // K-split partial sums: up to [2][R][512] = [4][R][128] ... sized on the host
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s0 = blockIdx.x * S;

  for (int i = threadIdx.x; i < S * PR_DIM; i += PR_THREADS) {
    const int s = i / PR_DIM, c = i % PR_DIM;
    xcur[i] = (s0 + s < p.B) ? p.x_init[(int64_t)(s0 + s) * PR_DIM + c] : 0.f;
  }
  __syncthreads();

  for (int step = 0; step < p.steps; ++step) {
    // ---- tokens: [text, time, x + learned_query]   (diffusion_prior.py:283-304)
    for (int i = threadIdx.x; i < S * PR_DIM; i += PR_THREADS) {
      const int s = i / PR_DIM, c = i % PR_DIM;
      const bool ok = s0 + s < p.B;
      xtok[(3 * s + 0) * PR_DIM + c] = ok ? __ldg(p.text + (int64_t)(s0 + s) * PR_DIM + c) : 0.f;
      xtok[(3 * s + 1) * PR_DIM + c] = __ldg(p.temb + step * PR_DIM + c);
      xtok[(3 * s + 2) * PR_DIM + c] = xcur[i] + __ldg(p.learned_q + c);
    }
    __syncthreads();

    for (int l = 0; l < p.depth; ++l) {
      const float* W = p.layers + (int64_t)l * PR_LAYER_FLOATS;
      // ---- attention (dalle2_pytorch.Attention.forward)
      ln_rows_T<R>(xtok, W + PR_OFF_ATT_G, inT, nullptr, false);
      matmul_rows<R, 2>(W + PR_OFF_WQ, PR_DIM, PR_INNER, inT, buf, partial);      // q   [R][512]
      matmul_rows<R, 1>(W + PR_OFF_WKV, PR_DIM, 2 * PR_DH, inT, kv, partial);      // k|v [R][128]
      // keys / values per sample: null first, then the 3 tokens; rotary (first 32 dims, interleaved pairs), l2norm, * sqrt(16)
      for (int item = warp; item < S * PR_NKEY; item += PR_THREADS / 32) {
        const int s = item / PR_NKEY, j = item % PR_NKEY;
        float k0, k1, v0, v1;
        if (j == 0) {
          k0 = __ldg(W + PR_OFF_NULLK + 2 * lane);
          k1 = __ldg(W + PR_OFF_NULLK + 2 * lane + 1);
          v0 = __ldg(W + PR_OFF_NULLV + 2 * lane);
          v1 = __ldg(W + PR_OFF_NULLV + 2 * lane + 1);
        } else {
          const float* row = kv + (3 * s + j - 1) * 2 * PR_DH;
          k0 = row[2 * lane];
          k1 = row[2 * lane + 1];
          v0 = row[PR_DH + 2 * lane];
          v1 = row[PR_DH + 2 * lane + 1];
          if (lane < 16) {
            const float c = __ldg(p.rot + ((j - 1) * 16 + lane) * 2), sn = __ldg(p.rot + ((j - 1) * 16 + lane) * 2 + 1);
            const float a = k0 * c - k1 * sn, b = k1 * c + k0 * sn;
            k0 = a;
            k1 = b;
          }
        }
        const float nrm = fmaxf(sqrtf(warp_sum(k0 * k0 + k1 * k1)), 1e-12f);
        khat[(s * PR_NKEY + j) * PR_DH + 2 * lane] = k0 / nrm * 4.f;
        khat[(s * PR_NKEY + j) * PR_DH + 2 * lane + 1] = k1 / nrm * 4.f;
        vall[(s * PR_NKEY + j) * PR_DH + 2 * lane] = v0;
        vall[(s * PR_NKEY + j) * PR_DH + 2 * lane + 1] = v1;
      }
      __syncthreads();
      // one warp per (sample, head, query token): cosine-sim scores against the 4 keys + T5 bias, softmax, weighted values
      for (int item = warp; item < S * PR_HEADS * PR_NTOK; item += PR_THREADS / 32) {
        const int s = item / (PR_HEADS * PR_NTOK), h = (item / PR_NTOK) % PR_HEADS, i = item % PR_NTOK;
        const int r = 3 * s + i;
        float q0 = buf[r * PR_INNER + h * PR_DH + 2 * lane] * 16.f, q1 = buf[r * PR_INNER + h * PR_DH + 2 * lane + 1] * 16.f;
        if (lane < 16) {
          const float c = __ldg(p.rot + (i * 16 + lane) * 2), sn = __ldg(p.rot + (i * 16 + lane) * 2 + 1);
          const float a = q0 * c - q1 * sn, b = q1 * c + q0 * sn;
          q0 = a;
          q1 = b;
        }
        const float nrm = fmaxf(sqrtf(warp_sum(q0 * q0 + q1 * q1)), 1e-12f);
        q0 = q0 / nrm * 4.f;
        q1 = q1 / nrm * 4.f;
        float sim[PR_NKEY];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < PR_NKEY; ++j) {
          const float* kk = khat + (s * PR_NKEY + j) * PR_DH;
          sim[j] = warp_sum(q0 * kk[2 * lane] + q1 * kk[2 * lane + 1]) + __ldg(p.rel_bias + (h * PR_NTOK + i) * PR_NKEY + j);
          mx = fmaxf(mx, sim[j]);
        }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < PR_NKEY; ++j) {
          sim[j] = expf(sim[j] - mx);
          den += sim[j];
        }
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int j = 0; j < PR_NKEY; ++j) {
          const float pj = sim[j] / den;
          const float* vv = vall + (s * PR_NKEY + j) * PR_DH;
          o0 = fmaf(pj, vv[2 * lane], o0);
          o1 = fmaf(pj, vv[2 * lane + 1], o1);
        }
        inT[(h * PR_DH + 2 * lane) * R + r] = o0;
        inT[(h * PR_DH + 2 * lane + 1) * R + r] = o1;
      }
      __syncthreads();
      matmul_rows<R, 1>(W + PR_OFF_WO, PR_INNER, PR_DIM, inT, buf, partial);       // to_out.0  [R][128]
      {  // to_out.1 LayerNorm, residual add
        for (int r = warp; r < R; r += PR_THREADS / 32) {
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = buf[r * PR_DIM + lane + 32 * u];
          const float mean = warp_sum((v[0] + v[1]) + (v[2] + v[3])) * (1.f / PR_DIM);
          float q = 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) q += (v[u] - mean) * (v[u] - mean);
          const float rstd = rsqrtf(warp_sum(q) * (1.f / PR_DIM) + 1e-5f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = lane + 32 * u;
            xtok[r * PR_DIM + c] += (v[u] - mean) * rstd * __ldg(W + PR_OFF_OUT_G + c);
          }
        }
        __syncthreads();
      }
      // ---- feed-forward (dalle2_pytorch.FeedForward: LayerNorm, Linear(128 -> 1024), SwiGLU, Linear(512 -> 128))
      ln_rows_T<R>(xtok, W + PR_OFF_FF_G, inT, nullptr, false);
      matmul_rows<R, 2>(W + PR_OFF_W1, PR_DIM, 2 * PR_FF, inT, buf, partial);
      for (int i = threadIdx.x; i < R * PR_FF; i += PR_THREADS) {
        const int r = i / PR_FF, c = i % PR_FF;
        const float a = buf[r * 2 * PR_FF + c], gte = buf[r * 2 * PR_FF + PR_FF + c];
        inT[c * R + r] = a * (gte / (1.f + expf(-gte)));
      }
      __syncthreads();
      matmul_rows<R, 1>(W + PR_OFF_W2, PR_FF, PR_DIM, inT, buf, partial);
      for (int i = threadIdx.x; i < R * PR_DIM; i += PR_THREADS) xtok[i] += buf[i];
      __syncthreads();
    }
    // ---- stable LayerNorm + project_out; the last token of every sample is the x0 prediction (:165-166, :311)
    ln_rows_T<R>(xtok, p.norm_g, inT, nullptr, true);
    matmul_rows<R, 1>(p.proj_t, PR_DIM, PR_DIM, inT, buf, partial);
    // ---- DDPM / DDIM update (p_sample :329-341, q_posterior; dalle2 p_sample_loop_ddim)
    const float* sc = p.sched + step * 6;
    const int mode = (int)__ldg(sc);
    const float p0 = __ldg(sc + 1), p1 = __ldg(sc + 2), p2 = __ldg(sc + 3), p3 = __ldg(sc + 4), p4 = __ldg(sc + 5);
    for (int i = threadIdx.x; i < S * PR_DIM; i += PR_THREADS) {
      const int s = i / PR_DIM, c = i % PR_DIM;
      if (s0 + s >= p.B) continue;
      float x0 = buf[(3 * s + 2) * PR_DIM + c];
      if (p.null_pred != nullptr) {   // the null pass sees neither the text nor the noisy embedding: one vector per step, precomputed
        const float nl = __ldg(p.null_pred + (int64_t)step * PR_DIM + c);
        x0 = nl + (x0 - nl) * p.cond_scale;
      }
      const float x = xcur[i];
      const float nz = __ldg(p.noise + ((int64_t)step * p.B + s0 + s) * PR_DIM + c);
      float xn;
      if (mode == 0) {            // DDPM: mean = coef1 * x0 + coef2 * x ; x = mean + sigma * noise
        xn = (p0 * x0 + p1 * x) + p2 * nz;
      } else if (mode == 1) {     // DDIM: eps = (sqrt_recip * x - x0) / sqrt_recipm1 ; x = x0 * sqrt(a_next) + c1 * noise + c2 * eps
        const float eps = (p0 * x - x0) / p1;
        xn = (x0 * p2 + p3 * nz) + p4 * eps;
      } else {                    // last DDIM pair (time_next < 0): x = x0
        xn = x0;
      }
      xcur[i] = xn;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < S * PR_DIM; i += PR_THREADS) {
    const int s = i / PR_DIM, c = i % PR_DIM;
    if (s0 + s < p.B) p.out[(int64_t)(s0 + s) * PR_DIM + c] = xcur[i] * p.out_scale;
  }
}

// time-token embeddings for every step: SinusoidalPosEmb(128) -> MLP(128 -> 256 -> 256 -> 128, SiLU)   (dalle2_pytorch)
__global__ void __launch_bounds__(256) prior_time_embed_kernel(const float* __restrict__ times, const float* __restrict__ w0t,
                                                                const float* __restrict__ b0, const float* __restrict__ w1t,
                                                                const float* __restrict__ b1, const float* __restrict__ w2t,
                                                                const float* __restrict__ b2, float* __restrict__ temb) {
  __shared__ float e[PR_DIM], h0[2 * PR_DIM], h1[2 * PR_DIM];
  const int step = blockIdx.x, tid = threadIdx.x;
  const float t = times[step];
  if (tid < PR_DIM) {
    const int half = PR_DIM / 2;
    const int j = tid % half;
    const float f = expf((float)j * -(logf(10000.f) / (float)(half - 1)));
    e[tid] = tid < half ? sinf(t * f) : cosf(t * f);
  }
  __syncthreads();
  {
    float a = b0[tid];
    for (int k = 0; k < PR_DIM; ++k) a = fmaf(e[k], w0t[k * 2 * PR_DIM + tid], a);
    h0[tid] = a / (1.f + expf(-a));
  }
  __syncthreads();
  {
    float a = b1[tid];
    for (int k = 0; k < 2 * PR_DIM; ++k) a = fmaf(h0[k], w1t[k * 2 * PR_DIM + tid], a);
    h1[tid] = a / (1.f + expf(-a));
  }
  __syncthreads();
  if (tid < PR_DIM) {
    float a = b2[tid];
    for (int k = 0; k < 2 * PR_DIM; ++k) a = fmaf(h1[k], w2t[k * PR_DIM + tid], a);
    temb[step * PR_DIM + tid] = a;
  }
}

// y = GELU(LayerNorm(x)) (+ res): BrainNetwork blocks (models/diffusion_prior.py:63-75,104-110: Linear -> LayerNorm -> GELU
// -> dropout(eval) ; x += residual) and its projector (:83-93). One block per row, C <= 4096.
__global__ void __launch_bounds__(256) ln_gelu_res_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, const float* __restrict__ res,
                                                           float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int C,
                                                           float eps) {
  __shared__ float red[64];
  const int64_t row = blockIdx.x;
  const float* xr = x + row * C;
  float v[16];
  float s = 0.f, dummy = 0.f;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = threadIdx.x + 256 * u;
    v[u] = c < C ? xr[c] : 0.f;
    s += v[u];
  }
  block_sum2(s, dummy, red);
  const float mean = s / C;
  float q = 0.f;
  dummy = 0.f;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = threadIdx.x + 256 * u;
    if (c < C) q += (v[u] - mean) * (v[u] - mean);
  }
  block_sum2(q, dummy, red);
  const float rstd = rsqrtf(q / C + eps);
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = threadIdx.x + 256 * u;
    if (c < C) {
      float y = gelu_erf((v[u] - mean) * rstd * w[c] + b[c]);
      if (res) y += res[row * C + c];
      if (out_f32) out_f32[row * C + c] = y;
      if (out_bf16) out_bf16[row * C + c] = __float2bfloat16_rn(y);
    }
  }
}

template <int S>
static int launch_prior(const PriorParams& p, cudaStream_t st) {
  constexpr int R = PR_NTOK * S;
  const size_t floats = (size_t)R * PR_DIM + (size_t)PR_FF * R + (size_t)R * 2 * PR_FF + (size_t)R * 2 * PR_DH +
                        2 * (size_t)S * PR_NKEY * PR_DH + (size_t)S * PR_DIM + (size_t)2 * R * PR_INNER;
  const size_t bytes = floats * sizeof(float);
  static SmemOptIn optin;
  const cudaError_t attr_err = smem_optin(prior_sample_kernel<S>, (int)bytes, optin);
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_prior_sample: cannot opt in to %zu bytes of shared memory: %s", bytes,
              cudaGetErrorString(attr_err));
  prior_sample_kernel<S><<<(p.B + S - 1) / S, PR_THREADS, bytes, st>>>(p);
  return check_launch("prior_sample");
}

}  // namespace avi

using namespace avi;

extern "C" int avi_prior_layer_floats(void) { return PR_LAYER_FLOATS; }

extern "C" int avi_prior_time_embed(const float* times, const float* w0t, const float* b0, const float* w1t, const float* b1,
                                    const float* w2t, const float* b2, float* temb, int32_t steps, void* stream) {
  AVI_REQUIRE(steps > 0, "avi_prior_time_embed: steps must be positive");
  prior_time_embed_kernel<<<steps, 256, 0, (cudaStream_t)stream>>>(times, w0t, b0, w1t, b1, w2t, b2, temb);
  return check_launch("prior_time_embed");
}

extern "C" int avi_prior_sample(const AviPriorNet* net, const float* temb, const float* sched, const float* text_embed,
                                const float* x_init, const float* noise, float* out, int32_t B, int32_t steps, float out_scale,
                                int32_t samples_per_cta, void* stream) {
  return avi_prior_sample_cfg(net, temb, sched, text_embed, x_init, noise, nullptr, 1.f, out, B, steps, out_scale, samples_per_cta, stream);
}

extern "C" int avi_prior_sample_cfg(const AviPriorNet* net, const float* temb, const float* sched, const float* text_embed,
                                    const float* x_init, const float* noise, const float* null_pred, float cond_scale, float* out,
                                    int32_t B, int32_t steps, float out_scale, int32_t samples_per_cta, void* stream) {
  AVI_REQUIRE(net != nullptr && net->dim == PR_DIM && net->heads == PR_HEADS && net->dim_head == PR_DH && net->ff_inner == PR_FF,
              "avi_prior_sample: only the reference configuration (dim 128, 8 heads x 64, ff inner 512) is built");
  AVI_REQUIRE(B > 0 && steps > 0 && net->depth > 0, "avi_prior_sample: bad sizes");
  PriorParams p;
  p.layers = net->layers;
  p.learned_q = net->learned_query;
  p.rel_bias = net->rel_bias;
  p.rot = net->rotary;
  p.norm_g = net->norm_g;
  p.proj_t = net->project_out_t;
  p.temb = temb;
  p.sched = sched;
  p.text = text_embed;
  p.x_init = x_init;
  p.noise = noise;
  p.out = out;
  p.B = B;
  p.steps = steps;
  p.depth = net->depth;
  p.out_scale = out_scale;
  p.null_pred = null_pred;
  p.cond_scale = cond_scale;
  int S = samples_per_cta;
  const int half_sms = device_sms() / 2;
  if (S <= 0) S = (B + 3) / 4 >= half_sms ? 4 : ((B + 1) / 2 >= half_sms ? 2 : 1);  // fill the SMs before batching per CTA
  if (S >= 4) return launch_prior<4>(p, (cudaStream_t)stream);
  if (S >= 2) return launch_prior<2>(p, (cudaStream_t)stream);
  return launch_prior<1>(p, (cudaStream_t)stream);
}

extern "C" int avi_ln_gelu_res(const float* x, const float* w, const float* b, const float* res, float* out_f32, void* out_bf16,
                               int64_t rows, int32_t C, float eps, void* stream) {
  AVI_REQUIRE(C > 0 && C <= 4096 && rows > 0, "avi_ln_gelu_res: C must be in 1..4096");
  ln_gelu_res_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, w, b, res, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                                       C, eps);
  return check_launch("ln_gelu_res");
}
