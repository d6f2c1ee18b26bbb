// wav2vec2 layer 0 on tensor cores: Conv1d(1 -> 512, k = 10, s = 5, no bias) + GroupNorm(512 groups) + GELU, bf16 time-major output.
// (HF Wav2Vec2GroupNormConvLayer, reached from models/lib/wav2vec.py:97)
//
// The CUDA-core version spent 1.8 ms per 64 x 10 s batch issue-bound on 10 FMAs + 5 LDS per output (profiles/r1); the output
// is 2.1 GB of bf16, i.e. 0.32 ms at HBM speed. Here the 10-tap contraction is an MMA:
//   A tile  : 128 time steps x K=32 bf16, im2col rows [hi(x0..x9) | lo(x0..x9) | hi(x0..x9) | 0 0] built in shared memory by four
//             producer warps directly in the SWIZZLE_128B K-major layout (the row stride of 5 samples = 20 bytes rules TMA out)
//   W       : [512 x 32] bf16 rows [hi(w) | hi(w) | lo(w) | 0 0], resident in shared memory (split-bf16: the product equals the
//             fp32 convolution to ~2^-16 relative; only lo*lo is dropped)
//   D       : four TMEM accumulators of 128 lanes x 128 columns (one per 128-channel group), each released as soon as its epilogue
//             warps have drained it, so the MMAs of tile i+1 overlap the epilogue of tile i
//   epilogue: 16 warps; y = acc * scale[b,c] + shift[b,c] (GroupNorm folded into one FMA), branch-free GELU, bf16, written row-per-lane
//             into a double-buffered SWIZZLE_64B staging tile that one lane hands to the TMA unit (no shared loads, no global store
//             instructions: the LSU pipe carries the st.shared traffic only).
// GroupNorm statistics are exact and cost one pass over the AUDIO only: y is linear in the 10-sample window, so
//   sum_t y = w . S1,  sum_t y^2 = w^T R w   with  S1[j] = sum_t x[5t+j],  R[j][j'] = sum_t x[5t+j] x[5t+j']   (65 moments / clip, fp64).
#include "tc_common.cuh"

namespace avi {

constexpr int CZ_K = 10, CZ_S = 5, CZ_BM = 128, CZ_C = 512, CZ_NQ = 4, CZ_BN = 128;
constexpr int CZ_BUILD_WARPS = 4, CZ_EPI_WARPS = 16, CZ_THREADS = (CZ_BUILD_WARPS + 1 + CZ_EPI_WARPS) * 32;  // 672
constexpr int CZ_NMOM = 65, CZ_MOM_STRIDE = 72;
constexpr uint32_t CZ_A_BYTES = CZ_BM * 128;            // 16 KB per stage (128-byte rows, first 64 bytes used)
constexpr uint32_t CZ_W_BYTES = CZ_C * 128;             // 64 KB
constexpr uint32_t CZ_OFF_W = 2 * CZ_A_BYTES;
constexpr uint32_t CZ_OFF_X = CZ_OFF_W + CZ_W_BYTES;    // audio staging: 2 stages x 656 floats
constexpr uint32_t CZ_X_FLOATS = 656;
constexpr uint32_t CZ_OFF_SS = CZ_OFF_X + 2 * CZ_X_FLOATS * 4;   // scale | shift of the current clip: 2 x 512 floats, 2 stages
// epilogue staging: two 2 KB tiles per warp (32 rows x 64 bytes, SWIZZLE_64B boxes of the output tensor map), 1 KB aligned so that
// the hardware swizzle (address bits [7,9) into bits [4,6)) is the (row >> 1) & 3 pattern the writers use
constexpr uint32_t CZ_OFF_TRANS = (CZ_OFF_SS + 2 * 2 * CZ_C * 4 + 1023) / 1024 * 1024;
constexpr uint32_t CZ_TRANS_TILE = 32 * 16 * 4;
constexpr uint32_t CZ_TRANS_WARP = 2 * CZ_TRANS_TILE;
constexpr uint32_t CZ_OFF_BAR = CZ_OFF_TRANS + CZ_EPI_WARPS * CZ_TRANS_WARP;
constexpr int CZ_CLC_STAGES = 4;                         // cluster-launch-control responses (dynamic tile scheduler, see gemm_tc2.cu)
constexpr uint32_t CZ_OFF_CLC = CZ_OFF_BAR + 256;        // [CZ_CLC_STAGES] 16-byte responses
constexpr uint32_t CZ_SMEM = CZ_OFF_CLC + CZ_CLC_STAGES * 16;
constexpr int CZ_CLC_CONSUMERS = CZ_BUILD_WARPS + 1 + CZ_EPI_WARPS;   // one arrival per builder warp, the MMA thread, every epilogue warp
static_assert(CZ_SMEM <= 232448, "shared memory budget");
static_assert(CZ_OFF_X % 16 == 0 && CZ_OFF_SS % 16 == 0 && CZ_OFF_TRANS % 1024 == 0 && CZ_OFF_BAR % 8 == 0, "alignment");

// ---------------------------------------------------------------------------------------------- GroupNorm statistics
__global__ void __launch_bounds__(256) conv0_moments_kernel(const float* __restrict__ audio, double* __restrict__ mom, int n_samples,
                                                            int L0, int t_per_block) {
  __shared__ double red[8][CZ_NMOM];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * t_per_block, t1 = min(L0, t0 + t_per_block);
  const float* x = audio + (int64_t)b * n_samples;
  float acc[CZ_NMOM];
#pragma unroll
  for (int i = 0; i < CZ_NMOM; ++i) acc[i] = 0.f;
  for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {  // <= 16 terms per thread in fp32, fp64 across threads / blocks
    float v[CZ_K];
#pragma unroll
    for (int j = 0; j < CZ_K; ++j) v[j] = x[(int64_t)t * CZ_S + j];
    int m = 0;
#pragma unroll
    for (int j = 0; j < CZ_K; ++j) acc[m++] += v[j];
#pragma unroll
    for (int j = 0; j < CZ_K; ++j)
#pragma unroll
      for (int k = j; k < CZ_K; ++k) {
        acc[m] = fmaf(v[j], v[k], acc[m]);
        ++m;
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < CZ_NMOM; ++i) {
    double d = (double)acc[i];
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) red[warp][i] = d;
  }
  __syncthreads();
  if (threadIdx.x < CZ_NMOM) {
    double d = 0.0;
    for (int w = 0; w < 8; ++w) d += red[w][threadIdx.x];
    atomicAdd(&mom[(int64_t)b * CZ_MOM_STRIDE + threadIdx.x], d);
  }
}

__global__ void __launch_bounds__(CZ_C) conv0_gn_coeffs_kernel(const double* __restrict__ mom, const float* __restrict__ w,
                                                                const float* __restrict__ gn_w, const float* __restrict__ gn_b,
                                                                float* __restrict__ scale, float* __restrict__ shift, int L0, float eps) {
  __shared__ double m[CZ_NMOM];
  const int b = blockIdx.x, c = threadIdx.x;
  if (c < CZ_NMOM) m[c] = mom[(int64_t)b * CZ_MOM_STRIDE + c];
  __syncthreads();
  double wv[CZ_K];
#pragma unroll
  for (int j = 0; j < CZ_K; ++j) wv[j] = (double)w[c * CZ_K + j];
  double s = 0.0, ss = 0.0;
  int idx = CZ_K;
#pragma unroll
  for (int j = 0; j < CZ_K; ++j) s += wv[j] * m[j];
#pragma unroll
  for (int j = 0; j < CZ_K; ++j)
#pragma unroll
    for (int k = j; k < CZ_K; ++k) ss += (j == k ? 1.0 : 2.0) * wv[j] * wv[k] * m[idx++];
  const double mean = s / L0;
  const double var = fmax(ss / L0 - mean * mean, 0.0);
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = rstd * gn_w[c];
  scale[(int64_t)b * CZ_C + c] = sc;
  shift[(int64_t)b * CZ_C + c] = gn_b[c] - (float)mean * sc;
}

// wp[c][0..63] bf16: [hi(w0..w9) | hi(w0..w9) | lo(w0..w9) | zeros]   (K-major rows of 128 bytes; TMA applies the swizzle)
__global__ void conv0_pack_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CZ_C * 64) return;
  const int c = i / 64, k = i % 64;
  float v = 0.f;
  if (k < 30) {
    const float x = w[c * CZ_K + (k % 10)];
    const float hi = __bfloat162float(__float2bfloat16_rn(x));
    v = k < 20 ? hi : x - hi;
  }
  wp[i] = __float2bfloat16_rn(v);
}

struct Conv0Params {
  const float* audio;
  const float* scale;   // [B][512]
  const float* shift;
  __nv_bfloat16* out;
  int64_t out_batch_stride;
  int n_samples, L0, tiles_per_clip, total_tiles;
  int dynamic;   // 1: one CTA per tile in the grid, the resident CTAs take the pending ones through cluster launch control
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(CZ_THREADS, 1) conv0_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                                                                 const Conv0Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CZ_OFF_BAR);
  uint64_t* w_full = bars;              // [1]
  uint64_t* a_full = bars + 1;          // [2] count = builder warps
  uint64_t* a_empty = bars + 3;         // [2] tcgen05.commit
  uint64_t* tmem_full = bars + 5;       // [4] tcgen05.commit
  uint64_t* tmem_empty = bars + 9;      // [4] count = 4 epilogue warps
  uint64_t* ss_full = bars + 13;        // [2] count = builder warps (scale/shift of the tile's clip staged)
  uint64_t* ss_empty = bars + 15;       // [2] count = epilogue warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);
  uint64_t* clc_full = bars + 18;       // [CLC_STAGES]
  uint64_t* clc_empty = bars + 22;      // [CLC_STAGES]
  uint8_t* clc_resp = smem + CZ_OFF_CLC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dyn = p.dynamic != 0;
  // next tile of this CTA (lane 0 of a warp asks, the warp shares the answer): the static walk or a tile taken from the launch queue
  auto next_tile = [&](int t, int& cslot, uint32_t& cphase) -> int {
    int nt = 0;
    if (lane == 0) {
      if (!dyn) {
        nt = t + (int)gridDim.x;
        if (nt >= p.total_tiles) nt = -1;
      } else {
        mbar_wait(smem_u32(&clc_full[cslot]), cphase);
        nt = clc_decode(smem_u32(clc_resp + 16 * cslot));
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&clc_empty[cslot]));
      }
    }
    if (++cslot == CZ_CLC_STAGES) {
      cslot = 0;
      cphase ^= 1;
    }
    return __shfl_sync(0xffffffffu, nt, 0);
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    mbar_init(smem_u32(w_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&a_full[s]), CZ_BUILD_WARPS);
      mbar_init(smem_u32(&a_empty[s]), 1);
      mbar_init(smem_u32(&ss_full[s]), CZ_BUILD_WARPS);
      mbar_init(smem_u32(&ss_empty[s]), CZ_EPI_WARPS);
    }
    for (int q = 0; q < CZ_NQ; ++q) {
      mbar_init(smem_u32(&tmem_full[q]), 1);
      mbar_init(smem_u32(&tmem_empty[q]), CZ_EPI_WARPS / CZ_NQ);
    }
    for (int s = 0; s < CZ_CLC_STAGES; ++s) {
      mbar_init(smem_u32(&clc_full[s]), 1);
      mbar_init(smem_u32(&clc_empty[s]), CZ_CLC_CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CZ_BUILD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < CZ_BUILD_WARPS) {
    // ===================== producers: W once (TMA), then one im2col A tile + the clip's scale/shift per tile =====================
    if (warp == 0 && lane == 0) {
      mbar_expect_tx(smem_u32(w_full), CZ_W_BYTES);
      for (int q = 0; q < CZ_NQ; ++q) tma_load_2d(smem_u32(smem + CZ_OFF_W + q * (CZ_BN * 128)), &map_w, smem_u32(w_full), 0, q * CZ_BN);
    }
    const int r = threadIdx.x;  // row of the tile, 0..127
    int it = 0, cslot = 0, islot = 0;
    uint32_t cphase = 0, iphase = 0;
    auto clc_issue = [&]() {     // thread 0: ask for the tile after the one about to be built (never after a failed request)
      mbar_wait(smem_u32(&clc_empty[islot]), iphase ^ 1);
      mbar_expect_tx(smem_u32(&clc_full[islot]), 16);
      clc_try_cancel(smem_u32(clc_resp + 16 * islot), smem_u32(&clc_full[islot]));
      if (++islot == CZ_CLC_STAGES) {
        islot = 0;
        iphase ^= 1;
      }
    };
    if (dyn && threadIdx.x == 0) clc_issue();
    for (int t = blockIdx.x; t >= 0; ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int b = t / p.tiles_per_clip, t0 = (t % p.tiles_per_clip) * CZ_BM;
      const int nrows = min(CZ_BM, p.L0 - t0);
      const int need = (nrows - 1) * CZ_S + CZ_K;
      mbar_wait(smem_u32(&a_empty[s]), ph ^ 1);
      mbar_wait(smem_u32(&ss_empty[s]), ph ^ 1);
      float* xs = reinterpret_cast<float*>(smem + CZ_OFF_X) + s * CZ_X_FLOATS;
      const float* xg = p.audio + (int64_t)b * p.n_samples + (int64_t)t0 * CZ_S;
      for (int i = threadIdx.x; i < (int)CZ_X_FLOATS; i += CZ_BUILD_WARPS * 32) xs[i] = i < need ? __ldg(xg + i) : 0.f;
      float* ssd = reinterpret_cast<float*>(smem + CZ_OFF_SS) + s * (2 * CZ_C);
      for (int i = threadIdx.x; i < CZ_C; i += CZ_BUILD_WARPS * 32) {
        ssd[i] = __ldg(p.scale + (int64_t)b * CZ_C + i);
        ssd[CZ_C + i] = __ldg(p.shift + (int64_t)b * CZ_C + i);
      }
      asm volatile("bar.sync 2, %0;" ::"n"(CZ_BUILD_WARPS * 32) : "memory");
      float hi[CZ_K], lo[CZ_K];
#pragma unroll
      for (int j = 0; j < CZ_K; ++j) {
        const float x = xs[r * CZ_S + j];
        hi[j] = __bfloat162float(__float2bfloat16_rn(x));
        lo[j] = x - hi[j];
      }
      // 16-byte chunks of the row: [h0..h7] [h8 h9 l0..l5] [l6..l9 h0..h3] [h4..h9 0 0]; chunk c lives at position c ^ (r & 7)
      const uint32_t row = smem_u32(smem + s * CZ_A_BYTES) + r * 128;
      const uint32_t sw = r & 7;
      sts128(row + ((0 ^ sw) << 4), pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]));
      sts128(row + ((1 ^ sw) << 4), pack_bf16x2(hi[8], hi[9]), pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]));
      sts128(row + ((2 ^ sw) << 4), pack_bf16x2(lo[6], lo[7]), pack_bf16x2(lo[8], lo[9]), pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]));
      sts128(row + ((3 ^ sw) << 4), pack_bf16x2(hi[4], hi[5]), pack_bf16x2(hi[6], hi[7]), pack_bf16x2(hi[8], hi[9]), 0u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&a_full[s]));
        mbar_arrive(smem_u32(&ss_full[s]));
      }
      t = next_tile(t, cslot, cphase);
      if (dyn && threadIdx.x == 0 && t >= 0) clc_issue();
    }
  } else if (warp == CZ_BUILD_WARPS) {
    // ===================== MMA issuer =====================
    // D = f32, A = B = bf16, K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CZ_BN >> 3) << 17) | ((uint32_t)(CZ_BM >> 4) << 24);
    if (lane == 0) mbar_wait(smem_u32(w_full), 0);
    int it = 0, cslot = 0;
    uint32_t cphase = 0;
    for (int t = blockIdx.x; t >= 0; t = next_tile(t, cslot, cphase), ++it) {   // lane 0 issues; the warp walks the tiles together
      if (lane == 0) {
        const int s = it & 1;
        mbar_wait(smem_u32(&a_full[s]), (it >> 1) & 1);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem + s * CZ_A_BYTES));
#pragma unroll
        for (int q = 0; q < CZ_NQ; ++q) {
          mbar_wait(smem_u32(&tmem_empty[q]), (it & 1) ^ 1);
          tc_fence_after();
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + CZ_OFF_W + q * (CZ_BN * 128)));
          umma_bf16(tmem_base + q * CZ_BN, adesc, bdesc, idesc, 0u);          // K 0..15
          umma_bf16(tmem_base + q * CZ_BN, adesc + 2, bdesc + 2, idesc, 1u);  // K 16..31
          umma_commit(smem_u32(&tmem_full[q]));
        }
        umma_commit(smem_u32(&a_empty[s]));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - (CZ_BUILD_WARPS + 1);   // 0..15
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int q = ew >> 2;                        // channel group [128 q, +128)
    const uint32_t tile = smem_u32(smem + CZ_OFF_TRANS) + ew * CZ_TRANS_WARP;
    int it = 0, cslot = 0;
    uint32_t cphase = 0;
    for (int t = blockIdx.x; t >= 0; t = next_tile(t, cslot, cphase), ++it) {
      const int s = it & 1;
      const int b = t / p.tiles_per_clip, t0 = (t % p.tiles_per_clip) * CZ_BM;
      const int rows_valid = p.L0 - (t0 + quarter * 32);
      mbar_wait(smem_u32(&ss_full[s]), (it >> 1) & 1);
      mbar_wait(smem_u32(&tmem_full[q]), it & 1);
      tc_fence_after();
      const uint32_t ssa = smem_u32(smem + CZ_OFF_SS) + s * (2 * CZ_C * 4);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int c0 = q * CZ_BN + ch * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * CZ_BN + ch * 32), v);
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 sc = lds128f(ssa + (c0 + 4 * j) * 4);
          const float4 sh = lds128f(ssa + (CZ_C + c0 + 4 * j) * 4);
          // packed fp32 (FFMA2): the GroupNorm affine and the GELU polynomial of two channels per instruction
          float a0, a1, a2, a3;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack_f32x2(sc.x, sc.y),
                                 pack_f32x2(sh.x, sh.y)), a0, a1);
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack_f32x2(sc.z, sc.w),
                                 pack_f32x2(sh.z, sh.w)), a2, a3);
          gelu_fast2(a0, a1, f[4 * j + 0], f[4 * j + 1]);
          gelu_fast2(a2, a3, f[4 * j + 2], f[4 * j + 3]);
        }
        const uint32_t buf = tile + (ch & 1) * CZ_TRANS_TILE;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that used this buffer has read it
        __syncwarp();
        const uint32_t wr_row = buf + lane * 64, wr_sw = (lane >> 1) & 3;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          sts128(wr_row + 16 * (g ^ wr_sw), pack_bf16x2(f[8 * g], f[8 * g + 1]), pack_bf16x2(f[8 * g + 2], f[8 * g + 3]),
                 pack_bf16x2(f[8 * g + 4], f[8 * g + 5]), pack_bf16x2(f[8 * g + 6], f[8 * g + 7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && rows_valid > 0) {                     // rows >= L0 are clipped by the tensor map
          tma_store_3d(buf, &map_out, c0, t0 + quarter * 32, b);
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&tmem_empty[q]));
        mbar_arrive(smem_u32(&ss_empty[s]));
      }
    }
  }

  if (warp > CZ_BUILD_WARPS && lane == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == CZ_BUILD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace avi

using namespace avi;

extern "C" int avi_w2v_conv0_pack_tc(const float* w, void* w_packed, void* stream) {
  conv0_pack_w_kernel<<<(CZ_C * 64 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<__nv_bfloat16*>(w_packed));
  return check_launch("conv0_pack_w");
}

extern "C" int avi_w2v_conv0_gn_gelu_tc(const float* audio, const float* w, const void* w_packed, const float* gn_w, const float* gn_b,
                                        void* stats, void* out, int64_t out_batch_stride, int32_t B, int32_t n_samples, int32_t C,
                                        float eps, void* stream) {
  AVI_REQUIRE(C == CZ_C, "avi_w2v_conv0_gn_gelu_tc: built for 512 channels (got %d)", C);
  AVI_REQUIRE(B > 0 && n_samples >= CZ_K, "avi_w2v_conv0_gn_gelu_tc: bad shape B=%d n=%d", B, n_samples);
  const int L0 = (n_samples - CZ_K) / CZ_S + 1;
  AVI_REQUIRE(out_batch_stride >= (int64_t)L0 * C && out_batch_stride % 8 == 0 && ((uintptr_t)out % 16 == 0) && ((uintptr_t)w_packed % 16 == 0),
              "avi_w2v_conv0_gn_gelu_tc: out_batch_stride too small or unaligned buffers");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch layout inside `stats` (B*C*2 doubles provided): [B][72] fp64 moments | scale [B][C] fp32 | shift [B][C] fp32
  double* mom = reinterpret_cast<double*>(stats);
  float* scale = reinterpret_cast<float*>(mom + (size_t)B * CZ_MOM_STRIDE);
  float* shift = scale + (size_t)B * C;
  if (cudaMemsetAsync(mom, 0, sizeof(double) * (size_t)B * CZ_MOM_STRIDE, st) != cudaSuccess) {
    set_error("avi_w2v_conv0_gn_gelu_tc: memset failed");
    return 1;
  }
  const int t_per_block = 4096;
  conv0_moments_kernel<<<dim3((L0 + t_per_block - 1) / t_per_block, B), 256, 0, st>>>(audio, mom, n_samples, L0, t_per_block);
  if (check_launch("conv0_moments")) return 1;
  conv0_gn_coeffs_kernel<<<B, CZ_C, 0, st>>>(mom, w, gn_w, gn_b, scale, shift, L0, eps);
  if (check_launch("conv0_gn_coeffs")) return 1;
  CUtensorMap map_w;
  {
    uint64_t dims[2] = {64, (uint64_t)CZ_C};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, CZ_BN};
    if (encode_map(&map_w, w_packed, 2, dims, strides, box)) return 1;
  }
  CUtensorMap map_out;
  {
    uint64_t dims[3] = {(uint64_t)CZ_C, (uint64_t)L0, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)CZ_C * 2, (uint64_t)out_batch_stride * 2};
    uint32_t box[3] = {32, 32, 1};
    if (encode_map(&map_out, out, 3, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  }
  Conv0Params p;
  p.audio = audio;
  p.scale = scale;
  p.shift = shift;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.out_batch_stride = out_batch_stride;
  p.n_samples = n_samples;
  p.L0 = L0;
  p.tiles_per_clip = (L0 + CZ_BM - 1) / CZ_BM;
  p.total_tiles = p.tiles_per_clip * B;
  static SmemOptIn optin;
  const cudaError_t attr_err = smem_optin(conv0_tc_kernel, (int)CZ_SMEM, optin);
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_w2v_conv0_gn_gelu_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int sms = device_sms();
  p.dynamic = (dynamic_tiles_mask() & AVI_DYN_CONV0) != 0 && p.total_tiles > 1 ? 1 : 0;
  const int grid = p.dynamic ? p.total_tiles : (p.total_tiles < sms ? p.total_tiles : sms);
  conv0_tc_kernel<<<grid, CZ_THREADS, CZ_SMEM, st>>>(map_w, map_out, p);
  return check_launch("conv0_tc");
}
