// bf16 tensor-core GEMM for sm_100a: tcgen05.mma (cta_group::1, M=128, N=256, K=16) with the accumulator in TMEM,
// operands staged in shared memory by TMA (SWIZZLE_128B, K-major), a 4-stage mbarrier ring, two TMEM accumulator
// stages so the epilogue of tile i overlaps the MMAs of tile i+1, persistent over the SMs.
//
//   C[b, r, n] = epi( sum_k A[b, r, k] * W[n, k] + bias[n] ) (+ residual[b, r, n])
//
// A is addressed through a 4-D tensor map (channel, phase, super-row, clip) so that a strided Conv1d over
// time-major activations is the same code path as a plain GEMM: output row r, tap j reads input row
// r*stride + j = super-row (r + j/stride), phase (j % stride).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue
// (TMEM lane quarter = warp_id % 4, column half = (warp_id - 2) / 4).
#include "tc_common.cuh"

namespace avi {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64, TC_STAGES = 4, TC_UMMA_K = 16;
constexpr int TC_THREADS = 320, TC_EPI_WARPS = 8;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB
constexpr uint32_t TC_B_BYTES = TC_BN * TC_BK * 2;   // 32 KB
constexpr uint32_t TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr uint32_t TC_TRANS_BYTES = TC_EPI_WARPS * 32 * 32 * 4;  // per-warp 32x32 fp32 transpose tiles (XOR-swizzled, no padding)
constexpr uint32_t TC_BIAS_BYTES = TC_BN * 4;
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + TC_TRANS_BYTES + TC_BIAS_BYTES + 256 /*barriers*/ + 1024 /*align*/;
static_assert(TC_SMEM_BYTES <= 232448, "shared memory budget");

struct TcParams {
  const float* bias;
  const float* residual;
  void* C;
  void* C2;
  int c_dtype;  // dtype of C; C2 (if any) is the other one
  int act;
  int batch, rows, N;
  int num_kb;           // K / 64
  int kb_per_tap;       // C_in / 64
  int conv_stride;
  int m_tiles, n_tiles, total_tiles;
  int64_t c_ld, c_batch_stride, res_ld, res_batch_stride;
  int vec_ok;           // outputs (and residual) are 16-byte addressable per 32-column chunk
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == AVI_ACT_GELU) return gelu_erf(v);
  if (act == AVI_ACT_RELU) return fmaxf(v, 0.f);
  if (act == AVI_ACT_QUICK_GELU) return quick_gelu(v);
  return v;
}

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + TC_STAGES * TC_A_BYTES;
  float* trans = reinterpret_cast<float*>(smem + TC_STAGES * TC_STAGE_BYTES);
  float* sbias = reinterpret_cast<float*>(smem + TC_STAGES * TC_STAGE_BYTES + TC_TRANS_BYTES);  // [BN], single-buffered
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES + TC_TRANS_BYTES + TC_BIAS_BYTES);
  uint64_t* full_bar = bars;                        // [STAGES]
  uint64_t* empty_bar = bars + TC_STAGES;           // [STAGES]
  uint64_t* tmem_full = bars + 2 * TC_STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * TC_STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tmem_full[s]), 1);
      mbar_init(smem_u32(&tmem_empty[s]), TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // whole warp: allocate all 512 TMEM columns (2 accumulator stages x 256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int n_blk = t % p.n_tiles;
        const int mt = t / p.n_tiles;
        const int b = mt / p.m_tiles, m_blk = mt % p.m_tiles;
        int kin = 0, ph = 0, sr = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, TC_STAGE_BYTES);
          tma_load_4d(smem_u32(smem_a + stage * TC_A_BYTES), &map_a, fb, kin * TC_BK, ph, m_blk * TC_BM + sr, b);
          tma_load_2d(smem_u32(smem_b + stage * TC_B_BYTES), &map_w, fb, kb * TC_BK, n_blk * TC_BN);
          if (++kin == p.kb_per_tap) {
            kin = 0;
            if (++ph == p.conv_stride) {
              ph = 0;
              ++sr;
            }
          }
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N=256, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&tmem_empty[as]), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * TC_BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * TC_A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * TC_B_BYTES));
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[stage]));  // frees the smem slot once these MMAs have read it
          if (++stage == TC_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(smem_u32(&tmem_full[as]));  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> smem transpose -> coalesced global =====================
    // Phase A (thread = accumulator row): tcgen05.ld 32 columns, + bias, activation, written to a per-warp staging tile.
    // Phase B (lanes span the columns of a few rows): 16-byte shared loads and fully coalesced 16-byte global accesses
    // (residual read + output write).  Three variants: packed bf16 rows, fp32 rows (+ residual), and a scalar one for
    // unaligned / ragged outputs (the 15069-wide vertex rows).  The XOR swizzles keep every shared access conflict-free.
    const int ew = warp - 2;              // 0..7
    const int quarter = warp & 3;         // TMEM lanes [32*quarter, +32) are the only ones this warp may touch
    const int half = ew >> 2;             // columns [128*half, +128)
    const int etid = threadIdx.x - 64;    // 0..255
    const uint32_t tile = smem_u32(trans) + ew * (32 * 32 * 4);
    const uint32_t sbias_a = smem_u32(sbias);
    const bool fast_bf16 = p.vec_ok && p.c_dtype == AVI_DT_BF16 && p.C2 == nullptr && p.residual == nullptr;
    const bool fast_f32 = p.vec_ok && p.c_dtype == AVI_DT_F32 && p.C2 == nullptr;
    int it = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int n_blk = t % p.n_tiles;
      const int mt = t / p.n_tiles;
      const int b = mt / p.m_tiles, m_blk = mt % p.m_tiles;
      const int n0 = n_blk * TC_BN;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's bias reads are done
      {
        const int n = n0 + etid;
        sbias[etid] = (p.bias != nullptr && n < p.N) ? p.bias[n] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(smem_u32(&tmem_full[as]), aphase);
      tc_fence_after();
      const int row_base = m_blk * TC_BM + quarter * 32;
      const int rows_valid = p.rows - row_base;  // may be <= 0 or > 32
      const int64_t c_row0 = (int64_t)b * p.c_batch_stride + (int64_t)row_base * p.c_ld;
      const int64_t r_row0 = (int64_t)b * p.res_batch_stride + (int64_t)row_base * p.res_ld;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int col0 = half * 128 + ch * 32;
        const int n_base = n0 + col0;
        if (n_base >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TC_BN + col0), v);
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = lds128f(sbias_a + (col0 + 4 * j) * 4);
          f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x;
          f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z;
          f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
        }
        if (p.act == AVI_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
        } else if (p.act == AVI_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act == AVI_ACT_QUICK_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
        }
        const bool full_n = n_base + 32 <= p.N;
        if (fast_bf16 && full_n) {
          // tile: 32 rows x 16 packed words, pitch 16 words; 4-word group g of row r lives at group g ^ ((r >> 1) & 3)
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * g + 2 * u], f[8 * g + 2 * u + 1]);
              w[u] = *reinterpret_cast<uint32_t*>(&h2);
            }
            sts128(tile + (lane * 16 + 4 * (g ^ ((lane >> 1) & 3))) * 4, w[0], w[1], w[2], w[3]);
          }
          __syncwarp();
          __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(p.C) + c_row0 + n_base;
          const int g = lane & 3;
#pragma unroll
          for (int i8 = 0; i8 < 4; ++i8) {
            const int r = i8 * 8 + (lane >> 2);
            if (r < rows_valid) {
              const uint4 q = lds128(tile + (r * 16 + 4 * (g ^ ((r >> 1) & 3))) * 4);
              *reinterpret_cast<uint4*>(cb + (int64_t)r * p.c_ld + 8 * g) = q;
            }
          }
          __syncwarp();
        } else if (fast_f32 && full_n) {
          // tile: 32 rows x 32 words, pitch 32; 4-word group g of row r lives at group g ^ (r & 7)
#pragma unroll
          for (int g = 0; g < 8; ++g)
            sts128(tile + (lane * 32 + 4 * (g ^ (lane & 7))) * 4, __float_as_uint(f[4 * g]), __float_as_uint(f[4 * g + 1]),
                   __float_as_uint(f[4 * g + 2]), __float_as_uint(f[4 * g + 3]));
          __syncwarp();
          float* cf = reinterpret_cast<float*>(p.C) + c_row0 + n_base;
          const float* rp = p.residual ? p.residual + r_row0 + n_base : nullptr;
          const int g = lane & 7;
          float4 rr[8];
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const int r = i4 * 4 + (lane >> 3);
            rr[i4] = (rp != nullptr && r < rows_valid) ? __ldg(reinterpret_cast<const float4*>(rp + (int64_t)r * p.res_ld + 4 * g))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const int r = i4 * 4 + (lane >> 3);
            if (r < rows_valid) {
              float4 q = lds128f(tile + (r * 32 + 4 * (g ^ (r & 7))) * 4);
              q.x += rr[i4].x;
              q.y += rr[i4].y;
              q.z += rr[i4].z;
              q.w += rr[i4].w;
              *reinterpret_cast<float4*>(cf + (int64_t)r * p.c_ld + 4 * g) = q;
            }
          }
          __syncwarp();
        } else {
          // generic: element (row, col j) at word row*32 + (j ^ row); lane = column in phase B
#pragma unroll
          for (int j = 0; j < 32; ++j) sts32(tile + (lane * 32 + (j ^ lane)) * 4, __float_as_uint(f[j]));
          __syncwarp();
          const int n = n_base + lane;
          if (n < p.N) {
            float* cf = nullptr;
            __nv_bfloat16* cb = nullptr;
            if (p.c_dtype == AVI_DT_F32) {
              cf = reinterpret_cast<float*>(p.C);
              cb = reinterpret_cast<__nv_bfloat16*>(p.C2);
            } else {
              cb = reinterpret_cast<__nv_bfloat16*>(p.C);
              cf = reinterpret_cast<float*>(p.C2);
            }
            const int rmax = rows_valid < 32 ? rows_valid : 32;
            int64_t co = c_row0 + n;
            int64_t ro = r_row0 + n;
#pragma unroll 4
            for (int r = 0; r < rmax; ++r) {
              float val = __uint_as_float(lds32(tile + (r * 32 + (lane ^ r)) * 4));
              if (p.residual) val += p.residual[ro];
              if (cf) cf[co] = val;
              if (cb) cb[co] = __float2bfloat16_rn(val);
              co += p.c_ld;
              ro += p.res_ld;
            }
          }
          __syncwarp();
        }
      }
      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[as]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

static const char* tc_check(const AviGemmArgs* a) {
  if (!a) return "null args";
  if (a->a_dtype != AVI_DT_BF16) return "A/W must be bf16";
  if (a->batch <= 0 || a->rows <= 0 || a->N <= 0 || a->K <= 0) return "bad shape";
  if (a->conv_taps < 1 || a->conv_stride < 1) return "bad conv params";
  if (a->K % a->conv_taps != 0) return "K must be taps * C_in";
  const int cin = a->K / a->conv_taps;
  if (cin % TC_BK != 0) return "C_in (K per tap) must be a multiple of 64";
  if (a->a_ld < cin) return "a_ld smaller than the channel window";
  if (a->a_ld % 8 != 0 || a->a_batch_stride % 8 != 0) return "a_ld and a_batch_stride must be multiples of 8 elements";
  if (((uintptr_t)a->A | (uintptr_t)a->W) % 16 != 0) return "A and W must be 16-byte aligned";
  if (a->a_rows_alloc < (int64_t)(a->rows - 1) * a->conv_stride + a->conv_taps) return "a_rows_alloc smaller than the rows read";
  if (a->conv_taps_x != 0) return "2-D taps are not built in the single-CTA kernel";
  if (a->C == nullptr) return "C is null";
  return nullptr;
}

}  // namespace avi

using namespace avi;

// single-CTA variant (kept for A/B measurements against the CTA-pair kernel of gemm_tc2.cu: AVI_GEMM_V1=1)
extern "C" int avi_gemm_bf16_tc_v1(const AviGemmArgs* a, void* stream) {
  const char* why = tc_check(a);
  AVI_REQUIRE(why == nullptr, "avi_gemm_bf16_tc: %s", why);
  const int cin = a->K / a->conv_taps;
  const int s = a->conv_stride;
  CUtensorMap map_a, map_w;
  {
    // (channel, phase, super-row, clip)
    const uint64_t q_rows = (uint64_t)(a->a_rows_alloc / s);
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)s, q_rows, (uint64_t)a->batch};
    uint64_t strides[3] = {(uint64_t)a->a_ld * 2, (uint64_t)a->a_ld * s * 2, (uint64_t)a->a_batch_stride * 2};
    if (a->batch == 1) strides[2] = dims[2] * strides[1];  // unused; keep it well-formed
    uint32_t box[4] = {TC_BK, 1, TC_BM, 1};
    if (encode_map(&map_a, a->A, 4, dims, strides, box)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t strides[1] = {(uint64_t)a->K * 2};
    uint32_t box[2] = {TC_BK, TC_BN};
    if (encode_map(&map_w, a->W, 2, dims, strides, box)) return 1;
  }
  TcParams p;
  p.bias = a->bias;
  p.residual = a->residual;
  p.C = a->C;
  p.C2 = a->C2;
  p.c_dtype = a->c_dtype;
  p.act = a->act;
  p.batch = a->batch;
  p.rows = a->rows;
  p.N = a->N;
  p.num_kb = a->K / TC_BK;
  p.kb_per_tap = cin / TC_BK;
  p.conv_stride = s;
  p.m_tiles = (a->rows + TC_BM - 1) / TC_BM;
  p.n_tiles = (a->N + TC_BN - 1) / TC_BN;
  p.total_tiles = p.m_tiles * p.n_tiles * a->batch;
  p.c_ld = a->c_ld;
  p.c_batch_stride = a->c_batch_stride;
  p.res_ld = a->res_ld;
  p.res_batch_stride = a->res_batch_stride;
  // vector path: every 32-column chunk of every row is 16-byte addressable in both output dtypes and the residual
  bool vec = (a->c_ld % 8 == 0) && (a->c_batch_stride % 8 == 0) && ((uintptr_t)a->C % 16 == 0) &&
             (a->C2 == nullptr || (uintptr_t)a->C2 % 16 == 0);
  if (a->residual) vec = vec && (a->res_ld % 4 == 0) && (a->res_batch_stride % 4 == 0) && ((uintptr_t)a->residual % 16 == 0);
  p.vec_ok = vec ? 1 : 0;

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gemm_bf16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
  });
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_gemm_bf16_tc: cannot opt in to %u bytes of shared memory: %s", TC_SMEM_BYTES,
              cudaGetErrorString(attr_err));
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  gemm_bf16_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(map_a, map_w, p);
  return check_launch("gemm_bf16_tc");
}
