// Fused multi-head attention for the wav2vec2 encoder on tcgen05 (sm_100a):  out = softmax(Q K^T * scale) V
// per (clip, head, 128-query tile), whole key range (T <= 256) in one shot. PERSISTENT and software-pipelined (round 2): one CTA per
// SM walks its share of the B*H*ceil(T/128) work items, with three roles that overlap across items:
//   warp 0   TMA producer : Q [128x64] + K [256x64] into a 2-stage ring, V [256x64] into its own 2-stage ring (swizzled bf16 tiles
//                           of the packed qkv activation). Q/K of item i+2 are requested as soon as MMA 1 of item i has read them.
//   warp 1   MMA issuer   : MMA 1  S[128x256] (fp32, TMEM, 2 stages) = Q K^T        4 x tcgen05.mma (M128 N256 K16), K-major operands
//                           MMA 2  O[128x64]  (TMEM, overlays S)     = P V         16 x tcgen05.mma (M128 N64  K16), V MN-major
//                           issue order MMA1(0), { MMA1(i+1), MMA2(i) }: the tensor pipe computes the next item's scores while the
//                           softmax warps work on the current one.
//   warps 2..17 softmax   : thread = (query row, 64-key quarter). ONE pass over TMEM (the 64 scores stay in registers), row max and
//                           row sum exchanged between the 4 threads of a row through shared memory, P written as bf16 into the
//                           K-major SWIZZLE_128B layout MMA 2 consumes; then O / rowsum -> bf16 -> global.
// Round 1 launched one CTA per item (1536 CTAs of 5 serial phases each: TMA round trip, MMA, softmax, MMA, store): 63 us per layer,
// tensor pipe 10 % active, a quarter of the stall samples on the Q/K load. TMEM read bandwidth (one pass over S: 128 KB per item)
// is what bounds this version.
// (HF Wav2Vec2Attention / eager_attention_forward as called from models/lib/wav2vec.py:142; no mask on this path.)
#include "tc_common.cuh"

namespace avi {

constexpr int AT_D = 64, AT_BM = 128, AT_BN = 256;
constexpr uint32_t AT_Q_BYTES = AT_BM * AT_D * 2;   // 16 KB
constexpr uint32_t AT_K_BYTES = AT_BN * AT_D * 2;   // 32 KB
constexpr uint32_t AT_QK_BYTES = AT_Q_BYTES + AT_K_BYTES;
constexpr uint32_t AT_P_BYTES = AT_BM * AT_BN * 2;  // 64 KB: 4 chunks (64 keys) of [128 rows x 128 B]
constexpr uint32_t AT_OFF_V = 2 * AT_QK_BYTES;               // 96 KB
constexpr uint32_t AT_OFF_P = AT_OFF_V + 2 * AT_K_BYTES;     // 160 KB
constexpr uint32_t AT_OFF_XCH = AT_OFF_P + AT_P_BYTES;       // 224 KB: [4 key quarters][128 rows] floats (row max, then row sum)
constexpr uint32_t AT_OFF_BAR = AT_OFF_XCH + 4 * AT_BM * 4;
constexpr uint32_t AT_SMEM = AT_OFF_BAR + 256;
constexpr int AT_SM_WARPS = 16, AT_THREADS = (2 + AT_SM_WARPS) * 32;
static_assert(AT_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

struct AttnParams {
  __nv_bfloat16* out;
  int T, H, B, q_tiles, total_items;
  float scale_log2e;
};

__global__ void __launch_bounds__(AT_THREADS, 1)   // 18 warps = 5 on one SM sub-partition: 16 K registers / (5 x 32) caps a thread at 96 (a 112-register build fails to launch)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQK = smem;                      // [2][Q 16 KB | K 32 KB]
  uint8_t* sV = smem + AT_OFF_V;            // [2][32 KB]
  uint8_t* sP = smem + AT_OFF_P;            // 64 KB
  float* xch = reinterpret_cast<float*>(smem + AT_OFF_XCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t* qk_full = bars;         // [2] TMA bytes of Q and K landed
  uint64_t* qk_empty = bars + 2;    // [2] MMA 1 has read them
  uint64_t* v_full = bars + 4;      // [2]
  uint64_t* v_empty = bars + 6;     // [2] MMA 2 has read V (and P)
  uint64_t* s_full = bars + 8;      // [2] S complete in TMEM stage
  uint64_t* s_empty = bars + 10;    // [2] the softmax warps have read O out of the stage (16 warps arrive)
  uint64_t* p_full = bars + 12;     // [1] P written (16 warps arrive)
  uint64_t* o_full = bars + 13;     // [1] O complete (MMA 2 done: P may be rewritten)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = p.H * AT_D;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&qk_full[s]), 1);
      mbar_init(smem_u32(&qk_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1);
      mbar_init(smem_u32(&v_empty[s]), 1);
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), AT_SM_WARPS);
    }
    mbar_init(smem_u32(p_full), AT_SM_WARPS);
    mbar_init(smem_u32(o_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // both S stages: all 512 TMEM columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // the prologue above may overlap the tail of the QKV GEMM (programmatic dependent launch)

  // item -> (clip, head, query tile); the query tiles of one (clip, head) are neighbours, so K / V are shared through L2
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_items = first < p.total_items ? (p.total_items - first + stride - 1) / stride : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < n_items; ++i) {
        const int item = first + i * stride;
        const int qt = item % p.q_tiles, bh = item / p.q_tiles;
        const int h = bh % p.H, b = bh / p.H;
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(smem_u32(&qk_empty[s]), ph ^ 1);
        mbar_expect_tx(smem_u32(&qk_full[s]), AT_QK_BYTES);
        tma_load_2d(smem_u32(sQK + s * AT_QK_BYTES), &map_q, smem_u32(&qk_full[s]), h * AT_D, b * p.T + qt * AT_BM);
        tma_load_2d(smem_u32(sQK + s * AT_QK_BYTES + AT_Q_BYTES), &map_kv, smem_u32(&qk_full[s]), E + h * AT_D, b * p.T);
        mbar_wait(smem_u32(&v_empty[s]), ph ^ 1);
        mbar_expect_tx(smem_u32(&v_full[s]), AT_K_BYTES);
        tma_load_2d(smem_u32(sV + s * AT_K_BYTES), &map_kv, smem_u32(&v_full[s]), 2 * E + h * AT_D, b * p.T);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_items > 0) {
      // S = Q K^T : D=f32, A=B=bf16 K-major, M=128, N=256.   O = P V : B (V) MN-major, M=128, N=64
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BN >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AT_D >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      auto mma1 = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(smem_u32(&s_empty[s]), ph ^ 1);     // O of item i-2 has been read out of this TMEM stage
        mbar_wait(smem_u32(&qk_full[s]), ph);
        tc_fence_after();
        const uint64_t qd = umma_desc_sw128(smem_u32(sQK + s * AT_QK_BYTES));
        const uint64_t kd = umma_desc_sw128(smem_u32(sQK + s * AT_QK_BYTES + AT_Q_BYTES));
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tmem_base + s * AT_BN, qd + 2 * k, kd + 2 * k, idesc1, k != 0 ? 1u : 0u);
        umma_commit(smem_u32(&qk_empty[s]));
        umma_commit(smem_u32(&s_full[s]));
      };
      mma1(0);
      for (int i = 0; i < n_items; ++i) {
        if (i + 1 < n_items) mma1(i + 1);
        const int s = i & 1;
        mbar_wait(smem_u32(p_full), i & 1);
        mbar_wait(smem_u32(&v_full[s]), (i >> 1) & 1);
        tc_fence_after();
        const uint64_t vd = umma_desc_sw128_mn(smem_u32(sV + s * AT_K_BYTES));
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k) {
          const uint64_t pd = umma_desc_sw128(smem_u32(sP + (k >> 2) * (AT_BM * 128))) + 2 * (k & 3);
          umma_bf16(tmem_base + s * AT_BN, pd, vd + (uint64_t)k * (16 * 128 >> 4), idesc2, k != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&v_empty[s]));
        umma_commit(smem_u32(o_full));
      }
    }
  } else {
    // ===================== softmax + epilogue warps =====================
    const int ew = warp - 2;                    // 0..15
    const int qr = warp & 3;                    // TMEM lane quarter this warp may touch
    const int kq = ew >> 2;                     // key quarter [64 kq, +64) = chunk kq of P; output features [16 kq, +16)
    const int row = qr * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qr * 32) << 16);
    const int bar_id = 1 + qr;                  // the 4 warps (128 threads) that share these 32 rows
#define AT_LD32(arr, addr)                                                                                                              \
  asm volatile(                                                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                                         \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                                         \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                         \
      : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]), "=r"(arr[8]),   \
        "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]), "=r"(arr[15]), "=r"(arr[16]),          \
        "=r"(arr[17]), "=r"(arr[18]), "=r"(arr[19]), "=r"(arr[20]), "=r"(arr[21]), "=r"(arr[22]), "=r"(arr[23]), "=r"(arr[24]),         \
        "=r"(arr[25]), "=r"(arr[26]), "=r"(arr[27]), "=r"(arr[28]), "=r"(arr[29]), "=r"(arr[30]), "=r"(arr[31])                         \
      : "r"(addr))
// Stage A of item I: the 64 scores of this thread (ONE pass over TMEM), keys beyond T set to -inf once (only the last key quarter of
// a ragged T sees any: warp-uniform branch, no per-element predicates later), row max exchanged between the 4 threads of the row.
#define AT_STAGE_A(I, A0, A1, MXS)                                                                                                      \
  {                                                                                                                                     \
    const int s_ = (I) & 1;                                                                                                             \
    if (lane == 0) mbar_wait(smem_u32(&s_full[s_]), ((I) >> 1) & 1);                                                                    \
    __syncwarp();                                                                                                                       \
    tc_fence_after();                                                                                                                   \
    const uint32_t ta_ = lane_addr + (uint32_t)(s_ * AT_BN + kq * 64);                                                                  \
    AT_LD32(A0, ta_);                                                                                                                   \
    AT_LD32(A1, ta_ + 32);                                                                                                              \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");                                                                        \
    if (nvalid < 64) {                                                                                                                  \
      _Pragma("unroll") for (int j = 0; j < 32; ++j) {                                                                                  \
        if (j >= nvalid) A0[j] = 0xff800000u;                                                                                           \
        if (32 + j >= nvalid) A1[j] = 0xff800000u;                                                                                      \
      }                                                                                                                                 \
    }                                                                                                                                   \
    float mx_ = -INFINITY;                                                                                                              \
    _Pragma("unroll") for (int j = 0; j < 32; ++j) mx_ = fmaxf(mx_, fmaxf(__uint_as_float(A0[j]), __uint_as_float(A1[j])));             \
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); /* the previous item's row sums have been read by the whole quarter */ \
    xch[kq * AT_BM + row] = mx_;                                                                                                        \
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");                                                                         \
    mx_ = fmaxf(fmaxf(xch[row], xch[AT_BM + row]), fmaxf(xch[2 * AT_BM + row], xch[3 * AT_BM + row]));                                  \
    MXS = mx_ * p.scale_log2e;                                                                                                          \
  }

    const int nvalid = p.T - kq * 64;
    uint32_t v0[32], v1[32], n0[32], n1[32];
    float mxs = 0.f, mxs_next = 0.f;
    if (n_items > 0) AT_STAGE_A(0, v0, v1, mxs);
    for (int i = 0; i < n_items; ++i) {
      const int item = first + i * stride;
      const int qt = item % p.q_tiles, bh = item / p.q_tiles;
      const int h = bh % p.H, b = bh / p.H;
      const int s = i & 1;
      // ---- stage B: probabilities (bf16, what MMA 2 consumes; the row sum is taken over the rounded values); ex2.approx(-inf) = 0
      uint32_t pk[32];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v0[j]), p.scale_log2e, -mxs));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v0[j + 1]), p.scale_log2e, -mxs));
        const float p2 = ex2_approx(fmaf(__uint_as_float(v1[j]), p.scale_log2e, -mxs));
        const float p3 = ex2_approx(fmaf(__uint_as_float(v1[j + 1]), p.scale_log2e, -mxs));
        const uint32_t ua = pack_bf16x2_rn(p0, p1), ub = pack_bf16x2_rn(p2, p3);
        // the rounded values back as fp32: low half << 16, high half masked
        sum += (__uint_as_float(ua << 16) + __uint_as_float(ua & 0xffff0000u)) + (__uint_as_float(ub << 16) + __uint_as_float(ub & 0xffff0000u));
        pk[j >> 1] = ua;
        pk[16 + (j >> 1)] = ub;
      }
      // P of item i-1 has been consumed: every warp waited for o_full(i-1) in stage C of the previous iteration
      {
        const uint32_t prow = smem_u32(sP) + kq * (AT_BM * 128) + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sts128(prow + ((q ^ (row & 7)) * 16), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      // row sums through the same exchange buffer (every thread of the quarter read the maxima before it got here)
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      xch[kq * AT_BM + row] = sum;
      fence_proxy_async_smem();     // P (generic-proxy stores) visible to the tensor core
      tc_fence_before();            // the TMEM reads of stage A are ordered before MMA 2, which overwrites the stage with O
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(p_full));
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      const float inv = 1.f / ((xch[row] + xch[AT_BM + row]) + (xch[2 * AT_BM + row] + xch[3 * AT_BM + row]));
      // ---- stage A of the NEXT item while the tensor pipe runs MMA 2 of this one (its scores were computed during stage B)
      if (i + 1 < n_items) AT_STAGE_A(i + 1, n0, n1, mxs_next);
      // ---- stage C: O / rowsum -> bf16 -> global
      if (lane == 0) mbar_wait(smem_u32(o_full), i & 1);
      __syncwarp();
      tc_fence_after();
      uint32_t o[16];
      tmem_ld16(lane_addr + (uint32_t)(s * AT_BN + kq * 16), o);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));          // the TMEM stage may take the scores of item i+2
      const int t = qt * AT_BM + row;
      if (t < p.T) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((int64_t)b * p.T + t) * E + h * AT_D + kq * 16);
        uint32_t w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = pack_bf16x2_rn(__uint_as_float(o[2 * u]) * inv, __uint_as_float(o[2 * u + 1]) * inv);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v0[j] = n0[j];
        v1[j] = n1[j];
      }
      mxs = mxs_next;
    }
#undef AT_STAGE_A
#undef AT_LD32
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace avi

using namespace avi;

extern "C" int avi_mha_fwd_tc_supported(int32_t dtype, int32_t T, int32_t D) {
  return (dtype == AVI_DT_BF16 && D == AT_D && T >= 1 && T <= AT_BN) ? 1 : 0;
}

extern "C" int avi_mha_fwd_tc(const void* qkv, void* out, int32_t B, int32_t T, int32_t H, int32_t D, float scale, void* stream) {
  AVI_REQUIRE(B > 0 && H > 0 && D == AT_D && T >= 1 && T <= AT_BN, "avi_mha_fwd_tc: needs head dim 64 and T <= 256 (T=%d D=%d)", T, D);
  AVI_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), "avi_mha_fwd_tc: unaligned pointers");
  const int E = H * D;
  CUtensorMap map_q, map_kv;
  uint64_t dims[2] = {(uint64_t)3 * E, (uint64_t)B * T};
  uint64_t strides[1] = {(uint64_t)3 * E * 2};
  uint32_t box_q[2] = {AT_D, AT_BM}, box_kv[2] = {AT_D, AT_BN};
  if (encode_map(&map_q, qkv, 2, dims, strides, box_q)) return 1;
  if (encode_map(&map_kv, qkv, 2, dims, strides, box_kv)) return 1;
  static SmemOptIn optin;
  const cudaError_t attr_err = smem_optin(attn_tc_kernel, (int)AT_SMEM, optin);
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_mha_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  AttnParams p;
  p.out = (__nv_bfloat16*)out;
  p.T = T;
  p.H = H;
  p.B = B;
  p.q_tiles = (T + AT_BM - 1) / AT_BM;
  p.total_items = B * H * p.q_tiles;
  p.scale_log2e = scale * 1.4426950408889634f;
  const int sms = device_sms();
  const int grid = p.total_items < sms ? p.total_items : sms;
  const cudaError_t le = launch_pdl(attn_tc_kernel, dim3(grid), dim3(AT_THREADS), AT_SMEM, (cudaStream_t)stream, map_q, map_kv, p);
  AVI_REQUIRE(le == cudaSuccess, "avi_mha_fwd_tc: launch failed: %s", cudaGetErrorString(le));
  return check_launch("attn_tc");
}
