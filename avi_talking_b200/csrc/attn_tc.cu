// Fused multi-head attention for the wav2vec2 encoder on tcgen05 (sm_100a):  out = softmax(Q K^T * scale) V
// per (clip, head, 128-query tile), whole key range (T <= 256) in one shot:
//   TMA   : Q [128x64], K [256x64], V [256x64] bf16 tiles of the packed qkv activation -> swizzled smem
//   MMA 1 : S[128x256] (fp32, TMEM) = Q K^T               4 x tcgen05.mma (M128 N256 K16), both operands K-major
//   warps : each thread owns one query row: two passes over TMEM (row max; exp2, row sum), P written as bf16 into smem
//           in the K-major SWIZZLE_128B layout the next MMA consumes (P overlays the Q/K tiles, which are dead by then)
//   MMA 2 : O[128x64] (TMEM, overlays S) = P V            16 x tcgen05.mma (M128 N64 K16), V is the MN-major B operand
//   epilogue: O / rowsum -> bf16 -> global
// 96 KB of smem and 256 TMEM columns per CTA, so two CTAs share an SM and overlap each other's serial phases.
// (HF Wav2Vec2Attention / eager_attention_forward as called from models/lib/wav2vec.py:142; no mask on this path.)
#include "tc_common.cuh"

namespace avi {

constexpr int AT_D = 64, AT_BM = 128, AT_BN = 256;
constexpr uint32_t AT_Q_BYTES = AT_BM * AT_D * 2;   // 16 KB
constexpr uint32_t AT_K_BYTES = AT_BN * AT_D * 2;   // 32 KB
constexpr uint32_t AT_P_BYTES = AT_BM * AT_BN * 2;  // 64 KB, overlays Q | K | pad
constexpr uint32_t AT_V_OFF = AT_P_BYTES;            // V after the P region
constexpr uint32_t AT_XCH_OFF = AT_P_BYTES + AT_K_BYTES + 64;   // row max / row sum exchange between the two column halves
constexpr uint32_t AT_SMEM = AT_XCH_OFF + 2 * 2 * AT_BM * 4 + 1024;
constexpr int AT_THREADS = 288, AT_CTRL_WARP = 8;

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
               __nv_bfloat16* __restrict__ out, int T, int H, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 16 KB  } overlaid by P (64 KB) once S has been computed
  uint8_t* sK = smem + AT_Q_BYTES;          // 32 KB  }
  uint8_t* sP = smem;
  uint8_t* sV = smem + AT_V_OFF;            // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_V_OFF + AT_K_BYTES);
  uint64_t* bar_load = bars;       // TMA bytes landed
  uint64_t* bar_s = bars + 1;      // S complete
  uint64_t* bar_p = bars + 2;      // P written by all 128 softmax threads
  uint64_t* bar_o = bars + 3;      // O complete
  uint64_t* bar_v = bars + 4;      // V landed (not needed before the second MMA: kept off the critical path to S and the softmax)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM, h = blockIdx.y, b = blockIdx.z;
  const int E = H * AT_D;

  if (warp == AT_CTRL_WARP) {
    if (lane == 0) {
      mbar_init(smem_u32(bar_load), 1);
      mbar_init(smem_u32(bar_s), 1);
      mbar_init(smem_u32(bar_p), 256);
      mbar_init(smem_u32(bar_o), 1);
      mbar_init(smem_u32(bar_v), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == AT_CTRL_WARP) {
    if (lane == 0) {
      const uint32_t lb = smem_u32(bar_load), vb = smem_u32(bar_v);
      mbar_expect_tx(lb, AT_Q_BYTES + AT_K_BYTES);
      tma_load_2d(smem_u32(sQ), &map_q, lb, h * AT_D, b * T + q0);
      tma_load_2d(smem_u32(sK), &map_kv, lb, E + h * AT_D, b * T);
      mbar_expect_tx(vb, AT_K_BYTES);
      tma_load_2d(smem_u32(sV), &map_kv, vb, 2 * E + h * AT_D, b * T);
      mbar_wait(lb, 0);
      tc_fence_after();
      // S = Q K^T : D=f32, A=B=bf16 K-major, M=128, N=256
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BN >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      const uint64_t qd = umma_desc_sw128(smem_u32(sQ)), kd = umma_desc_sw128(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tmem_base, qd + 2 * k, kd + 2 * k, idesc1, k != 0 ? 1u : 0u);
      umma_commit(smem_u32(bar_s));
      // O = P V : A = P (K-major over keys, 4 chunks of 64 keys), B = V (MN-major: d contiguous), M=128, N=64
      mbar_wait(smem_u32(bar_p), 0);
      mbar_wait(smem_u32(bar_v), 0);
      tc_fence_after();
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AT_D >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      const uint64_t vd = umma_desc_sw128_mn(smem_u32(sV));
#pragma unroll
      for (int k = 0; k < AT_BN / 16; ++k) {
        const uint64_t pd = umma_desc_sw128(smem_u32(sP + (k >> 2) * (AT_BM * 128))) + 2 * (k & 3);
        umma_bf16(tmem_base, pd, vd + (uint64_t)k * (16 * 128 >> 4), idesc2, k != 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(bar_o));
    }
  } else {
    // ---------------- softmax warps: two threads per query row (TMEM lane), one per half of the key range ----------------
    // warps w and w + 4 share TMEM lane quarter w % 4; `half` selects keys [128 half, +128) and output features [32 half, +32)
    const int row = (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float* xmax = reinterpret_cast<float*>(smem + AT_XCH_OFF);      // [2][128]
    float* xsum = xmax + 2 * AT_BM;                                 // [2][128]
    mbar_wait(smem_u32(bar_s), 0);
    tc_fence_after();
    float mx = -INFINITY;
    uint32_t v[32];
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int c = half * 4 + cc;
      if (c * 32 >= T) break;
      tmem_ld32(taddr + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < T) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    xmax[half * AT_BM + row] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mx = fmaxf(mx, xmax[(half ^ 1) * AT_BM + row]);
    const float mxs = mx * scale_log2e;
    float sum = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int c = half * 4 + cc;
      uint32_t pk[16];
      if (c * 32 < T) {
        tmem_ld32(taddr + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = (c * 32 + j < T) ? exp2f(fmaf(__uint_as_float(v[j]), scale_log2e, -mxs)) : 0.f;
          float p1 = (c * 32 + j + 1 < T) ? exp2f(fmaf(__uint_as_float(v[j + 1]), scale_log2e, -mxs)) : 0.f;
          // the row sum uses the bf16-rounded probabilities the second MMA will actually consume
          __nv_bfloat162 hp = __floats2bfloat162_rn(p0, p1);
          sum += __low2float(hp) + __high2float(hp);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hp);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
      }
      // keys [32c, 32c+32) of this row -> chunk (c/2) of P, 16-byte slots 4*(c&1) .. +3, XOR-swizzled with (row % 8)
      uint8_t* prow = sP + (c >> 1) * (AT_BM * 128) + row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int slot = ((c & 1) * 4 + q) ^ (row & 7);
        *reinterpret_cast<uint4*>(prow + slot * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    }
    xsum[half * AT_BM + row] = sum;
    // make the generic-proxy smem writes visible to the tensor core (async proxy), and order the TMEM reads before MMA 2
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    mbar_arrive(smem_u32(bar_p));
    mbar_wait(smem_u32(bar_o), 0);   // completes only after all 256 arrivals above, so both partial sums are visible
    tc_fence_after();
    const float inv = 1.f / (sum + xsum[(half ^ 1) * AT_BM + row]);
    const int t = q0 + row;
    {
      const int c = half;
      tmem_ld32(taddr + c * 32, v);
      if (t < T) {
        uint4* o = reinterpret_cast<uint4*>(out + ((int64_t)b * T + t) * E + h * AT_D + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            __nv_bfloat162 hp = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 2 * u]) * inv, __uint_as_float(v[8 * q + 2 * u + 1]) * inv);
            w[u] = *reinterpret_cast<uint32_t*>(&hp);
          }
          o[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AT_CTRL_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
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
  dim3 grid((T + AT_BM - 1) / AT_BM, H, B);
  attn_tc_kernel<<<grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(map_q, map_kv, (__nv_bfloat16*)out, T, H,
                                                             scale * 1.4426950408889634f);
  return check_launch("attn_tc");
}
