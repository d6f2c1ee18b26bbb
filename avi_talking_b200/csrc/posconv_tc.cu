// wav2vec2 positional convolution on tcgen05 with the ACTIVATION SLAB RESIDENT in shared memory (implicit convolution).
// (HF Wav2Vec2PositionalConvEmbedding: Conv1d(768, 768, k = 128, padding = 64, groups = 16)[..., :-1]; reached from
//  models/lib/wav2vec.py:142 through the encoder.)
//
//   pc[b, t, co] = bias[co] + sum_{j < 128} sum_{ci < 48} xpad[b, t + j, 48 g + ci] * w[co, ci, j],   g = co / 48
//
// Round 1 ran this as 4 block-diagonal conv-mode GEMMs: 4x the algorithmic MACs (3/4 structural zeros) and, worse, the 128-row A
// tile of every one of the 128 taps was fetched from L2 again - 0.5 ms per 64-clip step at 0.19 algorithmic efficiency, bound by the
// L2 -> SM operand stream. Here:
//   * work unit = (clip, 256-row time tile, group quad of 192 channels, half of the taps). Each CTA of the pair keeps the
//     (128 + 63)-row x 192-channel slab it needs in shared memory (3 SWIZZLE_128B tiles of 192 rows x 64 channels, loaded ONCE by
//     TMA); tap j is the same slab read from row j: the MMA's shared-memory descriptor simply starts j * 128 bytes further - the
//     128B swizzle is a function of the shared-memory ADDRESS, so any row offset is legal (profiles/probes/desc_rowoffset_probe.cu);
//   * only the non-zero BAND of the block-diagonal weight is contracted: the 64-channel k-block c of a quad meets output columns
//     [48 c, 48 c + 96) only (2 of the 4 groups), so each k-block is ONE tcgen05.mma.cta_group::2 of N = 96 (M = 256 across the
//     pair) into accumulator columns 48 c .. 48 c + 95 - 2x the algorithmic MACs instead of 4x, and the weight stream is 6 KB per
//     CTA per k-block (48 rows), through an 8-stage TMA ring;
//   * the two tap halves of a tile are separate units (keeps 74 pairs evenly loaded: 512 units at 64 clips) that meet in the
//     fp32 output through TMA reduce-add (the output is zeroed by the launch; two commutative additions onto zero: deterministic).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (leader CTA), warps 2..9 = epilogue.
#include <cstdlib>

#include "tc_common.cuh"

namespace avi {

constexpr int PI_BM = 128, PI_NQ = 192, PI_NB = 96, PI_NBH = 48, PI_CG = 48, PI_TAPS = 128, PI_KSPLIT = 2, PI_UT = PI_TAPS / PI_KSPLIT;
constexpr int PI_SLAB_ROWS = 192;                                 // >= PI_BM + PI_UT - 1 = 191 rows, whole 8-row swizzle groups
constexpr uint32_t PI_SLAB_C_BYTES = PI_SLAB_ROWS * 128;          // one 64-channel tile of the slab: 24 KB
constexpr uint32_t PI_SLAB_BYTES = 3 * PI_SLAB_C_BYTES;           // 72 KB
constexpr int PI_WSTAGES = 8;
constexpr uint32_t PI_W_BYTES = PI_NBH * 128;                     // 6 KB: this CTA's 48 rows of the 96-row band
constexpr int PI_EPI_WARPS = 8, PI_THREADS = (2 + PI_EPI_WARPS) * 32;
constexpr uint32_t PI_TRANS_WARP = 32 * 16 * 4;
constexpr uint32_t PI_OFF_W = 2 * PI_SLAB_BYTES;
constexpr uint32_t PI_OFF_TRANS = PI_OFF_W + PI_WSTAGES * PI_W_BYTES;
constexpr uint32_t PI_OFF_BIAS = PI_OFF_TRANS + PI_EPI_WARPS * PI_TRANS_WARP;
constexpr uint32_t PI_OFF_BAR = PI_OFF_BIAS + PI_NQ * 4;
constexpr uint32_t PI_SMEM = PI_OFF_BAR + 256;
static_assert(PI_SMEM <= 232448, "shared memory budget");
static_assert(PI_OFF_W % 1024 == 0 && PI_OFF_TRANS % 1024 == 0 && PI_SLAB_C_BYTES % 1024 == 0 && PI_W_BYTES % 1024 == 0, "swizzle alignment");
static_assert(2 * PI_WSTAGES + 8 <= 31, "barrier block");

struct PosconvParams {
  const float* bias;   // [C]
  int B, T, m_tiles, n_quads, total_units;
  int dbg;             // developer experiments (AVI_PC_DBG, timeline builds only): 1 = every tap reads slab row 0 (wrong result, timing only)
};

#ifdef AVI_GEMM_TIMELINE
// per unit of pair 0: [0] MMA start, [1] MMA end (issue), [2] cycles the MMA thread waited on full[], [3] cycles the producer waited on empty[],
// [4] epilogue start, [5] epilogue end
__device__ long long g_pc_timeline[16][8];
#define PTL(unit, slot, val) do { if (blockIdx.x == 0 && (unit) < 16) g_pc_timeline[unit][slot] = (val); } while (0)
#else
#define PTL(unit, slot, val) do { } while (0)
#endif

__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct PiUnit {
  int h, q, m_blk, b;
};
// consecutive units = consecutive (clip, time tile) of ONE (tap half, quad): the pairs running at the same time stream the same
// 2.4 MB of weights, which then stay L2-resident
__device__ __forceinline__ PiUnit pi_decode(int u, const PosconvParams& p) {
  const int per = p.B * p.m_tiles;
  PiUnit r;
  const int hq = u / per, bm = u % per;
  r.h = hq / p.n_quads;
  r.q = hq % p.n_quads;
  r.b = bm / p.m_tiles;
  r.m_blk = bm % p.m_tiles;
  return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PI_THREADS, 1)
posconv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_c, const PosconvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* slab = smem;                     // [2][3][192 rows x 128 B]
  uint8_t* wring = smem + PI_OFF_W;         // [8][48 rows x 128 B]
  float* trans = reinterpret_cast<float*>(smem + PI_OFF_TRANS);
  float* sbias = reinterpret_cast<float*>(smem + PI_OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PI_OFF_BAR);
  uint64_t* full_bar = bars;                          // [WSTAGES] (leader)
  uint64_t* empty_bar = bars + PI_WSTAGES;            // [WSTAGES]
  uint64_t* slab_full = bars + 2 * PI_WSTAGES;        // [2] (leader)
  uint64_t* slab_empty = bars + 2 * PI_WSTAGES + 2;   // [2]
  uint64_t* tmem_full = bars + 2 * PI_WSTAGES + 4;    // [2]
  uint64_t* tmem_empty = bars + 2 * PI_WSTAGES + 6;   // [2] (leader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * PI_WSTAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < PI_WSTAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&slab_full[s]), 1);
      mbar_init(smem_u32(&slab_empty[s]), 1);
      mbar_init(smem_u32(&tmem_full[s]), 1);
      mbar_init(smem_u32(&tmem_empty[s]), 2 * PI_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr int KB_PER_UNIT = PI_UT * 3;   // 192 k-blocks: 64 taps x 3 channel blocks

  if (warp == 0) {
    // ===================== TMA producer: the slab once per unit, the weight band k-block by k-block =====================
    if (lane == 0) {
      const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[0]), 0);
      const uint32_t slab_full_leader = mapa_shared(smem_u32(&slab_full[0]), 0);
      auto issue_slab = [&](int u, int it) {
        const PiUnit un = pi_decode(u, p);
        const int sb = it & 1;
        mbar_wait(smem_u32(&slab_empty[sb]), ((it >> 1) & 1) ^ 1);
        if (rank == 0) mbar_expect_tx(smem_u32(&slab_full[sb]), 2 * PI_SLAB_BYTES);
        const int row0 = un.m_blk * (2 * PI_BM) + (int)rank * PI_BM + un.h * PI_UT;   // xpad row of (output row 0 of this CTA, first tap)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          tma_load_3d_pair(smem_u32(slab + sb * PI_SLAB_BYTES + c * PI_SLAB_C_BYTES), &map_x, slab_full_leader + sb * 8,
                           un.q * PI_NQ + c * 64, row0, un.b);
      };
      int stage = 0, it = 0;
      uint32_t phase = 0;
      if (pair < p.total_units) issue_slab(pair, 0);
      for (int u = pair; u < p.total_units; u += num_pairs, ++it) {
        const PiUnit un = pi_decode(u, p);
        // weight stream rows: ((q * 128 + j) * 3 + i) * 96 + rank * 48, j = h * 64 + jj
        int wrow = ((un.q * PI_TAPS + un.h * PI_UT) * 3) * PI_NB + (int)rank * PI_NBH;
        long long waited = 0;
        for (int kb = 0; kb < KB_PER_UNIT; ++kb) {
          // the next unit's slab is requested a third of the way through this unit (its buffer was released when the previous unit's
          // MMAs retired), so it lands long before the MMA issuer needs it
          if (kb == KB_PER_UNIT / 3 && u + num_pairs < p.total_units) issue_slab(u + num_pairs, it + 1);
#ifdef AVI_GEMM_TIMELINE
          const long long w0 = clock64();
#endif
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
#ifdef AVI_GEMM_TIMELINE
          waited += clock64() - w0;
#endif
          if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[stage]), 2 * PI_W_BYTES);
          tma_load_2d_pair(smem_u32(wring + stage * PI_W_BYTES), &map_w, full_leader + stage * 8, 0, wrow);
          wrow += PI_NB;
          if (++stage == PI_WSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        PTL(it, 3, waited);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      // D = f32, A = B = bf16, K-major both, N = 96, M = 256 (pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PI_NB >> 3) << 17) | ((uint32_t)((2 * PI_BM) >> 4) << 24);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int u = pair; u < p.total_units; u += num_pairs, ++it) {
        const int as = it & 1, sb = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&tmem_empty[as]), aphase ^ 1);
        mbar_wait(smem_u32(&slab_full[sb]), aphase);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * PI_NQ;
        const uint64_t slab_desc = umma_desc_sw128(smem_u32(slab + sb * PI_SLAB_BYTES));
        long long waited = 0;
        PTL(it, 0, clock64());
        for (int jj = 0; jj < PI_UT; ++jj) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            // stream order of the channel blocks is c = 0, 2, 1: at tap 0 the first two MMAs INITIALISE columns [0,96) and [96,192)
            // (accumulate = 0), the third (c = 1, columns [48,144)) and everything after accumulate
            const int c = (i == 0) ? 0 : (i == 1 ? 2 : 1);
#ifdef AVI_GEMM_TIMELINE
            const long long w0 = clock64();
#endif
            mbar_wait(smem_u32(&full_bar[stage]), phase);
#ifdef AVI_GEMM_TIMELINE
            waited += clock64() - w0;
#endif
            tc_fence_after();
            const int row_off = (p.dbg & 1) ? 0 : jj;
            const uint64_t adesc = slab_desc + (uint64_t)((c * PI_SLAB_C_BYTES + row_off * 128) >> 4);   // tap jj = rows jj.. of the slab
            const uint64_t bdesc = umma_desc_sw128(smem_u32(wring + stage * PI_W_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_pair(tmem_d + PI_CG * c, adesc + 2 * k, bdesc + 2 * k, idesc, (jj == 0 && i < 2 && k == 0) ? 0u : 1u);
            umma_commit_pair(smem_u32(&empty_bar[stage]), 3);
            if (++stage == PI_WSTAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        PTL(it, 1, clock64());
        PTL(it, 2, waited);
        umma_commit_pair(smem_u32(&slab_empty[sb]), 3);   // both CTAs' slabs may be overwritten once these MMAs have read them
        umma_commit_pair(smem_u32(&tmem_full[as]), 3);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> (+ bias) -> SWIZZLE_64B staging tile -> TMA reduce-add into the fp32 output =====================
    const int ew = warp - 2;             // 0..7
    const int quarter = warp & 3;        // TMEM lanes [32 * quarter, +32)
    const int half = ew >> 2;            // accumulator columns [96 * half, +96)
    const int etid = threadIdx.x - 64;
    const uint32_t tile = smem_u32(trans) + ew * PI_TRANS_WARP;
    const uint32_t te_leader0 = mapa_shared(smem_u32(&tmem_empty[0]), 0);
    const uint32_t te_leader1 = mapa_shared(smem_u32(&tmem_empty[1]), 0);
    int it = 0;
    for (int u = pair; u < p.total_units; u += num_pairs, ++it) {
      const PiUnit un = pi_decode(u, p);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      // bias of this quad (added by the first tap half only); safe to overwrite: every warp finished reading it before it arrived
      // on tmem_empty of the previous unit... which the bar.sync below orders
      asm volatile("bar.sync 1, %0;" ::"n"(PI_EPI_WARPS * 32) : "memory");
      if (etid < PI_NQ) sts32(smem_u32(sbias) + etid * 4, __float_as_uint(un.h == 0 ? __ldg(p.bias + un.q * PI_NQ + etid) : 0.f));
      asm volatile("bar.sync 1, %0;" ::"n"(PI_EPI_WARPS * 32) : "memory");
      if (lane == 0) mbar_wait(smem_u32(&tmem_full[as]), aphase);
      __syncwarp();
      tc_fence_after();
      if (ew == 0 && lane == 0) PTL(it, 4, clock64());
      const int row_base = un.m_blk * (2 * PI_BM) + (int)rank * PI_BM + quarter * 32;
      if (row_base < p.T) {
#pragma unroll 1
        for (int ch = 0; ch < 3; ++ch) {
          const int col0 = half * PI_NB + ch * 32;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * PI_NQ + col0), v);
          const uint32_t wr_row = tile + lane * 64, wr_sw = (lane >> 1) & 3;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
              const float4 bb = lds128f(smem_u32(sbias) + (col0 + 16 * hh + 4 * gg) * 4);
              sts128(wr_row + 16 * (gg ^ wr_sw), __float_as_uint(__uint_as_float(v[16 * hh + 4 * gg]) + bb.x),
                     __float_as_uint(__uint_as_float(v[16 * hh + 4 * gg + 1]) + bb.y),
                     __float_as_uint(__uint_as_float(v[16 * hh + 4 * gg + 2]) + bb.z),
                     __float_as_uint(__uint_as_float(v[16 * hh + 4 * gg + 3]) + bb.w));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_reduce_add_3d(tile, &map_c, un.q * PI_NQ + col0 + 16 * hh, row_base, un.b);   // rows >= T are clipped by the TMA unit
              bulk_commit();
            }
          }
        }
      }
      if (ew == 0 && lane == 0) PTL(it, 5, clock64());
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(as ? te_leader1 : te_leader0);
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace avi

using namespace avi;

#ifdef AVI_GEMM_TIMELINE
extern "C" int avi_debug_posconv_timeline(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_pc_timeline, sizeof(long long) * 16 * 8) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int avi_w2v_posconv_tc(const void* xpad, const void* w_band, const float* bias, float* pc, int32_t B, int32_t T, int32_t Tp,
                                  int32_t C, int32_t groups, int32_t k, void* stream) {
  AVI_REQUIRE(xpad && w_band && bias && pc, "avi_w2v_posconv_tc: null pointer");
  AVI_REQUIRE(B > 0 && T > 0 && k == PI_TAPS && groups > 0 && C % groups == 0 && C / groups == PI_CG && groups % 4 == 0,
              "avi_w2v_posconv_tc: built for k = 128 and 48-channel groups in quads (C=%d groups=%d k=%d)", C, groups, k);
  AVI_REQUIRE(Tp >= T + k - 1, "avi_w2v_posconv_tc: xpad needs T + k - 1 rows per clip (Tp=%d T=%d)", Tp, T);
  AVI_REQUIRE(((uintptr_t)xpad | (uintptr_t)w_band | (uintptr_t)pc) % 16 == 0, "avi_w2v_posconv_tc: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int n_quads = groups / 4;
  CUtensorMap map_x, map_w, map_c;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)Tp, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)Tp * C * 2};
    uint32_t box[3] = {64, PI_SLAB_ROWS, 1};
    if (encode_map(&map_x, xpad, 3, dims, strides, box)) return 1;
  }
  {
    uint64_t dims[2] = {64, (uint64_t)n_quads * PI_TAPS * 3 * PI_NB};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, PI_NBH};
    if (encode_map(&map_w, w_band, 2, dims, strides, box)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)T * C * 4};
    uint32_t box[3] = {16, 32, 1};
    if (encode_map(&map_c, pc, 3, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
  }
  PosconvParams p;
  p.bias = bias;
  p.B = B;
  p.T = T;
  p.m_tiles = (T + 2 * PI_BM - 1) / (2 * PI_BM);
  p.n_quads = n_quads;
  p.total_units = PI_KSPLIT * n_quads * B * p.m_tiles;
  p.dbg = 0;
#ifdef AVI_GEMM_TIMELINE
  if (const char* e = getenv("AVI_PC_DBG")) p.dbg = atoi(e);
#endif
  // the tap halves meet in the output through reduce-add
  cudaError_t me = cudaMemsetAsync(pc, 0, (size_t)B * T * C * sizeof(float), st);
  AVI_REQUIRE(me == cudaSuccess, "avi_w2v_posconv_tc: cudaMemsetAsync: %s", cudaGetErrorString(me));
  static SmemOptIn optin;
  const cudaError_t attr_err = smem_optin(posconv_tc_kernel, (int)PI_SMEM, optin);
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_w2v_posconv_tc: cannot opt in to %u bytes of shared memory: %s", PI_SMEM, cudaGetErrorString(attr_err));
  const int max_pairs = device_sms() / 2;
  const int pairs = p.total_units < max_pairs ? p.total_units : max_pairs;
  posconv_tc_kernel<<<2 * pairs, PI_THREADS, PI_SMEM, st>>>(map_x, map_w, map_c, p);
  return check_launch("posconv_tc");
}
