// bf16 tensor-core GEMM for sm_100a on CTA PAIRS: tcgen05.mma.cta_group::2 (M=256 across two SMs, N=256, K=16) with the
// accumulator in TMEM, operands staged in shared memory by TMA (SWIZZLE_128B, K-major) through a 4-stage mbarrier ring, two
// TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1, persistent over the 74 SM pairs.
//
//   C[b, r, n] = epi( sum_k A[b, r, k] * W[n, k] + bias[n] ) (+ residual[b, r, n])
//
// Why pairs: the single-CTA 128x256 tile needs 48 KB of operands per 4.2 MFLOP and was pinned at the L2->SM delivery rate
// (~46 B/clk/SM measured, profiles/); in a pair each CTA stages its own 128 rows of A but only HALF of the W tile (128 of the
// 256 rows), 32 KB per 4.2 MFLOP, and the MMA reads both halves.
//
// A is addressed through a 4-D tensor map (channel, phase, super-row, clip) so that a strided Conv1d over time-major
// activations is the same code path as a plain GEMM: output row r, tap j reads input row r*stride + j = super-row
// (r + j/stride), phase (j % stride).
//
// Warp roles (576 threads per CTA): warp 0 = TMA producer (both CTAs), warp 1 = TMEM allocator (both) + MMA issuer (leader
// CTA only), warps 2..17 = epilogue (TMEM lane quarter = warp_id % 4, 64-column slice = (warp_id - 2) / 4).
// Barriers: full[s] lives in the leader (both CTAs' TMA bytes land on it), empty[s] / tmem_full[a] are multicast to both CTAs
// by tcgen05.commit, tmem_empty[a] lives in the leader and collects the epilogue warps of both CTAs.
//
// CL = 4 (two pairs per cluster; OPT-IN through avi_gemm_set_multicast(1) / AVI_GEMM_MULTICAST=1): the mainloop of a pair is paced
// by L2 -> SM delivery, not by the tensor pipe (32 KB per CTA per k-block arrive in ~800 clk against 512 clk of MMAs: ~40 B/clk/SM,
// the L2 slice output cap of the chip). The two pairs of a cluster work on two consecutive m-tiles of the SAME n-tile, and each
// CTA fetches only a 64-row QUARTER of the W tile, multicast to the CTA of the other pair that needs the same half: 24 KB per CTA
// per k-block are requested instead of 32. A slot is reusable once BOTH pairs' MMAs have read it (empty[s] counts two commits,
// each multicast to all four CTAs), so the pairs run in lockstep. MEASURED (profiles/r2/gemm_multicast_ab.txt): bit-identical
// results, but no gain - the whole step is unchanged (GEMM launches 5.43 vs 5.44 ms) and isolated L2-resident encoder shapes are
// 7-13 % SLOWER (qkv 62 vs 55 us): 2-way multicast does not lift the delivery cap (the L2 already merges near-simultaneous
// requests for a line from a few SMs) and the lockstep costs slack. Hence off by default; kept, tested, for wider clusters.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "tc_common.cuh"

namespace avi {

constexpr int P2_BM = 128, P2_BN = 256, P2_BNH = 128, P2_BK = 64, P2_STAGES = 6, P2_UMMA_K = 16;
constexpr int P2_EPI_WARPS = 16, P2_THREADS = (2 + P2_EPI_WARPS) * 32, P2_EPI_THREADS = P2_EPI_WARPS * 32;
constexpr uint32_t P2_A_BYTES = P2_BM * P2_BK * 2;    // 16 KB
constexpr uint32_t P2_B_BYTES = P2_BNH * P2_BK * 2;   // 16 KB
constexpr uint32_t P2_STAGE_BYTES = P2_A_BYTES + P2_B_BYTES;
// The operand ring is what hides the TMA round trip: with 4 stages both this kernel and its single-CTA predecessor sat at
// ~980 clk per 64-deep k-block (MMA floor 512), i.e. latency-bound; the pair's smaller stages (32 KB instead of 48 KB) buy 6.
constexpr uint32_t P2_TRANS_WARP = 32 * 16 * 4;                  // per-warp staging tile: 32 rows x 16 words (XOR-swizzled)
constexpr uint32_t P2_TRANS_BYTES = P2_EPI_WARPS * P2_TRANS_WARP;
constexpr uint32_t P2_BIAS_BYTES = 2 * P2_BN * 4;                // double-buffered by accumulator stage
constexpr uint32_t P2_OFF_B = P2_STAGES * P2_A_BYTES;
constexpr uint32_t P2_OFF_TRANS = P2_STAGES * P2_STAGE_BYTES;
constexpr uint32_t P2_OFF_BIAS = P2_OFF_TRANS + P2_TRANS_BYTES;
constexpr uint32_t P2_OFF_BAR = P2_OFF_BIAS + P2_BIAS_BYTES;
constexpr int P2_CLC_STAGES = 4;                                 // cluster-launch-control responses in flight (dynamic tile scheduler)
constexpr uint32_t P2_OFF_CLC = P2_OFF_BAR + 256 /*barriers*/;   // [P2_CLC_STAGES] 16-byte responses
constexpr uint32_t P2_SMEM_BYTES = P2_OFF_CLC + P2_CLC_STAGES * 16;
constexpr int P2_CLC_CONSUMERS = 2 * (1 + P2_EPI_WARPS) + 1;     // per pair: two producer threads, the MMA thread, 2 x 16 epilogue warps
static_assert(P2_SMEM_BYTES <= 232448, "shared memory budget");
static_assert(2 * P2_STAGES + 6 + 2 * P2_CLC_STAGES <= 32, "barrier block");

#ifdef AVI_GEMM_TIMELINE
// Build-time instrumentation (python __graft_entry__.py with AVI_NVCC_EXTRA=-DAVI_GEMM_TIMELINE): clock64() stamps of CTA 0's
// roles for the first tiles, read back with avi_debug_timeline(). Not compiled into the product library.
__device__ long long g_timeline[3][64][8];
#define TL(role, tile, slot) do { if (blockIdx.x == 0 && (tile) < 64) g_timeline[role][tile][slot] = clock64(); } while (0)
#else
#define TL(role, tile, slot) do { } while (0)
#endif

struct Tc2Params {
  const float* bias;
  const float* residual;
  void* C;
  void* C2;
  int c_dtype;  // dtype of C; C2 (if any) is the other one
  int act;
  int batch, rows, N;
  int num_kb;      // K / 64
  int kb_per_tap;  // C_in / 64
  int conv_stride;
  int taps_x, row_skip;  // 2-D taps: after every taps_x taps the input row advances by row_skip more (row pitch - taps_x); 0 = 1-D
  int m_tiles, n_tiles, total_tiles;  // m_tiles counts 256-row pair tiles per batch entry
  int64_t c_ld, c_batch_stride, res_ld, res_batch_stride;
  int vec_ok;      // outputs (and residual) are 16-byte addressable per 32-column chunk
  int tma_store;   // 0: threads store; 1: staged tiles leave through TMA stores (map_c); 2: TMA reduce-add (C += ..., residual == C)
  // 1 (CL = 2 only): the grid holds ONE cluster per tile and the resident pairs steal the pending ones through cluster launch
  // control (clusterlaunchcontrol.try_cancel) instead of walking tile = pair + k * pairs. The pairs then share the tiles by
  // availability: a launch that only gets part of the GPU (the other batch's autoregressive decoder holds 64 SMs for 1 ms) finishes
  // on the SMs it has instead of leaving the tiles of its not-yet-resident pairs for when they arrive.
  int dynamic;
};

// TF32 = true: operands are fp32 in shared memory (32 elements per 128-byte swizzle row instead of 64), consumed by
// tcgen05.mma.kind::tf32 (10-bit significand, fp32 accumulate): the same pipeline at half the MMA rate, for callers that need more
// than bf16's 8 bits (the 60-convolution FanEncoder). Everything below is written in BYTES per k-block (128) and in elements via BK.
template <bool TF32, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(P2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_c, const Tc2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // no static shared memory in this kernel: the dynamic window starts aligned
  if ((smem_u32(smem) & 1023u) != 0) __trap();        // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + P2_OFF_B;
  float* trans = reinterpret_cast<float*>(smem + P2_OFF_TRANS);
  float* sbias = reinterpret_cast<float*>(smem + P2_OFF_BIAS);  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P2_OFF_BAR);
  uint64_t* full_bar = bars;                        // [STAGES]  (used in the leader CTA)
  uint64_t* empty_bar = bars + P2_STAGES;           // [STAGES]
  uint64_t* tmem_full = bars + 2 * P2_STAGES;       // [2]
  uint64_t* tmem_empty = bars + 2 * P2_STAGES + 2;  // [2]       (used in the leader CTA)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * P2_STAGES + 4);
  uint64_t* clc_full = bars + 2 * P2_STAGES + 6;                    // [CLC_STAGES] response landed (every CTA of the pair)
  uint64_t* clc_empty = clc_full + P2_CLC_STAGES;                   // [CLC_STAGES] every consumer of the pair has read it (leader)
  uint8_t* clc_resp = smem + P2_OFF_CLC;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();             // 0 .. CL-1
  const uint32_t rank = crank & 1u;                     // position inside the CTA pair (0 = leader: issues the MMAs)
  const uint32_t leader = crank & ~1u;                  // cluster rank of this pair's leader
  const int q = (int)(crank >> 1);                      // which pair of the cluster (CL = 4: the pair's m-tile inside the super-tile)
  constexpr int PPC = CL / 2;                           // pairs per cluster
  const int pair = blockIdx.x / CL, num_pairs = gridDim.x / CL;   // work-item walkers (a cluster walks super-tiles of PPC m-tiles)
  const int mt_total = p.m_tiles * p.batch;
  const bool dyn = CL == 2 && p.dynamic != 0;
  const int first_tile = dyn ? (int)(blockIdx.x / CL) : pair;
  const uint32_t clc_empty_leader = mapa_shared(smem_u32(&clc_empty[0]), leader);
  // next tile of this pair: the static walk, or the index of a cluster taken back from the launch queue (-1: none left). Every role
  // of both CTAs consumes every response (its own slot / phase cursor) and reports to the leader's empty barrier.
  auto next_tile = [&](int t, int& cslot, uint32_t& cphase) -> int {
    if (!dyn) {
      t += num_pairs;
      return t < p.total_tiles ? t : -1;
    }
    mbar_wait(smem_u32(&clc_full[cslot]), cphase);
    const int nxt = clc_decode(smem_u32(clc_resp + 16 * cslot));
    fence_proxy_async_smem();                      // the slot is rewritten by the async proxy (the next try_cancel)
    mbar_arrive_cluster(clc_empty_leader + 8 * cslot);
    if (++cslot == P2_CLC_STAGES) {
      cslot = 0;
      cphase ^= 1;
    }
    return nxt < 0 ? -1 : nxt / CL;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < P2_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), PPC);   // one tcgen05.commit per pair of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tmem_full[s]), 1);
      mbar_init(smem_u32(&tmem_empty[s]), 2 * P2_EPI_WARPS);
    }
    for (int s = 0; s < P2_CLC_STAGES; ++s) {
      mbar_init(smem_u32(&clc_full[s]), 1);
      mbar_init(smem_u32(&clc_empty[s]), P2_CLC_CONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp per CTA: all 512 TMEM columns of both SMs (2 accumulator stages x 256)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation of BOTH CTAs visible before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();          // everything above may overlap the tail of the previous kernel of the stream (programmatic dependent launch)

  if (warp == 0) {
    // ===================== TMA producer (each CTA loads its 128 rows of A and its 128 rows of W) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[0]), leader);
      int cslot = 0, islot = 0, it = 0;
      uint32_t cphase = 0, iphase = 0;
      // the leader's producer thread also asks for the NEXT tile when it starts loading one (one request ahead: the answer is there
      // long before the k-blocks of the current tile have been issued, and a pair never reserves more than one tile beyond its own)
      auto clc_issue = [&]() {
        mbar_wait(smem_u32(&clc_empty[islot]), iphase ^ 1);
        mbar_expect_tx(smem_u32(&clc_full[islot]), 16);
        mbar_expect_tx_cluster(mapa_shared(smem_u32(&clc_full[islot]), 1), 16);
        clc_try_cancel_multicast(smem_u32(clc_resp + 16 * islot), smem_u32(&clc_full[islot]));
        if (++islot == P2_CLC_STAGES) {
          islot = 0;
          iphase ^= 1;
        }
      };
      if (dyn && crank == 0) clc_issue();
      for (int t = first_tile; t >= 0; ++it) {
        const int n_blk = t % p.n_tiles;
        const int mt = (t / p.n_tiles) * PPC + q;     // past the end for the idle pair of an odd last super-tile: TMA zero-fills
        const int b = mt / p.m_tiles, m_blk = mt % p.m_tiles;
        const int row0 = m_blk * (2 * P2_BM) + (int)rank * P2_BM;
        // the last n-tile may be narrower than 256: the MMA is issued with N = n_eff (a multiple of 32), of which each CTA of the pair
        // supplies n_eff / 2 rows of W (the box still brings 128 rows; the surplus is not read)
        const int n_eff = min(P2_BN, ((p.N - n_blk * P2_BN + 31) >> 5) << 5);
        const int wrow0 = n_blk * P2_BN + (int)rank * (n_eff >> 1);
        // tap / phase / super-row / channel-block counters advance incrementally: this single thread paces the whole pipeline,
        // and four runtime integer divisions per k-block cost about as much as the MMAs of that k-block
        int kin = 0, ph = 0, sr = 0, tx = 0;
        TL(0, it, 0);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          if (kb == 0) TL(0, it, 1);
          const uint32_t fb_local = smem_u32(&full_bar[stage]);
          if (rank == 0) mbar_expect_tx(fb_local, 2 * P2_STAGE_BYTES);
          constexpr int BK = TF32 ? P2_BK / 2 : P2_BK;   // elements per 128-byte k-block row
          tma_load_4d_pair(smem_u32(smem_a + stage * P2_A_BYTES), &map_a, full_leader + stage * 8, kin * BK, ph, row0 + sr, b);
          if constexpr (CL == 4) {
            // this CTA's quarter (64 rows) of the W tile, delivered to the same offset in both CTAs that hold this half
            tma_load_2d_pair_mc(smem_u32(smem_b + stage * P2_B_BYTES + q * (P2_B_BYTES / 2)), &map_w, full_leader + stage * 8, kb * BK,
                                wrow0 + q * (P2_BNH / 2), (uint16_t)(0x5u << rank));
          } else {
            tma_load_2d_pair(smem_u32(smem_b + stage * P2_B_BYTES), &map_w, full_leader + stage * 8, kb * BK, wrow0);
          }
          if (++kin == p.kb_per_tap) {
            kin = 0;
            if (++ph == p.conv_stride) {
              ph = 0;
              ++sr;
            }
            if (p.taps_x != 0 && ++tx == p.taps_x) {   // next line of a k x k window
              tx = 0;
              sr += p.row_skip;
            }
          }
          if (++stage == P2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        t = next_tile(t, cslot, cphase);
        if (dyn && crank == 0 && t >= 0) clc_issue();   // never after a failed request
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=n_eff (256 except in a narrower last n-tile), M=256 (pair)
      // (formats: kind::f16 -> 1 = bf16; kind::tf32 -> 2 = tf32)
      const uint32_t fmt = TF32 ? 2u : 1u;
      const uint32_t idesc_base = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)((2 * P2_BM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0, cslot = 0;
      uint32_t cphase = 0;
      for (int t = first_tile; t >= 0; t = next_tile(t, cslot, cphase), ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int n_eff = min(P2_BN, ((p.N - (t % p.n_tiles) * P2_BN + 31) >> 5) << 5);
        const uint32_t idesc = idesc_base | ((uint32_t)(n_eff >> 3) << 17);
        TL(1, it, 0);
        mbar_wait(smem_u32(&tmem_empty[as]), aphase ^ 1);
        TL(1, it, 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * P2_BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          if (kb == 0) TL(1, it, 2);
          if (kb == p.num_kb - 1) TL(1, it, 3);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * P2_A_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * P2_B_BYTES));
#pragma unroll
          for (int k = 0; k < P2_BK / P2_UMMA_K; ++k) {   // four MMAs of 32 bytes of K each (16 bf16 or 8 tf32 elements)
            if constexpr (TF32) umma_tf32_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(smem_u32(&empty_bar[stage]), (uint16_t)((1u << CL) - 1));  // one of the PPC arrivals that free the slot in every CTA of the cluster
          if (++stage == P2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(smem_u32(&tmem_full[as]), (uint16_t)(3u << leader));  // accumulator complete -> epilogues of both CTAs of the pair
        TL(1, it, 4);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> smem transpose -> coalesced global =====================
    // Phase A (thread = accumulator row): tcgen05.ld 32 columns, + bias, activation, written to a per-warp staging tile.
    // Phase B (lanes span the columns of a few rows): 16-byte shared loads and fully coalesced 16-byte global accesses.
    // Three variants: packed bf16 rows, fp32 rows (+ residual), and a scalar one for unaligned / ragged outputs (the
    // 15069-wide vertex rows).  The XOR swizzles keep every shared access conflict-free.
    const int ew = warp - 2;              // 0..15
    const int quarter = warp & 3;         // TMEM lanes [32*quarter, +32) are the only ones this warp may touch
    const int slice = ew >> 2;            // columns [64*slice, +64)
    const int etid = threadIdx.x - 64;    // 0..511
    const uint32_t tile = smem_u32(trans) + ew * P2_TRANS_WARP;
    const bool fast_bf16 = p.vec_ok && p.c_dtype == AVI_DT_BF16 && p.C2 == nullptr && p.residual == nullptr;
    const bool fast_f32 = p.vec_ok && p.c_dtype == AVI_DT_F32 && p.C2 == nullptr;
    const bool ragged_f32 = !p.vec_ok && p.c_dtype == AVI_DT_F32 && p.C2 == nullptr && p.residual == nullptr;
    const uint32_t te_leader0 = mapa_shared(smem_u32(&tmem_empty[0]), leader);
    const uint32_t te_leader1 = mapa_shared(smem_u32(&tmem_empty[1]), leader);
    int it = 0, cslot = 0;
    uint32_t cphase = 0;
    for (int t = first_tile; t >= 0; ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int n_blk = t % p.n_tiles;
      const int mt = (t / p.n_tiles) * PPC + q;
      const int b = mt / p.m_tiles, m_blk = mt % p.m_tiles;
      const int n0 = n_blk * P2_BN;
      const uint32_t sbias_a = smem_u32(sbias) + as * (P2_BN * 4);
      if (ew == 0 && lane == 0) TL(2, it, 0);
      if (etid < P2_BN) {
        const int n = n0 + etid;
        sts32(sbias_a + etid * 4, __float_as_uint((p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(P2_EPI_THREADS) : "memory");
      if (ew == 0 && lane == 0) TL(2, it, 1);
      if (lane == 0) mbar_wait(smem_u32(&tmem_full[as]), aphase);   // one polling lane per warp: 16 pollers on the barrier, not 512
      __syncwarp();
      if (ew == 0 && lane == 0) TL(2, it, 2);
      tc_fence_after();
      const int row_base = m_blk * (2 * P2_BM) + (int)rank * P2_BM + quarter * 32;
      const int rows_valid = mt < mt_total ? p.rows - row_base : 0;  // may be <= 0 or > 32 (0: the idle pair of an odd last super-tile)
      const int64_t c_row0 = (int64_t)b * p.c_batch_stride + (int64_t)row_base * p.c_ld;
      const int64_t r_row0 = (int64_t)b * p.res_batch_stride + (int64_t)row_base * p.res_ld;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const int col0 = slice * 64 + ch * 32;
        const int n_base = n0 + col0;
        if (n_base >= p.N || rows_valid <= 0) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * P2_BN + col0), v);
        if (ew == 0 && lane == 0) TL(2, it, 3 + 2 * ch);
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = lds128f(sbias_a + (col0 + 4 * j) * 4);   // packed fp32 adds (FADD2): two columns per instruction
          unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), pack_f32x2(bb.x, bb.y)),
                       f[4 * j + 0], f[4 * j + 1]);
          unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pack_f32x2(bb.z, bb.w)),
                       f[4 * j + 2], f[4 * j + 3]);
        }
        if (p.act == AVI_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_fast2(f[j], f[j + 1], f[j], f[j + 1]);   // packed fp32 (FFMA2): two columns per instruction
        } else if (p.act == AVI_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act == AVI_ACT_QUICK_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
        }
        const bool full_n = n_base + 32 <= p.N;
        // staging tile: 32 rows x 16 words (pitch 16); the 4-word group g of row r lives at group g ^ ((r >> 1) & 3), which
        // keeps the row-per-lane writes and the row-segment reads below conflict-free
        const uint32_t wr_row = tile + lane * 64, wr_sw = (lane >> 1) & 3;
        if (p.tma_store) {
          // The staging tile (32 rows x 64 bytes, 16-byte chunk g of row r at g ^ ((r >> 1) & 3)) IS the SWIZZLE_64B shared-memory
          // box of map_c: one thread hands it to the TMA unit, which clips rows >= p.rows and columns >= p.N. No shared loads,
          // no global store instructions: the LSU pipe carries the st.shared traffic only.
          if (p.c_dtype == AVI_DT_BF16) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t w[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * g + 2 * u], f[8 * g + 2 * u + 1]);
                w[u] = *reinterpret_cast<uint32_t*>(&h2);
              }
              sts128(wr_row + 16 * (g ^ wr_sw), w[0], w[1], w[2], w[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(tile, &map_c, n_base, row_base, b);
              bulk_commit();
            }
          } else {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (n_base + 16 * hh >= p.N) break;
              if (lane == 0) bulk_wait_read0();
              __syncwarp();
#pragma unroll
              for (int gg = 0; gg < 4; ++gg)
                sts128(wr_row + 16 * (gg ^ wr_sw), __float_as_uint(f[16 * hh + 4 * gg]), __float_as_uint(f[16 * hh + 4 * gg + 1]),
                       __float_as_uint(f[16 * hh + 4 * gg + 2]), __float_as_uint(f[16 * hh + 4 * gg + 3]));
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (p.tma_store == 2) tma_reduce_add_3d(tile, &map_c, n_base + 16 * hh, row_base, b);
                else tma_store_3d(tile, &map_c, n_base + 16 * hh, row_base, b);
                bulk_commit();
              }
            }
          }
        } else if (fast_bf16 && full_n) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * g + 2 * u], f[8 * g + 2 * u + 1]);
              w[u] = *reinterpret_cast<uint32_t*>(&h2);
            }
            sts128(wr_row + 16 * (g ^ wr_sw), w[0], w[1], w[2], w[3]);
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
          float* cf = reinterpret_cast<float*>(p.C) + c_row0 + n_base;
          const float* rp = p.residual ? p.residual + r_row0 + n_base : nullptr;
          const int g = lane & 3;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {  // two passes of 16 columns
            float4 rr[4];
#pragma unroll
            for (int i8 = 0; i8 < 4; ++i8) {
              const int r = i8 * 8 + (lane >> 2);
              rr[i8] = (rp != nullptr && r < rows_valid)
                           ? __ldg(reinterpret_cast<const float4*>(rp + (int64_t)r * p.res_ld + 16 * hh + 4 * g))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int gg = 0; gg < 4; ++gg)
              sts128(wr_row + 16 * (gg ^ wr_sw), __float_as_uint(f[16 * hh + 4 * gg]), __float_as_uint(f[16 * hh + 4 * gg + 1]),
                     __float_as_uint(f[16 * hh + 4 * gg + 2]), __float_as_uint(f[16 * hh + 4 * gg + 3]));
            __syncwarp();
#pragma unroll
            for (int i8 = 0; i8 < 4; ++i8) {
              const int r = i8 * 8 + (lane >> 2);
              if (r < rows_valid) {
                float4 q = lds128f(tile + (r * 16 + 4 * (g ^ ((r >> 1) & 3))) * 4);
                q.x += rr[i8].x;
                q.y += rr[i8].y;
                q.z += rr[i8].z;
                q.w += rr[i8].w;
                *reinterpret_cast<float4*>(cf + (int64_t)r * p.c_ld + 16 * hh + 4 * g) = q;
              }
            }
            __syncwarp();
          }
        } else if (ragged_f32) {
          // fp32 rows that are only 4-byte aligned (the 15069-wide vertex rows), no residual, no second copy: half a warp per row,
          // 64-byte row segments; pointers advance incrementally and the staging addresses are precomputed, so one pair of rows
          // costs one shared load, one store and one pointer update per lane (this path is HBM-write bound, not issue bound)
          const int c = lane & 15, rsub = lane >> 4;
          const uint32_t rd_base = tile + (rsub * 16 + (c & 3)) * 4;
          uint32_t xo[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) xo[k] = 16u * (uint32_t)((c >> 2) ^ k);
          const int64_t step2 = 2 * p.c_ld;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int gg = 0; gg < 4; ++gg)
              sts128(wr_row + 16 * (gg ^ wr_sw), __float_as_uint(f[16 * hh + 4 * gg]), __float_as_uint(f[16 * hh + 4 * gg + 1]),
                     __float_as_uint(f[16 * hh + 4 * gg + 2]), __float_as_uint(f[16 * hh + 4 * gg + 3]));
            __syncwarp();
            const bool col_ok = n_base + 16 * hh + c < p.N;
            float* cp = reinterpret_cast<float*>(p.C) + c_row0 + n_base + 16 * hh + c + (int64_t)rsub * p.c_ld;
            float val[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) val[u] = __uint_as_float(lds32(rd_base + u * 128 + xo[u & 3]));
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              if (col_ok && 2 * u + rsub < rows_valid) *cp = val[u];
              cp += step2;
            }
            __syncwarp();
          }
        } else {
          // generic (ragged / unaligned rows, residual, second output copy): half a warp per row, 64-byte row segments, scalar accesses
          float* cf = nullptr;
          __nv_bfloat16* cb = nullptr;
          if (p.c_dtype == AVI_DT_F32) {
            cf = reinterpret_cast<float*>(p.C);
            cb = reinterpret_cast<__nv_bfloat16*>(p.C2);
          } else {
            cb = reinterpret_cast<__nv_bfloat16*>(p.C);
            cf = reinterpret_cast<float*>(p.C2);
          }
          const float* rp = p.residual;
          const int c = lane & 15, rsub = lane >> 4;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int gg = 0; gg < 4; ++gg)
              sts128(wr_row + 16 * (gg ^ wr_sw), __float_as_uint(f[16 * hh + 4 * gg]), __float_as_uint(f[16 * hh + 4 * gg + 1]),
                     __float_as_uint(f[16 * hh + 4 * gg + 2]), __float_as_uint(f[16 * hh + 4 * gg + 3]));
            __syncwarp();
            const int n = n_base + 16 * hh + c;
            if (n < p.N) {
#pragma unroll
              for (int r8 = 0; r8 < 2; ++r8) {
                float val[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const int r = 2 * (8 * r8 + u) + rsub;
                  val[u] = __uint_as_float(lds32(tile + (r * 16 + 4 * ((c >> 2) ^ ((r >> 1) & 3)) + (c & 3)) * 4));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const int r = 2 * (8 * r8 + u) + rsub;
                  if (r < rows_valid) {
                    const int64_t co = c_row0 + (int64_t)r * p.c_ld + n;
                    float x = val[u];
                    if (rp) x += __ldg(rp + r_row0 + (int64_t)r * p.res_ld + n);
                    if (cf) cf[co] = x;
                    if (cb) cb[co] = __float2bfloat16_rn(x);
                  }
                }
              }
            }
            __syncwarp();
          }
        }
      }
      if (ew == 0 && lane == 0) TL(2, it, 7);
      // release the accumulator stage back to the MMA warp of the leader CTA
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(as ? te_leader1 : te_leader0);
      int nt = 0;
      if (lane == 0) nt = next_tile(t, cslot, cphase);   // one consumer per warp
      t = __shfl_sync(0xffffffffu, nt, 0);
    }
    if (p.tma_store && lane == 0) bulk_wait_all();   // every store of this thread's bulk groups has been performed
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's shared memory / TMEM / barriers stay valid until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

static const char* tc2_check(const AviGemmArgs* a, bool tf32 = false) {
  if (!a) return "null args";
  if (a->a_dtype != (tf32 ? AVI_DT_F32 : AVI_DT_BF16)) return tf32 ? "A/W must be fp32" : "A/W must be bf16";
  if (a->batch <= 0 || a->rows <= 0 || a->N <= 0 || a->K <= 0) return "bad shape";
  if (a->conv_taps < 1 || a->conv_stride < 1) return "bad conv params";
  if (a->K % a->conv_taps != 0) return "K must be taps * C_in";
  const int cin = a->K / a->conv_taps;
  const int bk = tf32 ? P2_BK / 2 : P2_BK, al = tf32 ? 4 : 8;   // elements per k-block; elements per 16 bytes
  if (cin % bk != 0) return tf32 ? "C_in (K per tap) must be a multiple of 32" : "C_in (K per tap) must be a multiple of 64";
  if (a->a_ld < cin) return "a_ld smaller than the channel window";
  if (a->a_ld % al != 0 || a->a_batch_stride % al != 0) return "a_ld and a_batch_stride must be multiples of 16 bytes";
  if (((uintptr_t)a->A | (uintptr_t)a->W) % 16 != 0) return "A and W must be 16-byte aligned";
  if (a->conv_taps_x != 0) {
    if (a->conv_taps_x < 0 || a->conv_stride != 1 || a->conv_taps % a->conv_taps_x != 0 || a->conv_row_pitch < a->conv_taps_x)
      return "2-D taps need conv_stride 1, conv_taps a multiple of conv_taps_x and conv_row_pitch >= conv_taps_x";
    if (a->a_rows_alloc < (int64_t)(a->rows - 1) + (int64_t)(a->conv_taps / a->conv_taps_x - 1) * a->conv_row_pitch + a->conv_taps_x)
      return "a_rows_alloc smaller than the rows read";
  } else if (a->a_rows_alloc < (int64_t)(a->rows - 1) * a->conv_stride + a->conv_taps) {
    return "a_rows_alloc smaller than the rows read";
  } else if (a->a_rows_alloc % a->conv_stride != 0) {
    // the A tensor map counts a_rows_alloc / conv_stride super-rows: a partial last super-row would be zero-filled by TMA
    return "a_rows_alloc must be a multiple of conv_stride (pad the time axis)";
  }
  if (a->C == nullptr) return "C is null";
  return nullptr;
}

}  // namespace avi

using namespace avi;

static std::atomic<int> g_gemm_multicast{getenv("AVI_GEMM_MULTICAST") != nullptr ? 1 : 0};


#ifdef AVI_GEMM_TIMELINE
extern "C" int avi_debug_timeline(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 3 * 64 * 8) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int avi_gemm_set_multicast(int on) {
  return g_gemm_multicast.exchange(on ? 1 : 0, std::memory_order_relaxed);
}

extern "C" int avi_gemm_bf16_tc_supported(const AviGemmArgs* a) { return tc2_check(a) == nullptr ? 1 : 0; }

// how many 4-CTA clusters of this kernel the device can hold at once (GPCs whose SM count is not a multiple of 4 strand SMs)
template <bool TF32>
static int mc_clusters_resident() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * (device_sms() / 4));
    cfg.blockDim = dim3(P2_THREADS);
    cfg.dynamicSmemBytes = P2_SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int got = 0;
    if (cudaOccupancyMaxActiveClusters(&got, gemm_tc2_kernel<TF32, 4>, &cfg) != cudaSuccess || got <= 0) {
      cudaGetLastError();
      got = -1;
    }
    n = got;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n > 0 ? n : 0;
}

template <bool TF32>
static int gemm_tc2_launch(const AviGemmArgs* a, void* stream) {
  const char* why = tc2_check(a, TF32);
  AVI_REQUIRE(why == nullptr, "%s: %s", TF32 ? "avi_gemm_tf32_tc" : "avi_gemm_bf16_tc", why);
  constexpr int BK = TF32 ? P2_BK / 2 : P2_BK;
  constexpr uint64_t ES = TF32 ? 4 : 2;
  constexpr CUtensorMapDataType DT = TF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int cin = a->K / a->conv_taps;
  const int s = a->conv_stride;
  // two pairs per cluster with the W tile multicast, whenever the pairs would walk more than one wave of tiles
  const bool no_mc = g_gemm_multicast.load(std::memory_order_relaxed) == 0;
  const int max_pairs = device_sms() / 2;
  const int m_tiles_all = ((a->rows + 2 * P2_BM - 1) / (2 * P2_BM)) * a->batch, n_tiles_all = (a->N + P2_BN - 1) / P2_BN;
  static SmemOptIn optin, optin_mc;   // one per template instance; per-device flags inside
  int mc_clusters = 0;
  if (!no_mc && (int64_t)m_tiles_all * n_tiles_all > max_pairs) {
    const cudaError_t e = smem_optin(gemm_tc2_kernel<TF32, 4>, (int)P2_SMEM_BYTES, optin_mc);
    AVI_REQUIRE(e == cudaSuccess, "avi_gemm_bf16_tc: cannot opt in to %u bytes of shared memory: %s", P2_SMEM_BYTES, cudaGetErrorString(e));
    mc_clusters = mc_clusters_resident<TF32>();
    if (2 * mc_clusters < max_pairs - 4) mc_clusters = 0;   // a device that strands more than 8 SMs keeps the pair kernel
  }
  const bool mc = mc_clusters > 0;
  CUtensorMap map_a, map_w;
  {
    // (channel, phase, super-row, clip)
    const uint64_t q_rows = (uint64_t)(a->a_rows_alloc / s);
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)s, q_rows, (uint64_t)a->batch};
    uint64_t strides[3] = {(uint64_t)a->a_ld * ES, (uint64_t)a->a_ld * s * ES, (uint64_t)a->a_batch_stride * ES};
    if (a->batch == 1) strides[2] = dims[2] * strides[1];  // unused; keep it well-formed
    uint32_t box[4] = {BK, 1, P2_BM, 1};
    if (encode_map(&map_a, a->A, 4, dims, strides, box, DT)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t strides[1] = {(uint64_t)a->K * ES};
    uint32_t box[2] = {BK, mc ? (uint32_t)P2_BNH / 2 : (uint32_t)P2_BNH};
    if (encode_map(&map_w, a->W, 2, dims, strides, box, DT)) return 1;
  }
  Tc2Params p;
  p.bias = a->bias;
  p.residual = a->residual;
  p.C = a->C;
  p.C2 = a->C2;
  p.c_dtype = a->c_dtype;
  p.act = a->act;
  p.batch = a->batch;
  p.rows = a->rows;
  p.N = a->N;
  p.num_kb = a->K / BK;
  p.kb_per_tap = cin / BK;
  p.conv_stride = s;
  p.taps_x = a->conv_taps_x;
  p.row_skip = a->conv_taps_x != 0 ? a->conv_row_pitch - a->conv_taps_x : 0;
  p.m_tiles = (a->rows + 2 * P2_BM - 1) / (2 * P2_BM);
  p.n_tiles = (a->N + P2_BN - 1) / P2_BN;
  p.total_tiles = mc ? ((m_tiles_all + 1) / 2) * p.n_tiles : p.m_tiles * p.n_tiles * a->batch;   // mc: super-tiles of two m-tiles
  p.c_ld = a->c_ld;
  p.c_batch_stride = a->c_batch_stride;
  p.res_ld = a->res_ld;
  p.res_batch_stride = a->res_batch_stride;
  // vector path: every 32-column chunk of every row is 16-byte addressable in both output dtypes and the residual
  bool vec = (a->c_ld % 8 == 0) && (a->c_batch_stride % 8 == 0) && ((uintptr_t)a->C % 16 == 0) &&
             (a->C2 == nullptr || (uintptr_t)a->C2 % 16 == 0);
  if (a->residual) vec = vec && (a->res_ld % 4 == 0) && (a->res_batch_stride % 4 == 0) && ((uintptr_t)a->residual % 16 == 0);
  p.vec_ok = vec ? 1 : 0;
  // TMA-store epilogue: single output, 16-byte addressable rows; a residual is supported when it IS the output (in-place update
  // of the fp32 residual stream through TMA reduce-add)
  static const bool no_tma_store = getenv("AVI_GEMM_NO_TMA_STORE") != nullptr;
  const bool inplace_res = a->residual != nullptr && (const void*)a->residual == (const void*)a->C && a->c_dtype == AVI_DT_F32 &&
                           a->res_ld == a->c_ld && a->res_batch_stride == a->c_batch_stride;
  p.tma_store = 0;
  p.dynamic = 0;
  CUtensorMap map_c = map_a;
  // The TMA unit clips stores at the tensor bounds in 16-byte granules of the innermost dimension (measured:
  // profiles/probes/tma_clip_probe.py): with N * elemsize not a multiple of 16 the granule holding column N-1 is written whole
  // (zeros in the surplus columns). That is only acceptable when those columns are the row's own padding, i.e. the pitch ends
  // exactly at that granule (the 15069-wide vertex rows on their 15072 pitch); a narrower window inside a wider buffer keeps the
  // per-thread store path.
  const uint64_t es_c = a->c_dtype == AVI_DT_F32 ? 4 : 2;
  const uint64_t row_bytes = (uint64_t)a->N * es_c, row_bytes16 = (row_bytes + 15) / 16 * 16;
  const bool clip_ok = row_bytes == row_bytes16 || row_bytes16 == (uint64_t)a->c_ld * es_c;
  if (!no_tma_store && vec && clip_ok && a->C2 == nullptr && (a->residual == nullptr || inplace_res)) {
    const uint64_t es = es_c;
    uint64_t dims[3] = {(uint64_t)a->N, (uint64_t)a->rows, (uint64_t)a->batch};
    uint64_t strides[2] = {(uint64_t)a->c_ld * es, (uint64_t)a->c_batch_stride * es};
    if (a->batch == 1) strides[1] = dims[1] * strides[0];
    uint32_t box[3] = {a->c_dtype == AVI_DT_F32 ? 16u : 32u, 32u, 1u};
    if (encode_map(&map_c, a->C, 3, dims, strides, box,
                   a->c_dtype == AVI_DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   CU_TENSOR_MAP_SWIZZLE_64B))
      return 1;
    p.tma_store = inplace_res ? 2 : 1;
    if (inplace_res) p.residual = nullptr;   // the memory system performs the addition
  }

  if (mc) {
    const int clusters = p.total_tiles < mc_clusters ? p.total_tiles : mc_clusters;
    const cudaError_t le = launch_pdl(gemm_tc2_kernel<TF32, 4>, dim3(4 * clusters), dim3(P2_THREADS), P2_SMEM_BYTES, (cudaStream_t)stream,
                                      map_a, map_w, map_c, p);
    AVI_REQUIRE(le == cudaSuccess, "avi_gemm_bf16_tc: launch failed: %s", cudaGetErrorString(le));
    return check_launch(TF32 ? "gemm_tf32_tc" : "gemm_bf16_tc");
  }
  const cudaError_t attr_err = smem_optin(gemm_tc2_kernel<TF32, 2>, (int)P2_SMEM_BYTES, optin);
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_gemm_bf16_tc: cannot opt in to %u bytes of shared memory: %s", P2_SMEM_BYTES,
              cudaGetErrorString(attr_err));
  // dynamic: one cluster per tile, of which at most max_pairs are ever resident; the rest are cancelled by the resident ones
  p.dynamic = ((dynamic_tiles_mask() & AVI_DYN_GEMM) != 0 && p.total_tiles > 1) ? 1 : 0;
  const int pairs = p.dynamic ? p.total_tiles : (p.total_tiles < max_pairs ? p.total_tiles : max_pairs);
  const cudaError_t le = launch_pdl(gemm_tc2_kernel<TF32, 2>, dim3(2 * pairs), dim3(P2_THREADS), P2_SMEM_BYTES, (cudaStream_t)stream, map_a,
                                    map_w, map_c, p);
  AVI_REQUIRE(le == cudaSuccess, "avi_gemm_bf16_tc: launch failed: %s", cudaGetErrorString(le));
  return check_launch(TF32 ? "gemm_tf32_tc" : "gemm_bf16_tc");
}

extern "C" int avi_gemm_bf16_tc(const AviGemmArgs* a, void* stream) {
  return gemm_tc2_launch<false>(a, stream);
}

extern "C" int avi_gemm_tf32_tc(const AviGemmArgs* a, void* stream) { return gemm_tc2_launch<true>(a, stream); }
