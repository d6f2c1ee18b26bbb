// FLAME blendshapes + pose correctives on tcgen05 with linear blend skinning fused as the epilogue (sm_100a).
//
//   v_posed[f, v, c] = v_template[v, c] + sum_l coef[f, l] * dirs[l, v, c]          (l = shape | expression | pose feature)
//   verts[f, v, :]   = sum_j W[v, j] * A[f, j] * [v_posed[f, v, :]; 1]              (lbs.py:188-232)
//
// Mapping: vertices are the MMA M dimension (TMEM lane = vertex), frames the N dimension, one accumulator per coordinate
// (x, y, z), so that after tcgen05.ld a thread holds x/y/z of ITS vertex for 32 frames and applies the per-frame skinning
// transform with its own 5 skinning weights in registers; consecutive lanes are consecutive vertices, so each frame's
// 32 x 3 floats form one contiguous 384-byte run (three 12-byte-strided stores per warp).
// Operands are fp16 (11-bit significand: ~1e-5 m worst case on ~1e-2 m displacements, vs ~1.7e-4 m for bf16); the
// template is added in fp32 in the epilogue so its 8 cm magnitude never passes through 16-bit rounding.
// A CTA keeps the direction slab of its 128-vertex tile (3 x [128 x 192] fp16 = 144 KB) resident in shared memory and
// streams 64-frame tiles of coefficients (24 KB) + joint transforms (15 KB) through a 2-stage ring; two TMEM accumulator
// stages overlap the MMAs of tile i+1 with the skinning/stores of tile i.  The kernel is HBM-write-bound by design
// (60 276 B written per frame vs 6.7 MFLOP per frame of 16-bit tensor work).
#include <cuda_fp16.h>

#include <algorithm>

#include "tc_common.cuh"

namespace avi {

constexpr int FT_BM = 128, FT_NF = 64, FT_K = 192, FT_KC = 64, FT_NCH = FT_K / FT_KC, FT_CHUNK_TILES = 16, FT_NJ = 5;
constexpr uint32_t FT_DIR_TILE = FT_BM * FT_KC * 2;              // 16 KB per (coord, k-chunk)
constexpr uint32_t FT_DIRS_BYTES = 3 * FT_NCH * FT_DIR_TILE;     // 144 KB
constexpr uint32_t FT_COEF_TILE = FT_NF * FT_KC * 2;             // 8 KB per k-chunk
constexpr uint32_t FT_COEF_BYTES = FT_NCH * FT_COEF_TILE;        // 24 KB per stage
constexpr uint32_t FT_AS_BYTES = FT_NF * FT_NJ * 12 * 4;         // 15 KB per stage
constexpr int FT_EPI_WARPS = 16, FT_THREADS = (2 + FT_EPI_WARPS) * 32;  // 4 epilogue warps per SM sub-partition
constexpr uint32_t FT_TR_BYTES = 0;
constexpr uint32_t FT_OFF_COEF = FT_DIRS_BYTES;
constexpr uint32_t FT_OFF_AS = FT_OFF_COEF + 2 * FT_COEF_BYTES;
constexpr uint32_t FT_OFF_TR = FT_OFF_AS + 2 * FT_AS_BYTES;
constexpr uint32_t FT_OFF_FLAG = FT_OFF_TR + FT_TR_BYTES;            // per-frame bitmask of joints whose transform is not the identity
constexpr uint32_t FT_OFF_BAR = FT_OFF_FLAG + 2 * FT_NF * 4;
constexpr uint32_t FT_SMEM = FT_OFF_BAR + 128 + 1024;
static_assert(FT_SMEM <= 232448, "shared memory budget");

struct FlameTcParams {
  const float* A;          // [F][5][12] relative joint transforms (3x4 each)
  const float* lbs_w;      // [V][5]
  const float* v_template; // [V][3]
  float* verts;            // [F][V][3]
  int F, V, V_pad;
  int n_vt, n_ftiles, n_items;
  // frame groups (clips): tiles never cross a group, so a per-group template (shape blendshapes hoisted out of the per-frame
  // contraction) is uniform within a tile. One group of F frames = the plain case.
  int frames_per_group, tiles_per_group;
  int64_t template_stride;  // floats between consecutive group templates (0: one global template)
  int64_t verts_stride;     // floats between consecutive output frames (>= V*3; a multiple of 4 keeps every frame 16-byte aligned)
};

// first frame / number of valid frames of frame tile ft
__device__ __forceinline__ void flame_tile_frames(const FlameTcParams& p, int ft, int& f0, int& nfr, int& group) {
  group = ft / p.tiles_per_group;
  const int t0 = (ft % p.tiles_per_group) * FT_NF;
  f0 = group * p.frames_per_group + t0;
  nfr = min(FT_NF, p.frames_per_group - t0);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);  // same instruction (kind::f16); idesc selects fp16 operands
}

__global__ void __launch_bounds__(FT_THREADS, 1)
flame_tc_kernel(const __grid_constant__ CUtensorMap map_dirs, const __grid_constant__ CUtensorMap map_coef, const FlameTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FT_OFF_BAR);
  uint64_t* dirs_full = bars;          // [1]
  uint64_t* dirs_empty = bars + 1;     // [1]
  uint64_t* cf_full = bars + 2;        // [2] coefficients + joint transforms landed
  uint64_t* coef_empty = bars + 4;     // [2] MMAs done reading the coefficient tiles
  uint64_t* as_empty = bars + 6;       // [2] epilogue done reading the joint transforms
  uint64_t* tmem_full = bars + 8;      // [2]
  uint64_t* tmem_empty = bars + 10;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dirs) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_coef) : "memory");
    mbar_init(smem_u32(dirs_full), 1);
    mbar_init(smem_u32(dirs_empty), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&cf_full[s]), 1);
      mbar_init(smem_u32(&coef_empty[s]), 1);
      mbar_init(smem_u32(&as_empty[s]), FT_EPI_WARPS);
      mbar_init(smem_u32(&tmem_full[s]), 1);
      mbar_init(smem_u32(&tmem_empty[s]), FT_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ============ TMA producer ============
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int item_it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++item_it) {
        const int vt = item % p.n_vt, chunk = item / p.n_vt;
        mbar_wait(smem_u32(dirs_empty), (item_it & 1) ^ 1);
        const uint32_t db = smem_u32(dirs_full);
        mbar_expect_tx(db, FT_DIRS_BYTES);
        for (int k = 0; k < 3; ++k)
          for (int c = 0; c < FT_NCH; ++c)
            tma_load_2d(smem_u32(smem + (k * FT_NCH + c) * FT_DIR_TILE), &map_dirs, db, c * FT_KC, k * p.V_pad + vt * FT_BM);
        const int ft0 = chunk * FT_CHUNK_TILES;
        const int ft1 = min(ft0 + FT_CHUNK_TILES, p.n_ftiles);
        for (int ft = ft0; ft < ft1; ++ft) {
          mbar_wait(smem_u32(&coef_empty[stage]), phase ^ 1);
          mbar_wait(smem_u32(&as_empty[stage]), phase ^ 1);
          int f0, nfr, group;
          flame_tile_frames(p, ft, f0, nfr, group);
          const uint32_t fb = smem_u32(&cf_full[stage]);
          mbar_expect_tx(fb, FT_COEF_BYTES + (uint32_t)nfr * FT_NJ * 12 * 4);
          for (int c = 0; c < FT_NCH; ++c)
            tma_load_2d(smem_u32(smem + FT_OFF_COEF + stage * FT_COEF_BYTES + c * FT_COEF_TILE), &map_coef, fb, c * FT_KC, f0);
          bulk_load_1d(smem_u32(smem + FT_OFF_AS + stage * FT_AS_BYTES), p.A + (int64_t)f0 * FT_NJ * 12,
                       (uint32_t)nfr * FT_NJ * 12 * 4, fb);
          stage ^= 1;
          if (stage == 0) phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ============ MMA issuer ============
    if (lane == 0) {
      // D = f32, A = B = fp16 (format 0), both K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(FT_NF >> 3) << 17) | ((uint32_t)(FT_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int item_it = 0, acc_it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++item_it) {
        const int chunk = item / p.n_vt;
        mbar_wait(smem_u32(dirs_full), item_it & 1);
        tc_fence_after();
        const int ft0 = chunk * FT_CHUNK_TILES;
        const int ft1 = min(ft0 + FT_CHUNK_TILES, p.n_ftiles);
        for (int ft = ft0; ft < ft1; ++ft, ++acc_it) {
          const int as = acc_it & 1;
          const uint32_t aph = (acc_it >> 1) & 1;
          mbar_wait(smem_u32(&tmem_empty[as]), aph ^ 1);
          mbar_wait(smem_u32(&cf_full[stage]), phase);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int c = 0; c < FT_NCH; ++c) {
              const uint64_t ad = umma_desc_sw128(smem_u32(smem + (k * FT_NCH + c) * FT_DIR_TILE));
              const uint64_t bd = umma_desc_sw128(smem_u32(smem + FT_OFF_COEF + stage * FT_COEF_BYTES + c * FT_COEF_TILE));
#pragma unroll
              for (int kk = 0; kk < FT_KC / 16; ++kk)
                umma_f16(tmem_base + as * (3 * FT_NF) + k * FT_NF, ad + 2 * kk, bd + 2 * kk, idesc, (c | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&coef_empty[stage]));
          umma_commit(smem_u32(&tmem_full[as]));
          stage ^= 1;
          if (stage == 0) phase ^= 1;
        }
        umma_commit(smem_u32(dirs_empty));  // the direction slab may be replaced once every MMA of this item has retired
      }
    }
  } else {
    // ============ epilogue: template add + skinning + coalesced stores ============
    // 16 warps: TMEM lane quarter = warp % 4 (32 vertices), frame quarter = (warp - 2) / 4 (16 of the tile's 64 frames)
    const int ew = warp - 2;
    const int quarter = warp & 3, fq = ew >> 2;
    const int64_t V3 = p.verts_stride;
    int stage = 0;
    uint32_t phase = 0;
    int acc_it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int vt = item % p.n_vt, chunk = item / p.n_vt;
      const int vw0 = vt * FT_BM + quarter * 32;  // first vertex of this warp
      const int v = vw0 + lane;
      float w[FT_NJ], vtp[3];
      const bool v_ok = v < p.V;
#pragma unroll
      for (int j = 0; j < FT_NJ; ++j) w[j] = (v < p.V) ? p.lbs_w[(int64_t)v * FT_NJ + j] : 0.f;
      const float wsum = (w[0] + w[1]) + (w[2] + w[3]) + w[4];
#pragma unroll
      for (int c = 0; c < 3; ++c) vtp[c] = (v < p.V) ? p.v_template[(int64_t)v * 3 + c] : 0.f;
      const int ft0 = chunk * FT_CHUNK_TILES;
      const int ft1 = min(ft0 + FT_CHUNK_TILES, p.n_ftiles);
      for (int ft = ft0; ft < ft1; ++ft, ++acc_it) {
        const int as = acc_it & 1;
        const uint32_t aph = (acc_it >> 1) & 1;
        int tf0, tnfr, group;
        flame_tile_frames(p, ft, tf0, tnfr, group);
        if (p.template_stride != 0 && v_ok) {  // per-group (per-clip) shaped template
          const float* tp = p.v_template + (int64_t)group * p.template_stride + (int64_t)v * 3;
          vtp[0] = __ldg(tp);
          vtp[1] = __ldg(tp + 1);
          vtp[2] = __ldg(tp + 2);
        }
        mbar_wait(smem_u32(&cf_full[stage]), phase);   // joint transforms of this tile are in smem
        // Joints whose relative transform is the identity contribute w_j * I: T = (sum_j w_j) I + sum_{active j} w_j (A_j - I).
        // On this path only the jaw moves (global / neck / eye poses are zero: faceformer_disentangle.py:425-433, Preprocessors.py:80),
        // so 4 of the 5 joints drop out: 12 FMAs and 3 shared loads per vertex-frame instead of 60 and 15.
        uint32_t* flags = reinterpret_cast<uint32_t*>(smem + FT_OFF_FLAG) + stage * FT_NF;
        {
          const int etid = threadIdx.x - 64;
          if (etid < FT_NF) {
            uint32_t m = 0;
            if (etid < tnfr) {
              const uint32_t Af = smem_u32(smem + FT_OFF_AS + stage * FT_AS_BYTES) + etid * (FT_NJ * 12 * 4);
#pragma unroll
              for (int j = 0; j < FT_NJ; ++j) {
                float d = 0.f;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                  float4 a = lds128f(Af + (j * 12 + q * 4) * 4);
                  // stored back as D = A - I: the epilogue below works on the deviation from the identity
                  a.x -= (q == 0 ? 1.f : 0.f);
                  a.y -= (q == 1 ? 1.f : 0.f);
                  a.z -= (q == 2 ? 1.f : 0.f);
                  sts128(Af + (j * 12 + q * 4) * 4, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
                  d = fmaxf(d, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
                }
                if (d > 1e-7f) m |= 1u << j;
              }
            }
            flags[etid] = m;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(FT_EPI_WARPS * 32) : "memory");
        }
        mbar_wait(smem_u32(&tmem_full[as]), aph);
        tc_fence_after();
        uint32_t vx[16], vy[16], vz[16];
        const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * (3 * FT_NF) + fq * 16);
        tmem_ld16(ta, vx);
        tmem_ld16(ta + FT_NF, vy);
        tmem_ld16(ta + 2 * FT_NF, vz);
        const uint32_t As = smem_u32(smem + FT_OFF_AS + stage * FT_AS_BYTES) + (fq * 16) * (FT_NJ * 12 * 4);
        const int fbase = tf0 + fq * 16;
        const int flimit = tf0 + tnfr;
        float* o = p.verts + (int64_t)fbase * V3 + (int64_t)v * 3;   // lanes = consecutive vertices: 384 contiguous bytes per frame
        // union of the non-identity joints over this warp's 16 frames: warp-uniform, hoisted out of the frame loop. A joint that is
        // the identity in SOME of those frames contributes w_j * (A_j - I) = exact zeros there, so the result is bit-identical to
        // testing every frame, without the ~30 branch / predicate instructions per frame that dominated the issue slots (ncu r1i).
        const uint32_t wmask = __reduce_or_sync(0xffffffffu, lane < 16 ? flags[fq * 16 + lane] : 0u);
        if ((wmask & (wmask - 1)) == 0) {
          // ---- at most ONE moving joint (the jaw, on every path of this repo): T = wsum I + w_j (A_j - I) ----
          const int j0 = wmask ? (31 - __clz((int)wmask)) : 0;
          const float wj = wmask ? w[0] * (j0 == 0) + w[1] * (j0 == 1) + w[2] * (j0 == 2) + w[3] * (j0 == 3) + w[4] * (j0 == 4) : 0.f;
          const uint32_t Aj = As + (j0 * 12) * 4;
#pragma unroll
          for (int n = 0; n < 16; ++n) {
            if (fbase + n < flimit) {  // warp-uniform
              // verts = sum_j w_j A_j [p; 1] = wsum p + w_j (D [p; 1]),  D = A_j - I (prepared in shared memory above)
              const float4 r0 = lds128f(Aj + (n * (FT_NJ * 12)) * 4);
              const float4 r1 = lds128f(Aj + (n * (FT_NJ * 12) + 4) * 4);
              const float4 r2 = lds128f(Aj + (n * (FT_NJ * 12) + 8) * 4);
              const float px = __uint_as_float(vx[n]) + vtp[0], py = __uint_as_float(vy[n]) + vtp[1], pz = __uint_as_float(vz[n]) + vtp[2];
              if (v_ok) {
                o[0] = fmaf(wj, fmaf(r0.x, px, fmaf(r0.y, py, fmaf(r0.z, pz, r0.w))), wsum * px);
                o[1] = fmaf(wj, fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w))), wsum * py);
                o[2] = fmaf(wj, fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w))), wsum * pz);
              }
            }
            o += V3;
          }
        } else {
#pragma unroll
          for (int n = 0; n < 16; ++n) {
            if (fbase + n < flimit) {  // warp-uniform
              float T[12];
#pragma unroll
              for (int e = 0; e < 12; ++e) T[e] = (e == 0 || e == 5 || e == 10) ? wsum : 0.f;
#pragma unroll
              for (int j = 0; j < FT_NJ; ++j) {
                if (wmask & (1u << j)) {
#pragma unroll
                  for (int q = 0; q < 3; ++q) {
                    const float4 a = lds128f(As + (n * (FT_NJ * 12) + j * 12 + q * 4) * 4);
                    T[q * 4 + 0] = fmaf(w[j], a.x, T[q * 4 + 0]);
                    T[q * 4 + 1] = fmaf(w[j], a.y, T[q * 4 + 1]);
                    T[q * 4 + 2] = fmaf(w[j], a.z, T[q * 4 + 2]);
                    T[q * 4 + 3] = fmaf(w[j], a.w, T[q * 4 + 3]);
                  }
                }
              }
              const float px = __uint_as_float(vx[n]) + vtp[0], py = __uint_as_float(vy[n]) + vtp[1], pz = __uint_as_float(vz[n]) + vtp[2];
              if (v_ok) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
                  o[i] = fmaf(T[i * 4 + 0], px, fmaf(T[i * 4 + 1], py, fmaf(T[i * 4 + 2], pz, T[i * 4 + 3])));
              }
            }
            o += V3;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&tmem_empty[as]));
          mbar_arrive(smem_u32(&as_empty[stage]));
        }
        stage ^= 1;
        if (stage == 0) phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// dirs16[c][v][l] = dirs32[l][v*3 + c] for l < n_dirs (shape | expression | pose rows; the template row is excluded), zero padded
__global__ void flame_pack_tc_kernel(const float* __restrict__ dirs32, __half* __restrict__ dirs16, int V, int V_pad, int row0, int n_dirs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)3 * V_pad * FT_K;
  if (i >= total) return;
  const int l = (int)(i % FT_K);
  const int v = (int)((i / FT_K) % V_pad);
  const int c = (int)(i / ((int64_t)FT_K * V_pad));
  float x = 0.f;
  if (v < V && l < n_dirs) x = dirs32[(int64_t)(row0 + l) * V * 3 + v * 3 + c];
  dirs16[i] = __float2half_rn(x);
}

// coef16[f][l] = fp16(coef32[f][l]) for l < n_dirs, zero padded to FT_K
__global__ void flame_coef16_kernel(const float* __restrict__ coef32, __half* __restrict__ coef16, int F, int K_pad32, int col0, int n_dirs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)F * FT_K) return;
  const int l = (int)(i % FT_K);
  const int64_t f = i / FT_K;
  coef16[i] = __float2half_rn(l < n_dirs ? coef32[f * K_pad32 + col0 + l] : 0.f);
}

}  // namespace avi

using namespace avi;

static std::atomic<int> g_flame_max_ctas{kNumSMs};

// Process-wide cap on the CTAs of the tensor-core FLAME kernel (default: one per SM). A caller that runs FLAME concurrently with
// a kernel occupying part of the GPU (the 64-CTA autoregressive decoder) sizes it to the SMs left over.
extern "C" int avi_flame_set_max_ctas(int32_t n) {
  AVI_REQUIRE(n >= 1 && n <= kNumSMs, "avi_flame_set_max_ctas: n must be in 1..%d", kNumSMs);
  g_flame_max_ctas.store(n);
  return 0;
}

extern "C" int avi_flame_tc_supported(int32_t NB) { return (NB + 36 <= FT_K) ? 1 : 0; }

extern "C" int avi_flame_pack_tc_rows(const float* dirs32, void* dirs16, int32_t V, int32_t row0, int32_t n_dirs, int32_t V_pad,
                                      void* stream) {
  AVI_REQUIRE(V > 0 && row0 >= 0 && n_dirs > 0 && n_dirs <= FT_K && V_pad % FT_BM == 0 && V_pad >= V,
              "avi_flame_pack_tc_rows: unsupported shape (n_dirs=%d V_pad=%d)", n_dirs, V_pad);
  const int64_t total = (int64_t)3 * V_pad * FT_K;
  flame_pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dirs32, (__half*)dirs16, V, V_pad, row0, n_dirs);
  return check_launch("flame_pack_tc");
}

extern "C" int avi_flame_pack_tc(const float* dirs32, void* dirs16, int32_t V, int32_t NB, int32_t V_pad, void* stream) {
  AVI_REQUIRE(NB + 36 <= FT_K, "avi_flame_pack_tc: NB + 36 must be <= %d", FT_K);
  return avi_flame_pack_tc_rows(dirs32, dirs16, V, 0, NB + 36, V_pad, stream);
}

// Grouped form: frames are G groups (clips) of frames_per_group; `templates` holds one [V,3] template per group (template_stride
// floats apart; 0 = one global template); the tensor-core contraction runs over coefficient columns [coef_col0, coef_col0 + n_dirs)
// of coef32 against the n_dirs direction rows packed by avi_flame_pack_tc_rows.
extern "C" int avi_flame_blend_skin_tc_grouped(const float* coef32, const float* A, const void* dirs16, const float* lbs_weights,
                                               const float* templates, int64_t template_stride, void* coef16, float* verts,
                                               int64_t verts_frame_stride, int32_t F, int32_t V, int32_t n_dirs, int32_t coef_col0,
                                               int32_t K_pad32, int32_t V_pad, int32_t frames_per_group, void* stream) {
  AVI_REQUIRE(verts_frame_stride >= (int64_t)V * 3, "avi_flame_blend_skin_tc: verts_frame_stride smaller than V*3");
  AVI_REQUIRE(F > 0 && V > 0 && n_dirs > 0 && n_dirs <= FT_K && V_pad % FT_BM == 0 && V_pad >= V && coef_col0 >= 0 &&
                  coef_col0 + n_dirs <= K_pad32,
              "avi_flame_blend_skin_tc: unsupported shape");
  AVI_REQUIRE(frames_per_group > 0 && F % frames_per_group == 0, "avi_flame_blend_skin_tc: F must be a multiple of frames_per_group");
  AVI_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)dirs16 % 16 == 0) && ((uintptr_t)coef16 % 16 == 0),
              "avi_flame_blend_skin_tc: unaligned pointers");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n16 = (int64_t)F * FT_K;
  flame_coef16_kernel<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>(coef32, (__half*)coef16, F, K_pad32, coef_col0, n_dirs);
  if (check_launch("flame_coef16")) return 1;
  CUtensorMap map_dirs, map_coef;
  {
    uint64_t dims[2] = {(uint64_t)FT_K, (uint64_t)3 * V_pad};
    uint64_t strides[1] = {(uint64_t)FT_K * 2};
    uint32_t box[2] = {FT_KC, FT_BM};
    if (encode_map(&map_dirs, dirs16, 2, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16)) return 1;
  }
  {
    uint64_t dims[2] = {(uint64_t)FT_K, (uint64_t)F};
    uint64_t strides[1] = {(uint64_t)FT_K * 2};
    uint32_t box[2] = {FT_KC, FT_NF};
    if (encode_map(&map_coef, coef16, 2, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16)) return 1;
  }
  FlameTcParams p;
  p.A = A;
  p.lbs_w = lbs_weights;
  p.v_template = templates;
  p.verts = verts;
  p.F = F;
  p.V = V;
  p.V_pad = V_pad;
  p.n_vt = V_pad / FT_BM;
  p.frames_per_group = frames_per_group;
  p.tiles_per_group = (frames_per_group + FT_NF - 1) / FT_NF;
  p.template_stride = template_stride;
  p.verts_stride = verts_frame_stride;
  p.n_ftiles = (F / frames_per_group) * p.tiles_per_group;
  const int n_chunks = (p.n_ftiles + FT_CHUNK_TILES - 1) / FT_CHUNK_TILES;
  p.n_items = p.n_vt * n_chunks;
  static SmemOptIn optin;
  const cudaError_t attr_err = smem_optin(flame_tc_kernel, (int)FT_SMEM, optin);
  AVI_REQUIRE(attr_err == cudaSuccess, "avi_flame_blend_skin_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int cap = std::min(g_flame_max_ctas.load(), device_sms());
  const int grid = p.n_items < cap ? p.n_items : cap;
  flame_tc_kernel<<<grid, FT_THREADS, FT_SMEM, st>>>(map_dirs, map_coef, p);
  return check_launch("flame_tc");
}

// The prologue (coefficient rows, joints, kinematic chain, optional landmark rows) is avi_flame_lbs_fwd's; this entry replaces
// only its blend+skin kernel. coef32 [F, K_pad32] and A [F,5,12] must already have been produced by the prologue on `stream`.
extern "C" int avi_flame_blend_skin_tc(const float* coef32, const float* A, const void* dirs16, const float* lbs_weights,
                                       const float* v_template, void* coef16, float* verts, int32_t F, int32_t V, int32_t NB,
                                       int32_t K_pad32, int32_t V_pad, void* stream) {
  AVI_REQUIRE(NB + 36 <= FT_K, "avi_flame_blend_skin_tc: NB + 36 must be <= %d", FT_K);
  return avi_flame_blend_skin_tc_grouped(coef32, A, dirs16, lbs_weights, v_template, 0, coef16, verts, (int64_t)V * 3, F, V, NB + 36, 0,
                                         K_pad32, V_pad, F, stream);
}
