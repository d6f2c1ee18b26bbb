// Probe: can a tcgen05.mma shared-memory descriptor start at an arbitrary ROW of a SWIZZLE_128B K-major slab?
// (Needed by the implicit positional conv: the (128 + 127)-row activation slab stays resident in shared memory and tap j reads
//  rows [j, j + 128) of it, instead of re-fetching a shifted A tile from L2 for every tap.)
//
// One CTA: TMA loads a [256 rows x 64 bf16] slab (two 128-row boxes, 32 KB, 1024-byte aligned) and a [64 x 64] identity W tile;
// for each row offset j the MMA D[128 x 64] = A_j * I^T is issued with the A descriptor's start address advanced by j * 128 bytes,
// once with base_offset = 0 and once with base_offset = (j & 7) (descriptor bits [49, 52)); D is compared with slab rows j.. .
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I avi_talking_b200/csrc -o build/probes/desc_rowoffset_probe \
//        profiles/probes/desc_rowoffset_probe.cu -lcuda && build/probes/desc_rowoffset_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

namespace avi {
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}
std::atomic<int64_t> g_launches{0};
}  // namespace avi
using namespace avi;

constexpr int ROWS = 256, KC = 64, NCOL = 64;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                                                       float* __restrict__ out, int row_off, int base_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;                      // 256 rows x 128 B
  uint8_t* sw = smem + ROWS * 128;         // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ROWS * 128 + NCOL * 128);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), ROWS * 128 + NCOL * 128);
    tma_load_2d(smem_u32(sa), &map_a, smem_u32(&bars[0]), 0, 0);
    tma_load_2d(smem_u32(sa + 128 * 128), &map_a, smem_u32(&bars[0]), 0, 128);
    tma_load_2d(smem_u32(sw), &map_w, smem_u32(&bars[0]), 0, 0);
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t adesc = umma_desc_sw128(smem_u32(sa) + (uint32_t)row_off * 128u) | ((uint64_t)(base_off & 7) << 49);
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sw));
#pragma unroll
    for (int k = 0; k < KC / 16; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0 ? 1u : 0u);
    umma_commit(smem_u32(&bars[1]));
  }
  __syncwarp();
  if (lane == 0) mbar_wait(smem_u32(&bars[1]), 0);
  __syncwarp();
  tc_fence_after();
  for (int c0 = 0; c0 < NCOL; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * NCOL + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
  }
}

int main() {
  std::vector<__nv_bfloat16> ha(ROWS * KC), hw(NCOL * KC);
  for (int r = 0; r < ROWS; ++r)
    for (int k = 0; k < KC; ++k) ha[r * KC + k] = __float2bfloat16((float)((r * 7 + k * 3) % 251) - 125.f);   // exact in bf16
  for (int n = 0; n < NCOL; ++n)
    for (int k = 0; k < KC; ++k) hw[n * KC + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *da, *dw;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&dw, hw.size() * 2);
  cudaMalloc(&dout, 128 * NCOL * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap map_a, map_w;
  {
    uint64_t dims[2] = {KC, ROWS}, strides[1] = {KC * 2};
    uint32_t box[2] = {KC, 128};
    if (encode_map(&map_a, da, 2, dims, strides, box)) return 1;
  }
  {
    uint64_t dims[2] = {KC, NCOL}, strides[1] = {KC * 2};
    uint32_t box[2] = {KC, NCOL};
    if (encode_map(&map_w, dw, 2, dims, strides, box)) return 1;
  }
  const int smem_bytes = ROWS * 128 + NCOL * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<float> hout(128 * NCOL);
  const int offs[] = {0, 1, 2, 3, 5, 7, 8, 9, 17, 64, 100, 127};
  for (int off : offs)
    for (int mode = 0; mode < 2; ++mode) {
      const int bo = mode == 0 ? 0 : (off & 7);
      if (mode == 1 && bo == 0) continue;
      cudaMemset(dout, 0, 128 * NCOL * 4);
      probe_kernel<<<1, 128, smem_bytes>>>(map_a, map_w, dout, off, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("row_off %3d base_offset %d: CUDA error %s\n", off, bo, cudaGetErrorString(e));
        return 2;
      }
      cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first_bad_row = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NCOL; ++n) {
          const float want = __bfloat162float(ha[(off + m) * KC + n]);
          if (hout[m * NCOL + n] != want) {
            if (!bad) first_bad_row = m;
            ++bad;
          }
        }
      printf("row_off %3d base_offset %d: %s (%d mismatches, first bad row %d)\n", off, bo, bad ? "MISMATCH" : "exact", bad, first_bad_row);
    }
  return 0;
}
