// Probe of cluster launch control on B200 (profiles/probes, not product code): a kernel launched with one 2-CTA cluster per tile whose
// resident clusters steal the pending ones (try_cancel, multicast to both CTAs). Checks that every tile is processed exactly once and
// times the kernel alone and beside a 64-CTA blocker kernel that holds 64 SMs for ~1 ms (the autoregressive decoder's footprint).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I avi_talking_b200/csrc -o /tmp/clc_probe profiles/probes/clc_probe.cu
#include <cstdio>
#include <vector>

#include "tc_common.cuh"
namespace avi { void set_error(const char*, ...) {} std::atomic<int64_t> g_launches{0}; }
using namespace avi;

constexpr int NS = 4;

template <bool DYNAMIC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
tile_kernel(int* hits, int* launched, int n_tiles, long long spin, int smem_dummy) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [NS]
  uint64_t* empty = full + NS;                                  // [NS] (rank 0)
  uint8_t* resp = smem + 128;                                   // [NS][16]
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 2 * 4);                    // 4 warps per CTA, 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (rank == 0) atomicAdd(launched, 1);
  }
  cluster_sync_all();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int slot = 0;
  uint32_t phase = 0;
  int islot = 0;
  uint32_t iphase = 0;
  auto issue = [&]() {   // rank 0, thread 0
    mbar_wait(smem_u32(&empty[islot]), iphase ^ 1);
    mbar_expect_tx(smem_u32(&full[islot]), 16);
    mbar_expect_tx_cluster(mapa_shared(smem_u32(&full[islot]), 1), 16);
    clc_try_cancel_multicast(smem_u32(resp + 16 * islot), smem_u32(&full[islot]));
    if (++islot == NS) { islot = 0; iphase ^= 1; }
  };
  int t = blockIdx.x / 2;
  if (!DYNAMIC) {
    for (; t < n_tiles; t += gridDim.x / 2) {
      if (rank == 0 && threadIdx.x == 0) atomicAdd(&hits[t], 1);
      const long long t0 = clock64();
      while (clock64() - t0 < spin) {}
    }
    return;
  }
  if (rank == 0 && threadIdx.x == 0) issue();
  while (true) {
    if (rank == 0 && threadIdx.x == 0) atomicAdd(&hits[t], 1);
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    int nxt = 0;
    if (lane == 0) {
      mbar_wait(smem_u32(&full[slot]), phase);
      nxt = clc_decode(smem_u32(resp + 16 * slot));
      fence_proxy_async_smem();
      mbar_arrive_cluster(mapa_shared(smem_u32(&empty[slot]), 0));
    }
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    if (++slot == NS) { slot = 0; phase ^= 1; }
    if (nxt < 0) break;
    t = nxt / 2;
    if (rank == 0 && threadIdx.x == 0) issue();
  }
  cluster_sync_all();   // the peer's barriers stay valid until both CTAs are done
}

__global__ void __launch_bounds__(256, 1) blocker(long long spin) {
  extern __shared__ uint8_t big[];
  big[threadIdx.x] = 1;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}
}

int main() {
  const int n_tiles = 2000;
  int *hits, *launched;
  cudaMalloc(&hits, n_tiles * 4);
  cudaMalloc(&launched, 4);
  const int smem = 200 * 1024;   // one CTA per SM, like the GEMM
  cudaFuncSetAttribute(tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(blocker, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaStream_t s1, s2;
  cudaStreamCreate(&s1);
  cudaStreamCreate(&s2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const long long spin = 10000;          // ~5 us per tile
  const long long block_spin = 2000000;  // ~1 ms
  for (int with_blocker = 0; with_blocker < 2; ++with_blocker) {
    for (int dyn = 0; dyn < 2; ++dyn) {
      for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(hits, 0, n_tiles * 4);
        cudaMemset(launched, 0, 4);
        cudaDeviceSynchronize();
        if (with_blocker) blocker<<<64, 256, smem, s2>>>(block_spin);
        cudaEventRecord(e0, s1);
        if (dyn) tile_kernel<true><<<2 * n_tiles, 128, smem, s1>>>(hits, launched, n_tiles, spin, 0);
        else tile_kernel<false><<<148, 128, smem, s1>>>(hits, launched, n_tiles, spin, 0);
        cudaEventRecord(e1, s1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<int> h(n_tiles);
        int l = 0;
        cudaMemcpy(h.data(), hits, n_tiles * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&l, launched, 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < n_tiles; ++i) bad += h[i] != 1;
        printf("blocker=%d dynamic=%d rep=%d: %.3f ms, clusters launched %d, tiles not hit exactly once %d, err=%s\n", with_blocker, dyn, rep, ms, l,
               bad, cudaGetErrorString(err));
      }
    }
  }
  return 0;
}
