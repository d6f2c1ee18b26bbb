// Probe: what limits the fp32-row epilogues (vertex head, FLAME)? 148 CTAs x 512 threads write a [15936 x 15072] fp32 matrix tile by
// tile (128 rows x 256 columns per CTA per tile, tiles ordered n-fastest across CTAs like the GEMM scheduler), with the warp-level
// store shape varied: SEG bytes of one row per instruction (64 = the GEMM epilogue today: 8 rows x 64 B; 128; 256; 512).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/store_probe profiles/probes/store_pattern_probe.cu && /tmp/store_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int SEG>
__global__ void __launch_bounds__(512, 1) probe(float* __restrict__ out, int rows, int ld, int n_tiles, int total_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int LPR = SEG >= 16 ? SEG / 16 : 1;          // lanes per row
  constexpr int RPI = 32 / LPR;          // rows per instruction
  // warp region: SEG <= 256: 32 rows x 64 columns (quarter = warp % 4, slice = warp / 4); SEG == 512: 8 rows x 256 columns
  const float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n_blk = t % n_tiles, m_blk = t / n_tiles;
    const int row0 = m_blk * 128, col0 = n_blk * 256;
    if (SEG == 16) {
      // thread = row (the raw tcgen05.ld 32x32b layout): each lane writes its own row, 16 B per instruction, 8 instructions = 128 B
      const int r = row0 + (warp & 3) * 32 + lane;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = col0 + (warp >> 2) * 64 + ch * 32 + j * 4;
          if (r < rows && c + 4 <= ld) *reinterpret_cast<float4*>(out + (size_t)r * ld + c) = v;
        }
    } else if (SEG <= 256) {
      const int r_base = row0 + (warp & 3) * 32, c_base = col0 + (warp >> 2) * 64;
      constexpr int COLS_PER_PASS = SEG / 4;                 // floats per row per instruction
#pragma unroll
      for (int cp = 0; cp < 64 / COLS_PER_PASS; ++cp)
#pragma unroll
        for (int rp = 0; rp < 32 / RPI; ++rp) {
          const int r = r_base + rp * RPI + lane / LPR, c = c_base + cp * COLS_PER_PASS + (lane % LPR) * 4;
          if (r < rows && c + 4 <= ld) *reinterpret_cast<float4*>(out + (size_t)r * ld + c) = v;
        }
    } else {
      const int r_base = row0 + warp * 8;
#pragma unroll
      for (int rp = 0; rp < 8; ++rp)
#pragma unroll
        for (int cp = 0; cp < 2; ++cp) {
          const int r = r_base + rp, c = col0 + cp * 128 + lane * 4;
          if (r < rows && c + 4 <= ld) *reinterpret_cast<float4*>(out + (size_t)r * ld + c) = v;
        }
    }
  }
}

template <int SEG>
void run(float* d, int rows, int ld) {
  const int n_tiles = (ld + 255) / 256, m_tiles = (rows + 127) / 128;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) probe<SEG><<<148, 512>>>(d, rows, ld, n_tiles, n_tiles * m_tiles);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) probe<SEG><<<148, 512>>>(d, rows, ld, n_tiles, n_tiles * m_tiles);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 10;
  printf("SEG %3d B/row/instr: %.4f ms  %.0f GB/s  (%s)\n", SEG, ms, (double)rows * ld * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int rows = 15936, ld = 15072;
  float* d;
  cudaMalloc(&d, (size_t)rows * ld * 4);
  run<16>(d, rows, ld);
  run<64>(d, rows, ld);
  run<128>(d, rows, ld);
  run<256>(d, rows, ld);
  run<512>(d, rows, ld);
  run<64>(d, rows, ld);
  return 0;
}
