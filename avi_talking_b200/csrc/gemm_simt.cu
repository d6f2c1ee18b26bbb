// fp32 CUDA-core GEMM (exact mode / odd shapes):  C[b,r,n] = epi(sum_k A[b,r,k] W[n,k] + bias[n]) (+ residual)
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, smem tiles stored k-major so the inner loop is
// two float4 loads + 16 FMAs per k.
#include "common.cuh"

namespace avi {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

struct SimtParams {
  const float* A;
  const float* W;
  const float* bias;
  const float* residual;
  float* C;
  __nv_bfloat16* C2;
  int batch, rows, N, K;
  int64_t a_row_stride, a_batch_stride, c_ld, c_batch_stride, res_ld, res_batch_stride;
  int act;
  int m_tiles;  // per batch entry
  int kx_span;      // 2-D taps: elements of one line of the window (taps_x * C_in); 0 = contiguous window
  int64_t kx_skip;  // elements to skip between the lines of the window ((row_pitch - taps_x) * a_ld)
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const SimtParams p) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Ws[SG_BK][SG_BN + 4];
  const int b = blockIdx.x / p.m_tiles;
  const int m0 = (blockIdx.x % p.m_tiles) * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const float* Ab = p.A + (int64_t)b * p.a_batch_stride;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> n, ty -> m
  // loader mapping: 64 rows x 16 k = 1024 elements / 256 threads = 4 each (one float4 along k when aligned)
  const int lr = tid / 4, lk = (tid % 4) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += SG_BK) {
    {
      const int r = m0 + lr;
      const float* src = Ab + (int64_t)r * p.a_row_stride + k0 + lk + (p.kx_span ? (int64_t)((k0 + lk) / p.kx_span) * p.kx_skip : 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) As[lk + u][lr] = (r < p.rows && k0 + lk + u < p.K) ? src[u] : 0.f;
      const int n = n0 + lr;
      const float* wsrc = p.W + (int64_t)n * p.K + k0 + lk;
#pragma unroll
      for (int u = 0; u < 4; ++u) Ws[lk + u][lr] = (n < p.N && k0 + lk + u < p.K) ? wsrc[u] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= p.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
      if (p.act == AVI_ACT_GELU) v = gelu_erf(v);
      else if (p.act == AVI_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == AVI_ACT_QUICK_GELU) v = quick_gelu(v);
      if (p.residual) v += p.residual[(int64_t)b * p.res_batch_stride + (int64_t)r * p.res_ld + n];
      const int64_t o = (int64_t)b * p.c_batch_stride + (int64_t)r * p.c_ld + n;
      if (p.C) p.C[o] = v;
      if (p.C2) p.C2[o] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace avi

using namespace avi;

extern "C" int avi_gemm_f32(const AviGemmArgs* a, void* stream) {
  AVI_REQUIRE(a != nullptr, "avi_gemm_f32: null args");
  AVI_REQUIRE(a->a_dtype == AVI_DT_F32, "avi_gemm_f32: A/W must be fp32");
  AVI_REQUIRE(a->batch > 0 && a->rows > 0 && a->N > 0 && a->K > 0, "avi_gemm_f32: bad shape b=%d r=%d N=%d K=%d", a->batch,
              a->rows, a->N, a->K);
  AVI_REQUIRE(a->conv_taps >= 1 && a->conv_stride >= 1, "avi_gemm_f32: bad conv params");
  AVI_REQUIRE(a->conv_taps == 1 || a->K == a->conv_taps * a->a_ld,
              "avi_gemm_f32: conv mode needs contiguous input rows (K == taps * a_ld)");
  AVI_REQUIRE(a->conv_taps_x == 0 || (a->conv_taps_x > 0 && a->conv_stride == 1 && a->conv_taps % a->conv_taps_x == 0 &&
                                      a->conv_row_pitch >= a->conv_taps_x && (a->conv_taps_x * a->a_ld) % 4 == 0),
              "avi_gemm_f32: 2-D taps need conv_stride 1, conv_taps a multiple of conv_taps_x, line span a multiple of 4 elements");
  SimtParams p;
  p.A = (const float*)a->A;
  p.W = (const float*)a->W;
  p.bias = a->bias;
  p.residual = a->residual;
  p.C = a->c_dtype == AVI_DT_F32 ? (float*)a->C : (float*)a->C2;
  p.C2 = a->c_dtype == AVI_DT_BF16 ? (__nv_bfloat16*)a->C : (__nv_bfloat16*)a->C2;
  p.batch = a->batch;
  p.rows = a->rows;
  p.N = a->N;
  p.K = a->K;
  p.a_row_stride = a->a_ld * a->conv_stride;
  p.a_batch_stride = a->a_batch_stride;
  p.c_ld = a->c_ld;
  p.c_batch_stride = a->c_batch_stride;
  p.res_ld = a->res_ld;
  p.res_batch_stride = a->res_batch_stride;
  p.act = a->act;
  p.m_tiles = (a->rows + SG_BM - 1) / SG_BM;
  p.kx_span = a->conv_taps_x != 0 ? a->conv_taps_x * (int)a->a_ld : 0;
  p.kx_skip = a->conv_taps_x != 0 ? (int64_t)(a->conv_row_pitch - a->conv_taps_x) * a->a_ld : 0;
  dim3 grid((unsigned)(p.m_tiles * a->batch), (unsigned)((a->N + SG_BN - 1) / SG_BN));
  AVI_REQUIRE(grid.y <= 65535, "avi_gemm_f32: N too large");
  gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("gemm_f32");
}
