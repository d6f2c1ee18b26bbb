// Shared helpers for the libavi_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/avi_b200.h"

namespace avi {

constexpr int kNumSMs = 148;  // B200 (upper bound used for static sizing; launches size their grids with device_sms())

// SM count of the CURRENT device, cached per device ordinal (a process may drive several GPUs: models can be moved with .to()).
inline int device_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// Opt-in to more than 48 KB of dynamic shared memory. The attribute belongs to (function, DEVICE): one flag per device ordinal,
// holding the largest size granted there. `state` is a zero-initialised static owned by the launch site (one per kernel instance).
struct SmemOptIn {
  std::atomic<int> granted[64];
};
template <typename Kernel>
inline cudaError_t smem_optin(Kernel kernel, int bytes, SmemOptIn& state) {
  int dev = 0;
  const bool cacheable = cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64;
  if (cacheable && state.granted[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && cacheable) state.granted[dev].store(bytes, std::memory_order_release);
  return e;
}

// Which persistent kernels take their tiles through cluster launch control instead of a static walk (avi_set_dynamic_tiles).
enum : int { AVI_DYN_GEMM = 1, AVI_DYN_CONV0 = 2 };
int dynamic_tiles_mask();

// Programmatic dependent launch (griddepcontrol): a kernel launched with launch_pdl() may be scheduled while the previous kernel of the
// stream is still draining (its CTAs have all exited or are in their last wave), runs its prologue (barrier init, TMEM allocation,
// tensor-map prefetch) and blocks in pdl_wait() until that kernel's memory is complete and visible. Only kernels that call pdl_wait()
// before their first global access may be launched this way. AVI_PDL=0 in the environment / avi_set_pdl(0) = plain launches.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

#define AVI_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      avi::set_error(__VA_ARGS__);  \
      return 2;                     \
    }                               \
  } while (0)

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Branch-free GELU for epilogues whose result is rounded to bf16 anyway (10 instructions, ONE of them MUFU; round 1 used the
// Abramowitz-Stegun erfc form: 14 instructions, two MUFU - the GELU epilogues of conv0 / conv1..6 / ffn1 are issue-bound):
//   gelu(x) = relu(x) - |x| * h(|x|),   h(a) = 0.5 * erfc(a / sqrt 2) = 2^P(a)
// P = degree-6 minimax fit of log2 h on [0, 5.5] (profiles/fit_gelu_poly.py): relative error of h <= 2.5e-5, |a h| error <= 3.8e-6 -
// three orders below the bf16 rounding of the result; no cancellation on either tail. Beyond 5.5 the argument is clamped
// (h(5.5) = 1.9e-8: the term is below 2e-6 for |x| < 100).
__device__ __forceinline__ float gelu_fast(float x) {
  const float a = fminf(fabsf(x), 5.5f);
  float pl = fmaf(2.615383824e-05f, a, -6.609828710e-04f);
  pl = fmaf(pl, a, 7.488321837e-03f);
  pl = fmaf(pl, a, -5.197044650e-02f);
  pl = fmaf(pl, a, -4.603294121e-01f);
  pl = fmaf(pl, a, -1.150584037e+00f);
  pl = fmaf(pl, a, -1.000036059e+00f);
  float h;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(pl));
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}

// Two GELUs per instruction stream: sm_100 has packed fp32 arithmetic (fma.rn.f32x2 -> FFMA2, two FMAs per lane per issue slot), and the
// GELU epilogues (2.7 G per 64-clip step; conv0 is issue-bound on them) are FMA chains. The polynomial runs in na = -min(|x|, 5.5) with
// the odd coefficients negated - every intermediate is the exact negation or the same value as in gelu_fast, so the result is
// bit-identical for |x| <= 5.5 (beyond, the 1e-7 tail term uses the clamped argument). 7 FFMA2 + 2 MUFU + 4 FMNMX per pair instead of
// 14 FFMA + 2 MUFU + 4 FMNMX (+ 2 for the separate -|x|).
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void gelu_fast2(float x0, float x1, float& y0, float& y1) {
  const unsigned long long na = pack_f32x2(fmaxf(-fabsf(x0), -5.5f), fmaxf(-fabsf(x1), -5.5f));
  unsigned long long pl = fma_f32x2(pack_f32x2(2.615383824e-05f, 2.615383824e-05f), na, pack_f32x2(6.609828710e-04f, 6.609828710e-04f));
  pl = fma_f32x2(pl, na, pack_f32x2(7.488321837e-03f, 7.488321837e-03f));
  pl = fma_f32x2(pl, na, pack_f32x2(5.197044650e-02f, 5.197044650e-02f));
  pl = fma_f32x2(pl, na, pack_f32x2(-4.603294121e-01f, -4.603294121e-01f));
  pl = fma_f32x2(pl, na, pack_f32x2(1.150584037e+00f, 1.150584037e+00f));
  pl = fma_f32x2(pl, na, pack_f32x2(-1.000036059e+00f, -1.000036059e+00f));
  float p0, p1, h0, h1;
  unpack_f32x2(pl, p0, p1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(p0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(p1));
  unpack_f32x2(fma_f32x2(na, pack_f32x2(h0, h1), pack_f32x2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), y0, y1);
}

// CLIP's quick_gelu: x * sigmoid(1.702 x)
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.f + __expf(-1.702f * x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of two values (blockDim.x multiple of 32, <= 1024). red must hold 64 floats.
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) {
    red[w] = a;
    red[32 + w] = b;
  }
  __syncthreads();
  a = (l < nw) ? red[l] : 0.f;
  b = (l < nw) ? red[32 + l] : 0.f;
  a = warp_sum(a);
  b = warp_sum(b);
}

__device__ __forceinline__ float load_as_float(const void* p, int dtype, int64_t i) {
  return dtype == AVI_DT_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                              : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_from_float(void* p, int dtype, int64_t i, float v) {
  if (dtype == AVI_DT_BF16)
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(p)[i] = v;
}

}  // namespace avi
