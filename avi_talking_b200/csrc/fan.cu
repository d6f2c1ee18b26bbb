// Row kernels of the FanEncoder image branch (third_party/pd_fgc_inference/lib/models/networks/
// FAN_feature_extractor.py:13-163): activations are NHWC fp32 rows [N*H*W, C]; every convolution is avi_im2col_affine + the GEMM.
#include "common.cuh"

namespace avi {

template <typename OutT>
__device__ __forceinline__ OutT fan_out(float v);
template <>
__device__ __forceinline__ float fan_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 fan_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// cols[(n, oy, ox), (ky, kx, c)] = act(x[n, oy*s - p + ky, ox*s - p + kx, c]) (zero outside the image: F.conv2d pads the ACTIVATED
// tensor), act(v) = relu(v * scale[c] + shift[c]) when scale != nullptr (the pre-activation BatchNorm + ReLU of ConvBlock), else v.
// Columns K..Kpad-1 are zero. One thread per 4 consecutive channels of one tap (C % 4 == 0) or per element otherwise.
// round-to-nearest to the 10-bit TF32 significand (the kind::tf32 MMA ignores the low 13 bits of its fp32 operands)
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

template <typename OutT, int VEC, bool RTF32 = false>
__global__ void __launch_bounds__(256) im2col_affine_kernel(const float* __restrict__ x, int64_t x_ld, OutT* __restrict__ cols, int N, int H,
                                                            int W, int C, int k, int stride, int pad, int Ho, int Wo, int Kpad,
                                                            const float* __restrict__ scale, const float* __restrict__ shift, int Wp_in) {
  const int KV = Kpad / VEC;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * Ho * Wo * KV) return;
  const int kc = (int)(i % KV) * VEC;
  const int64_t row = i / KV;
  const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), n = (int)(row / ((int64_t)Wo * Ho));
  float v[VEC];
#pragma unroll
  for (int u = 0; u < VEC; ++u) v[u] = 0.f;
  if (kc < k * k * C) {
    const int tap = kc / C, c = kc - tap * C;
    const int ky = tap / k, kx = tap - ky * k;
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float* src = x + ((int64_t)(n * H + iy) * Wp_in + ix) * x_ld + c;
#pragma unroll
      for (int u = 0; u < VEC; ++u) {
        float t = src[u];
        if (scale) t = fmaxf(fmaf(t, scale[c + u], shift[c + u]), 0.f);
        v[u] = RTF32 ? round_tf32(t) : t;
      }
    }
  }
  OutT* dst = cols + row * Kpad + kc;
#pragma unroll
  for (int u = 0; u < VEC; ++u) dst[u] = fan_out<OutT>(v[u]);
}

// Operand of the implicit 3x3 convolution: A[n, y+1, x+1, c] = act(x[n, y, x, c]) for the valid pixels, zero on the one-pixel border,
// stored with the SAME line pitch Wp = W + 2 as the activations (x has Wp pixels per line, the last two are don't-care), so that
// output pixel r = y*Wp + x' and tap (ky, kx) read operand row r + ky*Wp + kx: three conv-mode GEMM launches (taps = kx) per 3x3.
template <typename OutT, bool RTF32>
__global__ void __launch_bounds__(256) pad_act_kernel(const float* __restrict__ x, int64_t x_ld, OutT* __restrict__ a, int N, int H, int W,
                                                      int C, const float* __restrict__ scale, const float* __restrict__ shift) {
  const int Wp = W + 2, C4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * (H + 2) * Wp * C4) return;
  const int c = (int)(i % C4) * 4;
  const int64_t r = i / C4;
  const int xx = (int)(r % Wp), yy = (int)((r / Wp) % (H + 2)), n = (int)(r / ((int64_t)Wp * (H + 2)));
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (yy >= 1 && yy <= H && xx >= 1 && xx <= W) {
    const float* src = x + ((int64_t)(n * H + yy - 1) * Wp + xx - 1) * x_ld + c;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float t = src[u];
      if (scale) t = fmaxf(fmaf(t, scale[c + u], shift[c + u]), 0.f);
      v[u] = RTF32 ? round_tf32(t) : t;
    }
  }
  OutT* dst = a + r * C + c;
#pragma unroll
  for (int u = 0; u < 4; ++u) dst[u] = fan_out<OutT>(v[u]);
}

// F.max_pool2d(x, 2, stride=2) on NHWC rows (floor: a trailing odd row / column is dropped); line pitches Wp_in / Wp_out pixels
__global__ void maxpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int Wp_in, int Wp_out) {
  const int Ho = H / 2, Wo = W / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * Ho * Wo * C) return;
  const int c = (int)(i % C);
  const int64_t r = i / C;
  const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), n = (int)(r / ((int64_t)Wo * Ho));
  const float* p = x + ((int64_t)(n * H + 2 * oy) * Wp_in + 2 * ox) * C + c;
  y[((int64_t)(n * Ho + oy) * Wp_out + ox) * C + c] =
      fmaxf(fmaxf(p[0], p[C]), fmaxf(p[(int64_t)Wp_in * C], p[(int64_t)Wp_in * C + C]));
}

// out = up1 + F.interpolate(low, size=(Ho, Wo), mode="bilinear", align_corners=False)   (HourGlass._forward :97-101)
__global__ void upsample_bilinear_add_kernel(const float* __restrict__ low, const float* __restrict__ up1, float* __restrict__ out, int N,
                                             int Hi, int Wi, int Ho, int Wo, int C, int Wp_in, int Wp_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * Ho * Wo * C) return;
  const int c = (int)(i % C);
  const int64_t r = i / C;
  const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), n = (int)(r / ((int64_t)Wo * Ho));
  // PyTorch area_pixel_compute_source_index (align_corners = False): src = max((dst + 0.5) * (in / out) - 0.5, 0)
  const float sy = fmaxf(((float)oy + 0.5f) * ((float)Hi / (float)Ho) - 0.5f, 0.f);
  const float sx = fmaxf(((float)ox + 0.5f) * ((float)Wi / (float)Wo) - 0.5f, 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float* b = low + (int64_t)n * Hi * Wp_in * C + c;
  const float v00 = b[((int64_t)y0 * Wp_in + x0) * C], v01 = b[((int64_t)y0 * Wp_in + x1) * C];
  const float v10 = b[((int64_t)y1 * Wp_in + x0) * C], v11 = b[((int64_t)y1 * Wp_in + x1) * C];
  const int64_t o = ((int64_t)(n * Ho + oy) * Wp_out + ox) * C + c;
  out[o] = up1[o] + (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// x[r, c] = act(x[r, c] * scale[c] + shift[c]) in place (scale == nullptr: activation only); eval BatchNorm folded after a GEMM
__global__ void affine_act_kernel(float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift, int64_t n, int C,
                                  int relu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  float v = x[i];
  if (scale) v = fmaf(v, scale[c], shift[c]);
  x[i] = relu ? fmaxf(v, 0.f) : v;
}

}  // namespace avi

using namespace avi;

extern "C" int avi_im2col_affine(const float* x, int64_t x_ld, void* cols, int32_t cols_dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                                 int32_t k, int32_t stride, int32_t pad, int32_t Kpad, const float* scale, const float* shift, int32_t Wp_in,
                                 int32_t Wo_extra, void* stream) {
  AVI_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && pad >= 0 && Kpad >= k * k * C && x_ld >= C && Wp_in >= W &&
                  Wo_extra >= 0,
              "avi_im2col_affine: bad shape");
  // Wo_extra surplus output columns per image line (the padded-width layout of the implicit 3x3 convolutions): computed like any
  // other column, their taps beyond W read as zero
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1 + Wo_extra;
  AVI_REQUIRE(Ho > 0 && Wo > 0, "avi_im2col_affine: empty output");
  const bool vec = (C % 4 == 0) && (Kpad % 4 == 0);
  const int64_t n = (int64_t)N * Ho * Wo * (vec ? Kpad / 4 : Kpad);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (cols_dtype == AVI_DT_BF16) {
    if (vec) im2col_affine_kernel<__nv_bfloat16, 4><<<blocks, 256, 0, st>>>(x, x_ld, (__nv_bfloat16*)cols, N, H, W, C, k, stride, pad, Ho, Wo, Kpad, scale, shift, Wp_in);
    else im2col_affine_kernel<__nv_bfloat16, 1><<<blocks, 256, 0, st>>>(x, x_ld, (__nv_bfloat16*)cols, N, H, W, C, k, stride, pad, Ho, Wo, Kpad, scale, shift, Wp_in);
  } else if (cols_dtype == AVI_DT_TF32) {
    if (vec) im2col_affine_kernel<float, 4, true><<<blocks, 256, 0, st>>>(x, x_ld, (float*)cols, N, H, W, C, k, stride, pad, Ho, Wo, Kpad, scale, shift, Wp_in);
    else im2col_affine_kernel<float, 1, true><<<blocks, 256, 0, st>>>(x, x_ld, (float*)cols, N, H, W, C, k, stride, pad, Ho, Wo, Kpad, scale, shift, Wp_in);
  } else {
    if (vec) im2col_affine_kernel<float, 4><<<blocks, 256, 0, st>>>(x, x_ld, (float*)cols, N, H, W, C, k, stride, pad, Ho, Wo, Kpad, scale, shift, Wp_in);
    else im2col_affine_kernel<float, 1><<<blocks, 256, 0, st>>>(x, x_ld, (float*)cols, N, H, W, C, k, stride, pad, Ho, Wo, Kpad, scale, shift, Wp_in);
  }
  return check_launch("im2col_affine");
}

extern "C" int avi_maxpool2x2(const float* x, float* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t Wp_in, int32_t Wp_out,
                              void* stream) {
  AVI_REQUIRE(N > 0 && H >= 2 && W >= 2 && C > 0 && Wp_in >= W && Wp_out >= W / 2, "avi_maxpool2x2: bad shape");
  const int64_t n = (int64_t)N * (H / 2) * (W / 2) * C;
  maxpool2x2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C, Wp_in, Wp_out);
  return check_launch("maxpool2x2");
}

extern "C" int avi_pad_act(const float* x, int64_t x_ld, void* a, int32_t a_dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                           const float* scale, const float* shift, void* stream) {
  AVI_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && x_ld >= C && ((scale == nullptr) == (shift == nullptr)),
              "avi_pad_act: bad shape (C must be a multiple of 4)");
  const int64_t n = (int64_t)N * (H + 2) * (W + 2) * (C / 4);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (a_dtype == AVI_DT_BF16) pad_act_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>(x, x_ld, (__nv_bfloat16*)a, N, H, W, C, scale, shift);
  else if (a_dtype == AVI_DT_TF32) pad_act_kernel<float, true><<<blocks, 256, 0, st>>>(x, x_ld, (float*)a, N, H, W, C, scale, shift);
  else pad_act_kernel<float, false><<<blocks, 256, 0, st>>>(x, x_ld, (float*)a, N, H, W, C, scale, shift);
  return check_launch("pad_act");
}

extern "C" int avi_upsample_bilinear_add(const float* low, const float* up1, float* out, int32_t N, int32_t Hi, int32_t Wi, int32_t Ho,
                                         int32_t Wo, int32_t C, int32_t Wp_in, int32_t Wp_out, void* stream) {
  AVI_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0 && Wp_in >= Wi && Wp_out >= Wo, "avi_upsample_bilinear_add: bad shape");
  const int64_t n = (int64_t)N * Ho * Wo * C;
  upsample_bilinear_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(low, up1, out, N, Hi, Wi, Ho, Wo, C, Wp_in,
                                                                                             Wp_out);
  return check_launch("upsample_bilinear_add");
}

extern "C" int avi_affine_act(float* x, const float* scale, const float* shift, int64_t rows, int32_t C, int32_t relu, void* stream) {
  AVI_REQUIRE(rows > 0 && C > 0 && ((scale == nullptr) == (shift == nullptr)), "avi_affine_act: bad arguments");
  const int64_t n = rows * C;
  affine_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, scale, shift, n, C, relu);
  return check_launch("affine_act");
}
