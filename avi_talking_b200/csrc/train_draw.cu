// The stochastic draws of one TRAIN-mode faceformer_vert step, made on the device inside the step's CUDA graph.
//
// The reference draws them inside its modules, one torch RNG call per site, interleaved with the forward (nn.Dropout in HF Wav2Vec2 /
// PeriodicPositionalEncoding / nn.TransformerDecoderLayer, np.random.uniform for LayerDrop in Wav2Vec2Encoder.forward,
// _compute_mask_indices for SpecAugment, models/lib/wav2vec.py:16-63,120-131). Here every dropout site of the step lives in ONE flat
// fp32 buffer that one launch fills (HBM-write bound: 4 B per mask element), a second launch draws the LayerDrop decisions (written as
// the 0 / 1 blend rows the graph-stable LayerDrop of train.py multiplies with) and the SpecAugment spans, a third bumps the step counter.
// The generator is counter based (Philox4x32-10, Salmon et al. 2011), so the draws are a pure function of (seed, step, element index):
// a graph replay needs no host-side RNG state, and oracle/philox_oracle.py reproduces every mask bit for bit.
#include <algorithm>

#include "common.cuh"

namespace avi {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0, 1), 24 bits

// state: uint32[4] = {seed_lo, seed_hi, step, unused}. Element i of the flat buffer: group i / 4, word i % 4 of
// philox(ctr = (group_lo, group_hi, step, stream), key = seed); mask = u >= p ? 1 / (1 - p) : 0.
__global__ void __launch_bounds__(256) dropout_masks_kernel(float4* __restrict__ out, int64_t n4, float p, float scale,
                                                            const uint32_t* __restrict__ state, uint32_t stream) {
  const uint32_t k0 = state[0], k1 = state[1], step = state[2];
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n4; g += (int64_t)gridDim.x * blockDim.x) {
    const Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)((uint64_t)g >> 32), step, stream, k0, k1);
    out[g] = make_float4(u01(r.x) >= p ? scale : 0.f, u01(r.y) >= p ? scale : 0.f, u01(r.z) >= p ? scale : 0.f, u01(r.w) >= p ? scale : 0.f);
  }
}

constexpr uint32_t kStreamLayerDrop = 0x4C440000u, kStreamSpecStart = 0x53500000u, kStreamSpecCount = 0x534E0000u;
constexpr int kMaxLayers = 64;

// LayerDrop (skip layer l when u_l < layerdrop, Wav2Vec2Encoder.forward) as blend rows: blend[0][l][:] = keep, blend[1][l][:] = 1 - keep;
// keep_flags[l] likewise. Block 0 also draws the SpecAugment spans: n = max(min_spans, floor(rate + u)) spans of span_len rows per
// clip, start = floor(u' * (T - span_len + 1)) (models/lib/wav2vec.py:16-63 with min_masks = 2; spans may overlap, as upstream's).
__global__ void __launch_bounds__(256) layerdrop_spec_kernel(float* __restrict__ blend, float* __restrict__ keep_flags, int n_layers, int64_t rows,
                                                             float layerdrop, uint8_t* __restrict__ spec, int B, int T, int span_len,
                                                             float span_rate, int min_spans, const uint32_t* __restrict__ state) {
  __shared__ float keep[kMaxLayers];
  const uint32_t k0 = state[0], k1 = state[1], step = state[2];
  if ((int)threadIdx.x < n_layers) {
    const float u = u01(philox4x32_10(threadIdx.x, 0u, step, kStreamLayerDrop, k0, k1).x);
    keep[threadIdx.x] = u >= layerdrop ? 1.f : 0.f;
    if (blockIdx.x == 0 && keep_flags != nullptr) keep_flags[threadIdx.x] = keep[threadIdx.x];
  }
  __syncthreads();
  if (blend != nullptr) {
    const int64_t per = (int64_t)n_layers * rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
      const float k = keep[i / rows];
      blend[i] = k;
      blend[per + i] = 1.f - k;
    }
  }
  if (blockIdx.x != 0 || spec == nullptr) return;
  for (int i = threadIdx.x; i < B * T; i += blockDim.x) spec[i] = 0;
  __syncthreads();
  if (span_len >= T) return;
  const float un = u01(philox4x32_10(0u, 0u, step, kStreamSpecCount, k0, k1).x);
  int n_spans = (int)floorf(span_rate + un);
  n_spans = n_spans < min_spans ? min_spans : n_spans;
  for (int i = threadIdx.x; i < B * n_spans; i += blockDim.x) {
    const int b = i / n_spans, s = i - b * n_spans;
    const uint32_t x = philox4x32_10((uint32_t)b, (uint32_t)s, step, kStreamSpecStart, k0, k1).x;
    const int start = (int)(((uint64_t)x * (uint64_t)(T - span_len + 1)) >> 32);
    for (int t = 0; t < span_len; ++t) spec[b * T + start + t] = 1;     // overlapping spans write the same byte value
  }
}

__global__ void bump_step_kernel(uint32_t* state) { state[2] += 1u; }

}  // namespace avi

using namespace avi;

extern "C" int avi_dropout_masks(float* out, int64_t n, float p, const uint32_t* state, uint32_t stream_id, void* stream) {
  AVI_REQUIRE(out != nullptr && state != nullptr && n > 0 && n % 4 == 0, "avi_dropout_masks: n must be a positive multiple of 4 (n=%lld)",
              (long long)n);
  AVI_REQUIRE(p >= 0.f && p < 1.f && ((uintptr_t)out % 16) == 0, "avi_dropout_masks: 0 <= p < 1 and a 16-byte aligned buffer");
  AVI_REQUIRE(stream_id < 0x10000u, "avi_dropout_masks: stream ids above 0xFFFF are reserved for LayerDrop / SpecAugment");
  const int64_t n4 = n / 4;
  const unsigned grid = (unsigned)std::min<int64_t>((n4 + 255) / 256, (int64_t)device_sms() * 8);
  dropout_masks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)out, n4, p, 1.0f / (1.0f - p), state, stream_id);
  return check_launch("dropout_masks");
}

extern "C" int avi_layerdrop_spec_draw(float* blend, float* keep_flags, int32_t n_layers, int64_t rows, float layerdrop, uint8_t* spec,
                                       int32_t B, int32_t T, int32_t span_len, float span_rate, int32_t min_spans, const uint32_t* state,
                                       void* stream) {
  AVI_REQUIRE(state != nullptr && n_layers > 0 && n_layers <= kMaxLayers && rows >= 0, "avi_layerdrop_spec_draw: 1..64 layers");
  AVI_REQUIRE(spec == nullptr || (B > 0 && T > 0 && span_len > 0 && min_spans >= 0 && span_rate >= 0.f), "avi_layerdrop_spec_draw: bad span arguments");
  const int64_t per = (int64_t)n_layers * rows;
  const unsigned grid = blend ? (unsigned)std::max<int64_t>(1, std::min<int64_t>((per + 255) / 256, (int64_t)device_sms() * 4)) : 1u;
  layerdrop_spec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(blend, keep_flags, n_layers, rows, layerdrop, spec, B, T, span_len, span_rate,
                                                                min_spans, state);
  return check_launch("layerdrop_spec_draw");
}

extern "C" int avi_draw_bump_step(uint32_t* state, void* stream) {
  AVI_REQUIRE(state != nullptr, "avi_draw_bump_step: null state");
  bump_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
  return check_launch("draw_bump_step");
}
