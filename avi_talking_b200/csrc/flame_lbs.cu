// FLAME blendshapes + pose correctives + 5-joint kinematic chain + linear blend skinning (fp32 CUDA-core version).
// Follows third_party/inferno/inferno/utils/lbs.py:142-234 (lbs), :304-335 (batch_rodrigues), :351-408
// (batch_rigid_transform) and DecaFLAME.py:222-269.  Joint regression is folded into a [15, NB+1] matrix at pack time
// (J = J_regressor (v_template + S betas) = JT + JS betas), so no per-frame reduction over vertices is needed.
#include <cuda_fp16.h>

#include "common.cuh"

namespace avi {

constexpr int FL_NJ = 5;
constexpr int FL_FT = 64, FL_VT = 64, FL_KC = 16, FL_FPT = 16;  // frame tile, vertex tile, k chunk, frames per thread

// dirs[l][v*3+k] = shapedirs[v][k][l] ; rows NB..NB+35 = posedirs ; row NB+36 = v_template ; remaining rows zero
__global__ void flame_pack_dirs_kernel(const float* __restrict__ shapedirs, const float* __restrict__ posedirs,
                                       const float* __restrict__ v_template, float* __restrict__ dirs, int V3, int NB, int K_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)K_pad * V3) return;
  const int l = (int)(i / V3), col = (int)(i % V3);
  float v = 0.f;
  if (l < NB) v = shapedirs[(int64_t)col * NB + l];
  else if (l < NB + 36) v = posedirs[(int64_t)(l - NB) * V3 + col];
  else if (l == NB + 36) v = v_template[col];
  dirs[i] = v;
}

// jreg[j*3+k][l] = sum_v Jreg[j][v] * shapedirs[v][k][l]  (l < NB) ; column NB uses v_template
__global__ void flame_pack_jreg_kernel(const float* __restrict__ shapedirs, const float* __restrict__ v_template,
                                       const float* __restrict__ Jreg, float* __restrict__ jreg, int V, int NB) {
  const int l = blockIdx.x;   // 0..NB
  const int jk = blockIdx.y;  // 0..14
  const int j = jk / 3, k = jk % 3;
  double acc = 0.0;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float d = (l < NB) ? shapedirs[((int64_t)v * 3 + k) * NB + l] : v_template[v * 3 + k];
    acc += (double)Jreg[(int64_t)j * V + v] * (double)d;
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) jreg[(int64_t)jk * (NB + 1) + l] = (float)red[0];
}

__device__ __forceinline__ void rodrigues(const float* r, float* R) {
  // lbs.py:304-335: angle = || r + 1e-8 ||, axis = r / angle
  const float x = r[0] + 1e-8f, y = r[1] + 1e-8f, z = r[2] + 1e-8f;
  const float angle = sqrtf(x * x + y * y + z * z);
  const float rx = r[0] / angle, ry = r[1] / angle, rz = r[2] / angle;
  float s, c;
  sincosf(angle, &s, &c);
  const float oc = 1.f - c;
  // K = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]] ; R = I + s K + (1-c) K^2
  const float K[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float kk = 0.f;
#pragma unroll
      for (int m = 0; m < 3; ++m) kk = fmaf(K[i * 3 + m], K[m * 3 + j], kk);
      R[i * 3 + j] = (i == j ? 1.f : 0.f) + s * K[i * 3 + j] + oc * kk;
    }
}

// One warp per frame: coefficient row (betas | pose feature | 1), joints, kinematic chain, relative transforms A [5][3x4].
// rotmat = 1: full_pose holds the five 3x3 rotation matrices of a frame (lbs(pose2rot=False), lbs.py:205-209) instead of 15 axis-angle
// components.
__global__ void __launch_bounds__(128) flame_prologue_kernel(const float* __restrict__ betas, const float* __restrict__ full_pose,
                                                             const float* __restrict__ jreg, float* __restrict__ coef,
                                                             float* __restrict__ A, float* __restrict__ joints,
                                                             int32_t* __restrict__ dyn_rows, int F, int NB, int K_pad, int rotmat) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= F) return;
  __shared__ float sJ[4][16];
  __shared__ float sR[4][FL_NJ * 9];
  float* J = sJ[threadIdx.x >> 5];
  float* Rs = sR[threadIdx.x >> 5];
  float* cf = coef + (int64_t)f * K_pad;
  for (int l = lane; l < K_pad; l += 32) {
    if (l < NB) cf[l] = betas[(int64_t)f * NB + l];
    else if (l == NB + 36) cf[l] = 1.f;
    else if (l > NB + 36) cf[l] = 0.f;
  }
  {
    // joints = jreg [15, NB+1] . [betas; 1]: the NB-long contraction is split over the lanes (coalesced reads of both operands),
    // 15 butterfly reductions finish it; the five Rodrigues rotations run on lanes 0..4 meanwhile
    const float* bt = betas + (int64_t)f * NB;
    float part[15];
#pragma unroll
    for (int c = 0; c < 15; ++c) part[c] = 0.f;
    for (int l = lane; l < NB; l += 32) {
      const float bv = bt[l];
#pragma unroll
      for (int c = 0; c < 15; ++c) part[c] = fmaf(jreg[(int64_t)c * (NB + 1) + l], bv, part[c]);
    }
#pragma unroll
    for (int c = 0; c < 15; ++c) {
      const float tot = warp_sum(part[c]);
      if (lane == c) J[c] = tot + jreg[(int64_t)c * (NB + 1) + NB];
    }
    if (lane < FL_NJ) {
      float Rl[9];
      if (rotmat) {
#pragma unroll
        for (int e = 0; e < 9; ++e) Rl[e] = full_pose[(int64_t)f * 45 + lane * 9 + e];
      } else {
        rodrigues(full_pose + (int64_t)f * 15 + lane * 3, Rl);
      }
#pragma unroll
      for (int e = 0; e < 9; ++e) Rs[lane * 9 + e] = Rl[e];
    }
  }
  __syncwarp();
  if (lane == 0) {
    float R[FL_NJ][9];
#pragma unroll
    for (int j = 0; j < FL_NJ; ++j)
#pragma unroll
      for (int e = 0; e < 9; ++e) R[j][e] = Rs[j * 9 + e];
    if (dyn_rows) {
      // DecaFLAME.py:110-149 with neck_kin_chain = [neck, global]: rel = R_global R_neck; yaw in degrees picks the contour row
      float r00 = 0.f, r10 = 0.f, r20 = 0.f;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        r00 = fmaf(R[0][0 * 3 + m], R[1][m * 3 + 0], r00);
        r10 = fmaf(R[0][1 * 3 + m], R[1][m * 3 + 0], r10);
        r20 = fmaf(R[0][2 * 3 + m], R[1][m * 3 + 0], r20);
      }
      const float sy = sqrtf(r00 * r00 + r10 * r10);
      const float deg = fminf(atan2f(-r20, sy) * 180.0f / 3.14159265358979323846f, 39.f);
      const int y = (int)rintf(deg);
      dyn_rows[f] = (y < 0) ? ((y < -39) ? 78 : 39 - y) : y;
    }
    // pose feature = (R[1:] - I) flattened (lbs.py:201)
#pragma unroll
    for (int j = 1; j < FL_NJ; ++j)
#pragma unroll
      for (int e = 0; e < 9; ++e) cf[NB + (j - 1) * 9 + e] = R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    // chain: G_0 = [R_0 | J_0] ; G_j = G_parent [R_j | J_j - J_parent], parents = [-1,0,1,1,1]
    float G[FL_NJ][12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int c = 0; c < 3; ++c) G[0][i * 4 + c] = R[0][i * 3 + c];
      G[0][i * 4 + 3] = J[i];
    }
#pragma unroll
    for (int j = 1; j < FL_NJ; ++j) {
      const int par = (j == 1) ? 0 : 1;
      float rel[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) rel[i] = J[j * 3 + i] - J[par * 3 + i];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float acc = 0.f;
#pragma unroll
          for (int m = 0; m < 3; ++m) acc = fmaf(G[par][i * 4 + m], R[j][m * 3 + c], acc);
          G[j][i * 4 + c] = acc;
        }
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < 3; ++m) acc = fmaf(G[par][i * 4 + m], rel[m], acc);
        G[j][i * 4 + 3] = acc + G[par][i * 4 + 3];
      }
    }
    // A_j = G_j with translation minus G_j[:3,:3] J_j (lbs.py:404-406)
    float* Af = A + (int64_t)f * FL_NJ * 12;
#pragma unroll
    for (int j = 0; j < FL_NJ; ++j) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int m = 0; m < 3; ++m) acc = fmaf(G[j][i * 4 + m], J[j * 3 + m], acc);
#pragma unroll
        for (int c = 0; c < 3; ++c) Af[j * 12 + i * 4 + c] = G[j][i * 4 + c];
        Af[j * 12 + i * 4 + 3] = G[j][i * 4 + 3] - acc;
        if (joints) joints[((int64_t)f * FL_NJ + j) * 3 + i] = G[j][i * 4 + 3];
      }
    }
  }
}

// verts[f,v,:] = sum_j W[v,j] A[f,j] [v_posed(f,v); 1],  v_posed = coef[f,:] . dirs[:, v*3 .. v*3+2]
__global__ void __launch_bounds__(256) flame_blend_skin_kernel(const float* __restrict__ coef, const float* __restrict__ dirs,
                                                               const float* __restrict__ A, const float* __restrict__ lbs_w,
                                                               float* __restrict__ verts, int F, int V, int K_pad) {
  __shared__ __align__(16) float ds[FL_KC][FL_VT * 3];
  __shared__ __align__(16) float cs[FL_KC][FL_FT];
  __shared__ __align__(16) float As[FL_FT][FL_NJ * 12];
  const int v0 = blockIdx.x * FL_VT, f0 = blockIdx.y * FL_FT;
  const int tid = threadIdx.x;
  const int vl = tid % FL_VT, fg = tid / FL_VT;  // 64 vertices x 4 frame groups of 16
  const int V3 = V * 3;
  float acc[FL_FPT][3];
#pragma unroll
  for (int u = 0; u < FL_FPT; ++u) acc[u][0] = acc[u][1] = acc[u][2] = 0.f;
  for (int i = tid; i < FL_FT * FL_NJ * 12; i += 256) {
    const int fr = i / (FL_NJ * 12);
    (&As[0][0])[i] = (f0 + fr < F) ? A[(int64_t)(f0 + fr) * FL_NJ * 12 + (i % (FL_NJ * 12))] : 0.f;
  }
  for (int k0 = 0; k0 < K_pad; k0 += FL_KC) {
    __syncthreads();
    for (int i = tid; i < FL_KC * FL_VT * 3; i += 256) {
      const int kk = i / (FL_VT * 3), c = i % (FL_VT * 3);
      const int col = v0 * 3 + c;
      ds[kk][c] = (col < V3) ? dirs[(int64_t)(k0 + kk) * V3 + col] : 0.f;
    }
    for (int i = tid; i < FL_KC * FL_FT; i += 256) {
      const int fr = i / FL_KC, kk = i % FL_KC;
      cs[kk][fr] = (f0 + fr < F) ? coef[(int64_t)(f0 + fr) * K_pad + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < FL_KC; ++kk) {
      const float d0 = ds[kk][vl * 3 + 0], d1 = ds[kk][vl * 3 + 1], d2 = ds[kk][vl * 3 + 2];
      const float4* cp = reinterpret_cast<const float4*>(&cs[kk][fg * FL_FPT]);
#pragma unroll
      for (int q = 0; q < FL_FPT / 4; ++q) {
        const float4 c = cp[q];
        const float cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[q * 4 + u][0] = fmaf(cv[u], d0, acc[q * 4 + u][0]);
          acc[q * 4 + u][1] = fmaf(cv[u], d1, acc[q * 4 + u][1]);
          acc[q * 4 + u][2] = fmaf(cv[u], d2, acc[q * 4 + u][2]);
        }
      }
    }
  }
  const int v = v0 + vl;
  if (v >= V) return;
  float w[FL_NJ];
#pragma unroll
  for (int j = 0; j < FL_NJ; ++j) w[j] = lbs_w[(int64_t)v * FL_NJ + j];
#pragma unroll
  for (int u = 0; u < FL_FPT; ++u) {
    const int fr = fg * FL_FPT + u;
    const int f = f0 + fr;
    if (f >= F) break;
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
    for (int j = 0; j < FL_NJ; ++j) {
      const float4* ap = reinterpret_cast<const float4*>(&As[fr][j * 12]);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 a = ap[q];
        T[q * 4 + 0] = fmaf(w[j], a.x, T[q * 4 + 0]);
        T[q * 4 + 1] = fmaf(w[j], a.y, T[q * 4 + 1]);
        T[q * 4 + 2] = fmaf(w[j], a.z, T[q * 4 + 2]);
        T[q * 4 + 3] = fmaf(w[j], a.w, T[q * 4 + 3]);
      }
    }
    float* o = verts + ((int64_t)f * V + v) * 3;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      o[i] = fmaf(T[i * 4 + 0], acc[u][0], fmaf(T[i * 4 + 1], acc[u][1], fmaf(T[i * 4 + 2], acc[u][2], T[i * 4 + 3])));
  }
}

__global__ void flame_landmarks_kernel(const float* __restrict__ verts, const int64_t* __restrict__ faces,
                                       const int64_t* __restrict__ idx, const float* __restrict__ bary, float* __restrict__ out,
                                       int F, int V, int L, int per_frame) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F * L) return;
  const int f = i / L, l = i % L;
  const int64_t src = per_frame ? (int64_t)f * L + l : l;
  const int64_t face = idx[src];
  float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int64_t vi = faces[face * 3 + c];
    const float bw = bary[src * 3 + c];
    const float* vp = verts + ((int64_t)f * V + vi) * 3;
    o[0] = fmaf(bw, vp[0], o[0]);
    o[1] = fmaf(bw, vp[1], o[1]);
    o[2] = fmaf(bw, vp[2], o[2]);
  }
  out[(int64_t)i * 3 + 0] = o[0];
  out[(int64_t)i * 3 + 1] = o[1];
  out[(int64_t)i * 3 + 2] = o[2];
}

}  // namespace avi

using namespace avi;

extern "C" int avi_flame_pack(const float* shapedirs, const float* posedirs, const float* v_template, const float* J_regressor,
                              float* dirs, float* jreg, int32_t V, int32_t NB, int32_t K_pad, void* stream) {
  AVI_REQUIRE(V > 0 && NB > 0 && K_pad >= NB + 37 && K_pad % FL_KC == 0, "avi_flame_pack: K_pad must be >= NB+37 and a multiple of %d",
              FL_KC);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)K_pad * V * 3;
  flame_pack_dirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(shapedirs, posedirs, v_template, dirs, V * 3, NB, K_pad);
  if (check_launch("flame_pack_dirs")) return 1;
  flame_pack_jreg_kernel<<<dim3(NB + 1, 15), 256, 0, st>>>(shapedirs, v_template, J_regressor, jreg, V, NB);
  return check_launch("flame_pack_jreg");
}

extern "C" int avi_flame_prologue_ex(const float* betas, const float* full_pose, int32_t pose_is_rotmat, const float* jreg, float* coef,
                                     float* A, float* joints, int32_t* dyn_rows, int32_t F, int32_t NB, int32_t K_pad, void* stream) {
  AVI_REQUIRE(F > 0 && NB > 0 && K_pad >= NB + 37, "avi_flame_prologue: bad shape F=%d NB=%d K_pad=%d", F, NB, K_pad);
  flame_prologue_kernel<<<(F + 3) / 4, 128, 0, (cudaStream_t)stream>>>(betas, full_pose, jreg, coef, A, joints, dyn_rows, F, NB, K_pad,
                                                                       pose_is_rotmat ? 1 : 0);
  return check_launch("flame_prologue");
}

extern "C" int avi_flame_prologue(const float* betas, const float* full_pose, const float* jreg, float* coef, float* A, float* joints,
                                  int32_t* dyn_rows, int32_t F, int32_t NB, int32_t K_pad, void* stream) {
  return avi_flame_prologue_ex(betas, full_pose, 0, jreg, coef, A, joints, dyn_rows, F, NB, K_pad, stream);
}

extern "C" int avi_flame_lbs_fwd_ex(const float* betas, const float* full_pose, int32_t pose_is_rotmat, const float* dirs, const float* jreg,
                                    const float* lbs_weights, float* coef, float* A, float* verts, float* joints, int32_t* dyn_rows,
                                    int32_t F, int32_t V, int32_t NB, int32_t K_pad, void* stream);

extern "C" int avi_flame_lbs_fwd(const float* betas, const float* full_pose, const float* dirs, const float* jreg,
                                 const float* lbs_weights, float* coef, float* A, float* verts, float* joints,
                                 int32_t* dyn_rows, int32_t F, int32_t V, int32_t NB, int32_t K_pad, void* stream) {
  return avi_flame_lbs_fwd_ex(betas, full_pose, 0, dirs, jreg, lbs_weights, coef, A, verts, joints, dyn_rows, F, V, NB, K_pad, stream);
}

extern "C" int avi_flame_lbs_fwd_ex(const float* betas, const float* full_pose, int32_t pose_is_rotmat, const float* dirs, const float* jreg,
                                    const float* lbs_weights, float* coef, float* A, float* verts, float* joints, int32_t* dyn_rows,
                                    int32_t F, int32_t V, int32_t NB, int32_t K_pad, void* stream) {
  AVI_REQUIRE(F > 0 && V > 0 && NB > 0 && K_pad >= NB + 37 && K_pad % FL_KC == 0, "avi_flame_lbs_fwd: bad shape F=%d V=%d NB=%d K_pad=%d",
              F, V, NB, K_pad);
  cudaStream_t st = (cudaStream_t)stream;
  flame_prologue_kernel<<<(F + 3) / 4, 128, 0, st>>>(betas, full_pose, jreg, coef, A, joints, dyn_rows, F, NB, K_pad, pose_is_rotmat ? 1 : 0);
  if (check_launch("flame_prologue")) return 1;
  dim3 grid((V + FL_VT - 1) / FL_VT, (F + FL_FT - 1) / FL_FT);
  AVI_REQUIRE(grid.y <= 65535, "avi_flame_lbs_fwd: too many frames in one call (%d); split the batch", F);
  flame_blend_skin_kernel<<<grid, 256, 0, st>>>(coef, dirs, A, lbs_weights, verts, F, V, K_pad);
  return check_launch("flame_blend_skin");
}

extern "C" int avi_flame_landmarks(const float* verts, const int64_t* faces, const int64_t* idx, const float* bary, float* out,
                                   int32_t F, int32_t V, int32_t L, int32_t per_frame, void* stream) {
  AVI_REQUIRE(F > 0 && V > 0 && L > 0, "avi_flame_landmarks: bad shape");
  flame_landmarks_kernel<<<(F * L + 127) / 128, 128, 0, (cudaStream_t)stream>>>(verts, faces, idx, bary, out, F, V, L, per_frame);
  return check_launch("flame_landmarks");
}

// ------------------------------------------------------------------------------------------------ compact result sink (opt-in)
// out[r, c] = fp16(verts[r, c] - template[c]): the displacement from the neutral face is ~1e-2 m, so fp16 keeps ~5e-6 m while the
// device->host copy (the bound of the end-to-end path: 1.9 GB of fp32 vertices per 64-clip step) halves. Outside the fp32 contract.
namespace avi {
__global__ void __launch_bounds__(256) pack_disp_f16_kernel(const float* __restrict__ verts, const float* __restrict__ tpl,
                                                            __half* __restrict__ out, int64_t rows, int C, int64_t ld) {
  const int64_t row = blockIdx.y;
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) * 2; c < C; c += gridDim.x * blockDim.x * 2) {
    const float a = verts[row * ld + c] - tpl[c];
    if (c + 1 < C) {
      const float b = verts[row * ld + c + 1] - tpl[c + 1];
      if (((row * (int64_t)C + c) & 1) == 0) {
        *reinterpret_cast<__half2*>(out + row * (int64_t)C + c) = __floats2half2_rn(a, b);
      } else {
        out[row * (int64_t)C + c] = __float2half_rn(a);
        out[row * (int64_t)C + c + 1] = __float2half_rn(b);
      }
    } else {
      out[row * (int64_t)C + c] = __float2half_rn(a);
    }
  }
}
}  // namespace avi

extern "C" int avi_pack_disp_f16(const float* verts, const float* tpl, void* out_f16, int64_t rows, int32_t C, int64_t ld, void* stream) {
  AVI_REQUIRE(rows > 0 && rows <= 65535 * 64LL && C > 0 && ld >= C, "avi_pack_disp_f16: bad shape");
  for (int64_t r0 = 0; r0 < rows; r0 += 65535) {   // grid.y limit
    const int64_t n = rows - r0 < 65535 ? rows - r0 : 65535;
    pack_disp_f16_kernel<<<dim3(8, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(verts + r0 * ld, tpl, reinterpret_cast<__half*>(out_f16) + r0 * C,
                                                                               n, C, ld);
  }
  return check_launch("pack_disp_f16");
}
