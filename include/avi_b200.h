/*
 * avi_b200.h - C ABI of libavi_b200.so: hand-written sm_100a kernels for the
 * audio -> wav2vec2 -> FaceFormer decoder -> FLAME vertices hot path of
 * sunyasheng/AVI-Talking.
 *
 * The reference has no FFI of its own (it is pure Python/PyTorch); the
 * seam is its nn.Module classes.  Every entry point below replaces the ATen /
 * cuBLAS / cuDNN work a reference method does, cited as file:line relative to
 * the upstream repository root.  The Python drop-in classes under
 * avi_talking_b200/ bind these with ctypes.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - plain pointers and sizes only, no torch / C++ types;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return 0 on success, non-zero on error; avi_last_error() gives the message
 *     (thread-local).  The Python side turns non-zero into RuntimeError;
 *   - nothing allocates, frees or synchronises; all calls are asynchronous on
 *     `stream` and re-entrant across streams (the tensor-core GEMM keeps an
 *     internal, mutex-protected cache of TMA descriptors only);
 *   - row-major everywhere; activations are time-major [clip, frame, channel].
 */
#ifndef AVI_B200_H_
#define AVI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVI_B200_VERSION 100 /* 0.1.0 */

int avi_version(void);
const char* avi_last_error(void);
/* number of kernel launches this process has issued through the library (for bench.py's gpu_launches) */
int64_t avi_launch_count(void);

/* ------------------------------------------------------------------ dense contraction ------------------------------------------------------------------
 * C[b, r, n] = epi( sum_k A[b, r, k] * W[n, k] + bias[n] ) (+ residual[b, r, n])
 * Replaces torch.nn.Linear / Conv1d-as-GEMM calls: HF Wav2Vec2 q/k/v/out_proj, feed_forward, feature_projection,
 * conv_layers[1..6] (models/lib/wav2vec.py:97,120,142 via transformers), audio_feature_map / v_merge2hidden /
 * vertice_map_r (models/faceformer_disentangle.py:437,473,776).
 *
 * A rows may overlap (Conv1d as a GEMM over time-major activations): the row for (b, r) starts at
 *   A + b*a_batch_stride + r*conv_stride*a_ld  and is conv_taps*a_ld... see AviGemmArgs below.
 */
enum { AVI_ACT_NONE = 0, AVI_ACT_GELU = 1 /* exact erf */, AVI_ACT_RELU = 2, AVI_ACT_QUICK_GELU = 3 /* x * sigmoid(1.702 x), CLIP */,
       AVI_ACT_SILU = 4 /* x * sigmoid(x): avi_act_fwd / avi_act_bwd only (the prior's time MLP) */ };
enum { AVI_DT_F32 = 0, AVI_DT_BF16 = 1, AVI_DT_TF32 = 2 /* fp32 storage rounded to TF32 (operand producers only) */ };

typedef struct AviGemmArgs {
  const void* A;      /* activations, dtype a_dtype; element (b, row, c) at A + b*a_batch_stride + row*a_ld + c              */
  const void* W;      /* weights [N, K] row-major (nn.Linear layout; conv weights packed [N, taps*C] tap-major), dtype a_dtype */
  const float* bias;  /* [N] or NULL                                                                                         */
  const float* residual; /* fp32 [b, r, n] at residual + b*res_batch_stride + r*res_ld + n, or NULL (added AFTER the activation) */
  void* C;            /* output, dtype c_dtype; element (b, r, n) at C + b*c_batch_stride + r*c_ld + n                       */
  void* C2;           /* optional second copy of the output in the other dtype (same strides), or NULL                      */
  int32_t batch;      /* b extent                                                                                            */
  int32_t rows;       /* r extent (output rows per batch entry)                                                              */
  int32_t N;
  int32_t K;          /* = conv_taps * C_in                                                                                   */
  int32_t conv_taps;  /* 1 for a plain GEMM                                                                                   */
  int32_t conv_stride;/* 1 for a plain GEMM; output row r reads input rows r*conv_stride .. +conv_taps-1                      */
  int64_t a_ld;       /* elements between consecutive INPUT rows (= C_in for a conv, = K for a plain GEMM unless padded)       */
  int64_t a_batch_stride;
  int64_t a_rows_alloc; /* input rows that are safely readable per batch entry (tensor-core path: TMA bound)                 */
  int64_t c_ld, c_batch_stride;
  int64_t res_ld, res_batch_stride;
  int32_t a_dtype, c_dtype, act;
  /* 2-D taps (conv_stride must be 1): tap j reads input row r + (j / conv_taps_x) * conv_row_pitch + (j % conv_taps_x), i.e. a
   * k x k convolution over image lines of conv_row_pitch pixels is ONE contraction with conv_taps = k*k. 0 = the 1-D taps above. */
  int32_t conv_taps_x, conv_row_pitch;
} AviGemmArgs;

/* fp32 CUDA-core path (exact mode, any shape). a_dtype must be AVI_DT_F32. */
int avi_gemm_f32(const AviGemmArgs* args, void* stream);
/* bf16 tcgen05/TMEM/TMA path. a_dtype must be AVI_DT_BF16; K % 64 == 0, a_ld % 8 == 0, 16-byte aligned bases.
 * Output rows whose pitch c_ld ends exactly at the 16-byte granule holding column N-1 (padded rows, e.g. N = 15069 on c_ld = 15072)
 * get zeros in those padding columns; nothing else outside [rows, N] is written. residual == C (fp32) updates C in place. */
int avi_gemm_bf16_tc(const AviGemmArgs* args, void* stream);
/* same kernel on fp32 operands read as TF32 (tcgen05.mma.kind::tf32: 10-bit significand, fp32 accumulate; half the MMA rate).
 * a_dtype must be AVI_DT_F32; K per tap % 32 == 0, a_ld % 4 == 0, 16-byte aligned bases. */
int avi_gemm_tf32_tc(const AviGemmArgs* args, void* stream);
/* 1 if the tensor-core path accepts these shapes (host-side check only) */
int avi_gemm_bf16_tc_supported(const AviGemmArgs* args);
/* Opt in (1) / out (0, default; env AVI_GEMM_MULTICAST sets the initial value) of the 4-CTA-cluster schedule of the two kernels
 * above for launches with more than one wave of tiles: two CTA pairs per cluster on consecutive m-tiles, W-tile quarters
 * TMA-multicast between them. Same results bit for bit; measured no faster on B200 (csrc/gemm_tc2.cu). Returns the old value. */
int avi_gemm_set_multicast(int on);
/* Tile scheduling of the persistent tensor-core kernels. Bit clear (default) = the static walk tile = cta + k * ctas of a 148-CTA grid;
 * bit set = DYNAMIC: the grid holds one cluster per tile and the resident CTAs / CTA pairs take the pending ones through cluster launch
 * control (clusterlaunchcontrol.try_cancel), so a launch that shares the GPU with another kernel finishes on the SMs it has instead of
 * waiting for its not-yet-resident CTAs. Results are bit-identical; on the configs[1] step it measured no faster
 * (profiles/r2/clc_dynamic_tiles_ab.txt), hence opt-in. Bits: 1 avi_gemm_bf16_tc / avi_gemm_tf32_tc, 2 avi_w2v_conv0_gn_gelu_tc.
 * Returns the previous mask; AVI_DYNAMIC_TILES=<mask> in the environment sets the initial one. */
int avi_set_dynamic_tiles(int32_t mask);
/* Programmatic dependent launch (griddepcontrol.wait + cudaLaunchAttributeProgrammaticStreamSerialization) of the tensor-core GEMM, the
 * wav2vec2 attention and the LayerNorm kernels: each may be scheduled while the previous kernel of its stream drains and runs its
 * prologue (barrier init, TMEM allocation) before waiting for that kernel's memory. 0 = plain stream-ordered launches. Results are
 * identical. Returns the previous setting; AVI_PDL=<0|1> in the environment sets the initial one. */
int avi_set_pdl(int32_t on);
/* fp32 -> bf16 (weights packing / activation staging), n elements */
int avi_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* zero-padded bf16 copy of a time-major activation: dst[b, t, :] = (front <= t < front+T) ? src[b, t-front, :] : 0, rows_out rows
 * per clip (stages the input of the positional conv for the tensor-core path). */
int avi_pad_cast_bf16(const float* src, void* dst, int32_t B, int32_t T, int32_t C, int32_t front, int32_t rows_out, void* stream);
/* split-bf16 GEMM operand: dst [rows, 3K] = [hi(x) | lo(x) | hi(x)] with lo = bf16(x - hi); against weights packed
 * [hi(W) | hi(W) | lo(W)] the bf16 tensor path reproduces the fp32 product to ~2^-16 relative (vertex head, bf16 mode). */
int avi_split_bf16x3(const float* src, void* dst, int64_t rows, int32_t K, void* stream);
/* fp32-ACCURATE contractions on the bf16 tensor path (the "fp32 <= 1e-5" mode of north_star without the CUDA-core GEMM): every fp32
 * value is split into bf16 terms x = s0 + s1 + s2, s0 = bf16(x), s1 = bf16(x - s0), s2 = bf16(x - s0 - s1), and dst row r is the
 * concatenation of `nterms` K-wide blocks, block t holding s_{p_t} with p_t = bits 2t..2t+1 of `pattern`. A = [s0|s1|s0] against
 * W = [t0|t0|t1] keeps s0 t0 + s1 t0 + s0 t1 (2^-16 relative); A = [s0|s0|s0|s1|s1|s2] against W = [t0|t1|t2|t0|t1|t0] keeps every
 * product down to 2^-24 (fp32 accuracy) at six bf16 MMAs per product. src [rows, K] fp32 contiguous (K % 4 == 0), dst [rows, nterms*K]
 * bf16. Replaces the fp32 torch.matmul / F.linear / F.conv1d arithmetic of the reference in that mode (same call sites as avi_gemm_f32). */
int avi_split_bf16_terms(const float* src, void* dst, int64_t rows, int32_t K, int32_t nterms, uint32_t pattern, void* stream);

/* ------------------------------------------------------------------ wav2vec2 pieces ------------------------------------------------------------------ */
/* Conv1d(1->512,k=10,s=5,no bias) + GroupNorm(512 groups) + GELU   (HF Wav2Vec2GroupNormConvLayer; models/lib/wav2vec.py:97).
 * audio [B, n_samples] fp32 -> out [B, L0, C] (time-major, dtype out_dtype, batch stride out_batch_stride elements).
 * stats: scratch of B*C*2 doubles (zeroed by the call). */
int avi_w2v_conv0_gn_gelu(const float* audio, const float* w /*[C,10]*/, const float* gn_w, const float* gn_b,
                          void* stats, void* out, int32_t out_dtype, int64_t out_batch_stride,
                          int32_t B, int32_t n_samples, int32_t C, float eps, void* stream);

/* Same layer on the tcgen05 path (bf16 output, C == 512): the 10-tap contraction is a split-bf16 MMA (fp32-accurate to ~2^-16), the
 * GroupNorm statistics come from 65 audio moments per clip (exact, fp64), GroupNorm + GELU are the MMA epilogue.
 *   avi_w2v_conv0_pack_tc: w fp32 [512,10] -> w_packed bf16 [512,64] (once per weight version)
 *   stats: the same scratch as above (B*C*2 doubles). */
int avi_w2v_conv0_pack_tc(const float* w, void* w_packed, void* stream);
int avi_w2v_conv0_gn_gelu_tc(const float* audio, const float* w, const void* w_packed, const float* gn_w, const float* gn_b, void* stats,
                             void* out, int64_t out_batch_stride, int32_t B, int32_t n_samples, int32_t C, float eps, void* stream);

/* align_corners linear resample over time + LayerNorm(C): models/lib/wav2vec.py:67-73,108 + HF feature_projection.layer_norm.
 * in [B, T_in, C] (in_dtype, batch stride in elements) -> out_f32 / out_bf16 [B*T_out, C] (either may be NULL). */
int avi_w2v_lerp_layernorm(const void* in, int32_t in_dtype, int64_t in_batch_stride, const float* ln_w, const float* ln_b,
                           float* out_f32, void* out_bf16, int32_t B, int32_t T_in, int32_t T_out, int32_t C, float eps,
                           void* stream);

/* y = LayerNorm(x (+ res)) over the last dim; x, res fp32 [rows, C] (res may be NULL); writes fp32 and/or bf16 copies.
 * (HF encoder layer_norm / final_layer_norm; nn.TransformerDecoderLayer norm1-3; C <= 8192: one warp per row up to 1024, one block above) */
int avi_layernorm(const float* x, const float* res, const float* w, const float* b, float* out_f32, void* out_bf16, int64_t rows, int32_t C,
                  float eps, void* stream);

/* h = LayerNorm(x + GELU(grouped_conv1d(x)[..., :-1])) : HF Wav2Vec2PositionalConvEmbedding + encoder.layer_norm.
 * x fp32 [B, T, C]; w_packed fp32 [groups][k][C/groups in][C/groups out] (weight-norm g*v/||v|| already applied, packed
 * once at weight-load time); out_f32 is required (doubles as the conv scratch), optional bf16 copy. */
int avi_w2v_posconv_ln(const float* x, const float* w_packed, const float* conv_bias, const float* ln_w, const float* ln_b,
                       float* out_f32, void* out_bf16, int32_t B, int32_t T, int32_t C, int32_t groups, int32_t k, float eps,
                       void* stream);

/* second half of the above when the grouped conv itself ran as tensor-core GEMMs: out = LayerNorm(x + GELU(pc)); pc may alias out_f32 */
int avi_w2v_posconv_merge_ln(const float* x, const float* pc, const float* ln_w, const float* ln_b, float* out_f32, void* out_bf16,
                             int64_t rows, int32_t C, float eps, void* stream);

/* The grouped positional conv itself on tcgen05 with the activation slab resident in shared memory (csrc/posconv_tc.cu):
 *   pc[b, t, co] = bias[co] + sum_{j<k} sum_{ci<48} xpad[b, t + j, 48*(co/48) + ci] * w[co, ci, j]          (fp32, [B*T, C])
 * xpad: bf16 [B, Tp, C], Tp >= T + k - 1, rows [0, k/2) and [k/2 + T, Tp) zero (avi_pad_cast_bf16 with front = k/2);
 * w_band: bf16 [groups/4][k][3][96][64], the non-zero band of the block-diagonal weight of each 4-group quad, channel blocks in
 * stream order c = 0, 2, 1:  w_band[q][j][i][n][kk] = w[192q + 48c + n, (64c + kk) % 48, j] if (64c + kk) / 48 == (48c + n) / 48 else 0
 * (packed once per weight version by the host, wav2vec.py). Built for k = 128, 48-channel groups, groups % 4 == 0 (wav2vec2-base).
 * Replaces HF Wav2Vec2PositionalConvEmbedding.conv (reached from models/lib/wav2vec.py:142). pc is zeroed by the call. */
int avi_w2v_posconv_tc(const void* xpad, const void* w_band, const float* bias, float* pc, int32_t B, int32_t T, int32_t Tp, int32_t C,
                       int32_t groups, int32_t k, void* stream);

/* Faceformer.convert_coeff2verts prologue (models/faceformer_disentangle.py:426-430): exp_out [F, n_exp] = coeff[:, :n_exp] * std + mean
 * (coeff [F, n_coeff], mean / std [>= n_exp]); pose [F, n_pose]: columns 0..2 (global rotation) are zeroed IN PLACE, as upstream. */
int avi_ff_denorm_coeff(const float* coeff, const float* mean, const float* stdv, float* pose, float* exp_out, int32_t F, int32_t n_coeff,
                        int32_t n_exp, int32_t n_pose, void* stream);
/* Conditioning columns of the decoder input (models/faceformer_disentangle.py:808: cat[eye(6), emo(30), audio(fd)]):
 * out [B*T, ld] columns 0..5 = eye (eye_per_row == 0: ONE row of 6, the learnable embedding; else [B*T, 6]), columns 6..35 = emo
 * (clip b, frame t at emo + b * emo_clip_stride + t * 30). Columns >= 36 are not touched (the audio_feature_map GEMM writes them). */
int avi_ff_fill_cond(const float* eye, int32_t eye_per_row, const float* emo, int64_t emo_clip_stride, float* out, int32_t B, int32_t T,
                     int32_t ld, void* stream);

/* softmax(q k^T * scale) v per (clip, head); qkv [B, T, 3*H*D] packed (q | k | v), dtype qkv_dtype; out [B, T, H*D] same dtype.
 * (HF Wav2Vec2Attention / eager_attention_forward, no mask) */
int avi_mha_fwd(const void* qkv, void* out, int32_t dtype, int32_t B, int32_t T, int32_t H, int32_t D, float scale, void* stream);

/* same contract on the tcgen05 path (bf16, D == 64, T <= 256): QK^T and PV on tensor cores with S/O in TMEM, softmax from TMEM */
int avi_mha_fwd_tc(const void* qkv, void* out, int32_t B, int32_t T, int32_t H, int32_t D, float scale, void* stream);
int avi_mha_fwd_tc_supported(int32_t dtype, int32_t T, int32_t D);

/* ------------------------------------------------------------------ FaceFormer decoder (Path A) ------------------------------------------------------------------ */
typedef struct AviDecoderWeights { /* all fp32 device pointers; fd = feature_dim, 4 heads, dff = 2*fd */
  const float *sa_in_w, *sa_in_b;   /* [3fd, fd], [3fd]  self_attn.in_proj                                   */
  const float *sa_out_w, *sa_out_b; /* [fd, fd], [fd]                                                        */
  const float *ff1_w, *ff1_b;       /* [dff, fd]                                                             */
  const float *ff2_w, *ff2_b;       /* [fd, dff]                                                             */
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *ln3_w, *ln3_b;
  const float *fb_w, *fb_b;         /* [fd, fd], [fd]: composed feedback map vertice_map o vertice_map_r (AR only) */
  /* every *_w above is stored TRANSPOSED, [in, out] */
  const float *pe;                  /* [period, fd] periodic positional encoding table                       */
} AviDecoderWeights;

/* Autoregressive branch of Faceformer.forward_ff (models/faceformer_disentangle.py:461-476), KV-cached, one CTA per clip.
 * All matrices in AviDecoderWeights are TRANSPOSED nn.Linear weights ([in, out], packed once at weight-load time).
 * cross [B, T, fd]: the (degenerate) cross-attention term out_proj(v_proj(memory_t)) precomputed by GEMMs;
 * style [B, fd]; hidden_out [B, T, fd] = decoder output rows (input to vertice_map_r).
 * kv_scratch fp32 [B, 2, T, fd+1]: only used when the K/V cache does not fit in shared memory (may be NULL otherwise). */
int avi_ff_decoder_ar(const AviDecoderWeights* w, const float* cross, const float* style, float* hidden_out, float* kv_scratch,
                      int32_t B, int32_t T, int32_t fd, int32_t period, void* stream);

/* Biased causal self-attention over whole sequences (teacher-forced branch :445-458; mask = init_biased_mask :56-77 built on
 * the fly, 4 heads): qkv fp32 [B, T, 3*fd] (q | k | v) -> out [B, T, fd]. */
int avi_ff_biased_attn(const float* qkv, float* out, int32_t B, int32_t T, int32_t fd, int32_t period, void* stream);

/* ------------------------------------------------------------------ FLAME ------------------------------------------------------------------ */
/* One-time packing of the static FLAME buffers (third_party/inferno/inferno/models/DecaFLAME.py:60-74):
 *   dirs  [NB+36+1 (padded to K_pad), V*3] : rows 0..NB-1 = shapedirs^T, NB..NB+35 = posedirs, NB+36 = v_template
 *   jreg  [15, NB+1]                       : J_regressor applied to shapedirs (cols 0..NB-1) and to v_template (col NB)  */
int avi_flame_pack(const float* shapedirs /*[V,3,NB]*/, const float* posedirs /*[36,V*3]*/, const float* v_template /*[V,3]*/,
                   const float* J_regressor /*[5,V]*/, float* dirs, float* jreg, int32_t V, int32_t NB, int32_t K_pad, void* stream);

/* lbs() with pose2rot=True (third_party/inferno/inferno/utils/lbs.py:142-234) for F frames, 5 joints, parents [-1,0,1,1,1].
 * betas [F, NB]; full_pose [F, 15]; lbs_weights [V, 5]; coef scratch fp32 [F, K_pad]; A scratch fp32 [F, 5, 12];
 * verts out [F, V, 3]; joints out [F, 5, 3] or NULL; dyn_rows out int32 [F] or NULL = the row of the dynamic contour
 * landmark table each frame selects (DecaFLAME.py:110-149, neck_kin_chain = [1, 0]). */
int avi_flame_lbs_fwd(const float* betas, const float* full_pose, const float* dirs, const float* jreg, const float* lbs_weights,
                      float* coef, float* A, float* verts, float* joints, int32_t* dyn_rows, int32_t F, int32_t V, int32_t NB,
                      int32_t K_pad, void* stream);

/* first half of avi_flame_lbs_fwd only (coefficient rows, joints, kinematic chain, landmark rows), for the tensor-core blend below */
int avi_flame_prologue(const float* betas, const float* full_pose, const float* jreg, float* coef, float* A, float* joints,
                       int32_t* dyn_rows, int32_t F, int32_t NB, int32_t K_pad, void* stream);
/* the same two entry points for lbs(pose2rot=False) (lbs.py:205-209): pose_is_rotmat != 0 -> full_pose is [F, 5, 3, 3] rotation
 * matrices (pose feature = R[1:] - I, the matrices feed the kinematic chain directly); 0 -> [F, 15] axis-angle as above */
int avi_flame_prologue_ex(const float* betas, const float* full_pose, int32_t pose_is_rotmat, const float* jreg, float* coef, float* A,
                          float* joints, int32_t* dyn_rows, int32_t F, int32_t NB, int32_t K_pad, void* stream);
int avi_flame_lbs_fwd_ex(const float* betas, const float* full_pose, int32_t pose_is_rotmat, const float* dirs, const float* jreg,
                         const float* lbs_weights, float* coef, float* A, float* verts, float* joints, int32_t* dyn_rows, int32_t F,
                         int32_t V, int32_t NB, int32_t K_pad, void* stream);

/* tcgen05 blend + fused skinning (fp16 operands, fp32 accumulate/template/skinning; NB + 36 <= 192).
 *   avi_flame_pack_tc : dirs16 fp16 [3][V_pad][192] from the packed fp32 dirs of avi_flame_pack (V_pad multiple of 128)
 *   avi_flame_blend_skin_tc : consumes coef [F, K_pad32] and A [F,5,12] written by avi_flame_prologue; coef16 = scratch fp16 [F,192] */
int avi_flame_tc_supported(int32_t NB);
int avi_flame_pack_tc(const float* dirs32, void* dirs16, int32_t V, int32_t NB, int32_t V_pad, void* stream);
int avi_flame_blend_skin_tc(const float* coef32, const float* A, const void* dirs16, const float* lbs_weights,
                            const float* v_template, void* coef16, float* verts, int32_t F, int32_t V, int32_t NB,
                            int32_t K_pad32, int32_t V_pad, void* stream);

/* Grouped form with the SHAPE blendshapes hoisted out of the per-frame contraction (FlamePreprocessor / BertPriorDecoder call FLAME
 * with one shape per clip, Preprocessors.py:136-150): frames = G groups of frames_per_group; templates = one shaped template
 * [V,3] per group (template_stride floats apart, 0 = one global template); the tensor-core contraction covers coefficient columns
 * [coef_col0, coef_col0 + n_dirs) of coef32 (expression | pose feature) against direction rows [row0, row0 + n_dirs) of dirs32. */
int avi_flame_pack_tc_rows(const float* dirs32, void* dirs16, int32_t V, int32_t row0, int32_t n_dirs, int32_t V_pad, void* stream);
int avi_flame_blend_skin_tc_grouped(const float* coef32, const float* A, const void* dirs16, const float* lbs_weights,
                                    const float* templates, int64_t template_stride, void* coef16, float* verts,
                                    int64_t verts_frame_stride /* floats between output frames, >= V*3 */, int32_t F, int32_t V,
                                    int32_t n_dirs, int32_t coef_col0, int32_t K_pad32, int32_t V_pad, int32_t frames_per_group,
                                    void* stream);

/* cap (1..148, default 148) on the CTAs of the tensor-core FLAME kernel, for running it beside a kernel that occupies part of the GPU */
int avi_flame_set_max_ctas(int32_t n);

/* barycentric landmark gather (lbs.py:103-139): out[f, l, :] = sum_c bary[f|0, l, c] * verts[f, faces[idx[f|0, l], c], :].
 * per_frame = 1 when idx/bary carry a frame dimension (dynamic contour landmarks). */
int avi_flame_landmarks(const float* verts, const int64_t* faces, const int64_t* idx, const float* bary, float* out,
                        int32_t F, int32_t V, int32_t L, int32_t per_frame, void* stream);

/* Opt-in compact result sink (no reference counterpart; evaluation_functions.py:624-638 writes fp32 pickles): out_f16[r, c] =
 * fp16(verts[r*ld + c] - template[c]) for a dense [rows, C] fp16 buffer - the displacement from the neutral face keeps ~5e-6 m in
 * fp16 and the device->host copy halves. Outside the fp32 contract: the default sink and every headline number ship fp32 vertices. */
int avi_pack_disp_f16(const float* verts, const float* tpl, void* out_f16, int64_t rows, int32_t C, int64_t ld, void* stream);

/* ------------------------------------------------------------------ diffusion prior (text embedding -> style embedding) ------------------------------------------------------------------ */
/* Weights of VersatileDiffusionPriorNetwork as built at train_diffusion_prior.py:963-991 (dim 128, depth <= 8, 8 heads x 64, one
 * shared K/V head, SwiGLU inner 512), packed once at weight-load time. All fp32 device pointers.
 *   layers: [depth][avi_prior_layer_floats()] = per layer, in this order, every matrix TRANSPOSED to [in][out]:
 *     attn norm.g [128] | null_k [64] | null_v [64] | to_q^T [128][512] | to_kv^T [128][128] | to_out.0^T [512][128] | to_out.1.g [128]
 *     | ff 0.g [128] | ff 1^T [128][1024] | ff 5^T [512][128]
 *   rel_bias: RelPosBias(n = 3, n + 1 = 4) -> [heads][3][4]; rotary: cos/sin of position * freq, [3][16][2]  */
typedef struct AviPriorNet {
  const float* layers;
  const float* learned_query;   /* [128] (learned_query_mode = "pos_emb")                 */
  const float* rel_bias;
  const float* rotary;
  const float* norm_g;          /* causal_transformer.norm.g (stable LayerNorm)           */
  const float* project_out_t;   /* causal_transformer.project_out.weight^T [128][128]     */
  int32_t dim, depth, heads, dim_head, ff_inner;
} AviPriorNet;
int avi_prior_layer_floats(void);

/* time-token embeddings for all steps: SinusoidalPosEmb(128) -> MLP(128->256->256->128, SiLU) (models/diffusion_prior.py:186-189);
 * times [steps] fp32 (the timestep value of each step), weights transposed [in][out]; temb out [steps][128] */
int avi_prior_time_embed(const float* times, const float* w0t, const float* b0, const float* w1t, const float* b1, const float* w2t,
                         const float* b2, float* temb, int32_t steps, void* stream);

/* The whole sampling loop of InstructDiffusionPrior.p_sample_loop (models/diffusion_prior.py:329-367 DDPM branch; dalle2_pytorch
 * DiffusionPrior.p_sample_loop_ddim for fewer steps) in ONE launch, cond_scale == 1:
 *   text_embed [B][128], x_init [B][128] (initial noise), noise [steps][B][128] (the draw of every step, caller-supplied so the
 *   reference's generator stream can be reproduced), sched [steps][6] = {mode, p0..p4}:
 *     mode 0 (DDPM)  x = (p0*x0 + p1*x) + p2*noise         p0,p1 = posterior_mean_coef1/2[t], p2 = [t>0]*exp(0.5*logvar[t])
 *     mode 1 (DDIM)  e = (p0*x - x0)/p1 ; x = x0*p2 + p3*noise + p4*e
 *     mode 2         x = x0                               (last DDIM pair)
 *   out [B][128] = final x * out_scale (1 / image_embed_scale). samples_per_cta: 1, 2, 4 or 0 = choose. */
int avi_prior_sample(const AviPriorNet* net, const float* temb, const float* sched, const float* text_embed, const float* x_init,
                     const float* noise, float* out, int32_t B, int32_t steps, float out_scale, int32_t samples_per_cta, void* stream);
/* the same loop with classifier-free guidance (forward_with_cond_scale :209-221, cond_scale != 1): the null pass replaces BOTH the
 * text token and the noisy image token by the null embeddings, so its output depends on the timestep only: null_pred [steps][128]
 * is that output per step (one avi_prior_sample call of B = 1, steps = 1, mode 2 each, with text = null_brain_embeds and
 * x_init = null_image_embed), and x0 = null + (x0 - null) * cond_scale before the DDPM / DDIM update. null_pred == NULL: as above. */
int avi_prior_sample_cfg(const AviPriorNet* net, const float* temb, const float* sched, const float* text_embed, const float* x_init,
                         const float* noise, const float* null_pred, float cond_scale, float* out, int32_t B, int32_t steps,
                         float out_scale, int32_t samples_per_cta, void* stream);

/* y = GELU(LayerNorm(x)) (+ res), rows of C <= 4096: BrainNetwork blocks and projector (models/diffusion_prior.py:63-93,104-110) */
int avi_ln_gelu_res(const float* x, const float* w, const float* b, const float* res, float* out_f32, void* out_bf16, int64_t rows,
                    int32_t C, float eps, void* stream);

/* ------------------------------------------------------------------ EMOTE talking-head decoder (Path B, third_party/inferno) ------------------------------------------------------------------ */
/* per-clip zero-mean / unit-variance normalisation of the waveform (Wav2Vec2FeatureExtractor, called at
 * inferno/models/temporal/AudioEncoders.py:170-178): y = (x - mean) / sqrt(var + eps), x, y fp32 [B, n] */
int avi_audio_znorm(const float* x, float* y, int32_t B, int64_t n, float eps, void* stream);

/* softmax(q k^T * scale - slope_h * |i - j|) v for small heads (D = 16 or 32): the nn.TransformerEncoderLayer self-attention of
 * BertPriorDecoder (slopes = NULL, FaceFormerDecoder.py:996-1002) and of the L2L decoder (slopes [H] = ALiBi slopes, bias of
 * init_alibi_biased_mask_future, TransformerMasking.py:80-98). qkv fp32 [B, T, 3*H*D]; out [B, T, H*D] fp32 and/or bf16. */
int avi_mha_small_fwd(const float* qkv, float* out_f32, void* out_bf16, int32_t B, int32_t T, int32_t H, int32_t D, float scale,
                      const float* slopes, void* stream);

/* row staging for conv-mode GEMMs: dst [B, Lp, C] (dst_dtype) from src fp32 [B, L, C];
 * mode 0 zero padding (front rows of zeros first), 1 replicate padding, 2 zero insertion dst[front + 2t] = src[t]
 * (ConvTranspose1d(k=5, s=2, p=2, op=1) == zero insertion + 5-tap correlation with the flipped kernel; L2lMotionPrior.py:368-389) */
int avi_stage_rows(const float* src, void* dst, int32_t dst_dtype, int32_t B, int32_t L, int32_t Lp, int32_t C, int32_t front,
                   int32_t mode, void* stream);

/* y[b, rep*t + u, c] = bn_scale[c] * LeakyReLU(x[b, t, c]) + bn_shift[c]: expander activation + eval BatchNorm1d + repeat_interleave
 * (L2lMotionPrior.py:376-389,468-470); x fp32 [B, L, C] -> y fp32 [B, rep*L, C] */
int avi_lrelu_bn_repeat(const float* x, const float* bn_scale, const float* bn_shift, float* y, int32_t B, int32_t L, int32_t C,
                        int32_t repeat, float slope, void* stream);

/* out[b, t, :] = (a[b, t, :] - neutral[b, :]) + tpl[b, :]  (vertex offsets from the neutral shape re-attached to the template,
 * FaceFormerDecoder.py:1173-1175,690-694); a / out rows are row_stride floats apart (>= C); out may alias a */
int avi_sub_add_rows(const float* a, const float* neutral, const float* tpl, float* out, int32_t B, int32_t T, int32_t C,
                     int64_t row_stride, void* stream);

/* ------------------------------------------------------------------ CLIP text tower (models/diffusion_prior.py:30-55 -> HF CLIPTextModel)
 * out[b*T + t, :] = tok_emb[ids[b, t], :] + pos_emb[t, :]   (CLIPTextEmbeddings) */
int avi_embed_tokens(const int64_t* ids, const float* tok_emb, const float* pos_emb, float* out, int32_t B, int32_t T, int32_t C,
                     int32_t vocab, void* stream);
/* out[b, :] = mean_t x[b*T + t, :]   (the 77-token mean pooling of train_diffusion_prior.py:439,711) */
int avi_token_mean(const float* x, float* out, int32_t B, int32_t T, int32_t C, void* stream);

/* ------------------------------------------------------------------ FanEncoder image branch (SURVEY 8f row 1;
 * third_party/pd_fgc_inference/lib/models/networks/FAN_feature_extractor.py:13-163, encoder.py:89-126). NHWC fp32 rows [N*H*W, C].
 * cols[(n,oy,ox), (ky,kx,c)] = act(x[n, oy*stride-pad+ky, ox*stride-pad+kx, c]) (zero outside the image), act = relu(x*scale+shift)
 * when scale != NULL (ConvBlock's pre-activation BatchNorm + ReLU, :38-48); rows of x are x_ld floats apart; Kpad >= k*k*C;
 * cols_dtype AVI_DT_TF32 writes fp32 rounded to nearest TF32 (operand of avi_gemm_tf32_tc, whose MMA truncates);
 * image lines of x are Wp_in pixels apart (>= W); Wo_extra surplus output columns per line are produced too (padded-width layout) */
int avi_im2col_affine(const float* x, int64_t x_ld, void* cols, int32_t cols_dtype, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k,
                      int32_t stride, int32_t pad, int32_t Kpad, const float* scale, const float* shift, int32_t Wp_in, int32_t Wo_extra,
                      void* stream);
/* operand of the implicit 3x3 convolution: a[n, y+1, x+1, :] = act(x[n, y, x, :]) with a zero one-pixel border, both with line pitch
 * W + 2 pixels (a has (H+2)*(W+2) rows per image; allocate 2 spare rows at the very end). Output pixel r = y*(W+2) + x' and tap
 * (ky, kx) then read operand row r + ky*(W+2) + kx: ONE conv-mode GEMM with 2-D taps (conv_taps = 9, conv_taps_x = 3, conv_row_pitch = W+2).
 * a_dtype as cols_dtype */
int avi_pad_act(const float* x, int64_t x_ld, void* a, int32_t a_dtype, int32_t N, int32_t H, int32_t W, int32_t C, const float* scale,
                const float* shift, void* stream);
/* F.max_pool2d(x, 2, stride=2) (:86, :142); line pitches in pixels */
int avi_maxpool2x2(const float* x, float* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t Wp_in, int32_t Wp_out, void* stream);
/* out = up1 + bilinear upsample of low to (Ho, Wo), align_corners=False (HourGlass._forward :97-101); line pitches in pixels */
int avi_upsample_bilinear_add(const float* low, const float* up1, float* out, int32_t N, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo,
                              int32_t C, int32_t Wp_in, int32_t Wp_out, void* stream);
/* in place x = act(x * scale[c] + shift[c]) (eval BatchNorm after a GEMM; scale == shift == NULL: activation only) */
int avi_affine_act(float* x, const float* scale, const float* shift, int64_t rows, int32_t C, int32_t relu, void* stream);

/* ------------------------------------------------------------------ training step (BASELINE configs[4]) ------------------------------------------------------------------
 * Backward / optimizer pieces of the teacher-forced faceformer_vert step (models/faceformer_vert.py:360-482; feature extractor
 * frozen :154). Dense backward contractions use avi_gemm_bf16_tc on operands laid out by avi_transpose_cast_bf16. */
/* dst (dst_dtype) [C, R_pad] = transpose(src fp32 [R, C], row stride src_ld), zero padded to R_pad rows of the source */
int avi_transpose_cast(const float* src, void* dst, int32_t dst_dtype, int32_t R, int32_t C, int64_t src_ld, int32_t R_pad, void* stream);
/* dst (dst_dtype) [R_pad, C_pad] = src fp32 [R, C] zero padded (contraction dims that are not multiples of 64: 15069 -> 15104) */
int avi_cast_pad2d(const float* src, void* dst, int32_t dst_dtype, int32_t R, int32_t C, int64_t src_ld, int32_t R_pad, int32_t C_pad,
                   void* stream);
/* teacher-forcing input (faceformer_vert.py:443-444): out[b*T+t, :C] = gt[b, t-1] - template (zero row for t = 0), zero pad columns.
 * gt rows of all clips are gt_ld floats apart */
int avi_tf_input_rows(const float* gt, int64_t gt_ld, const float* tpl, float* out, int32_t B, int32_t T, int32_t C, int32_t C_pad,
                      void* stream);
/* x[b,t,:] += style[b*style_stride + :] + pe[t mod period, :]  (faceformer_vert.py:445-447) */
int avi_ff_add_style_pe(float* x, const float* style, int64_t style_stride, const float* pe, int32_t B, int32_t T, int32_t fd,
                        int32_t period, void* stream);
/* out[n] (+)= sum_r x[r, n]   (bias gradients) */
int avi_colsum(const float* x, float* out, int32_t R, int32_t N, int64_t ld, int32_t accumulate, void* stream);
/* act = AVI_ACT_GELU | AVI_ACT_RELU | AVI_ACT_SILU: out = act(pre) ; dpre = dout * act'(pre) */
int avi_act_fwd(const float* pre, float* out_f32, void* out_bf16, int64_t n, int32_t act, void* stream);
int avi_act_bwd(const float* pre, const float* dout, float* dpre, int64_t n, int32_t act, void* stream);
/* LayerNorm backward over the last dim (C <= 8192): dx (may be NULL), dw += , db += (may be NULL; accumulate with atomics, zero them first) */
int avi_layernorm_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int64_t rows, int32_t C, float eps,
                      void* stream);
/* attention for training (T <= 128): forward keeps P [B,H,T,T] (P may be NULL when no backward follows); bias_mode 0 none, 1 FaceFormer biased causal mask
 * (init_biased_mask, faceformer_vert.py via :56-77 of faceformer_disentangle.py), 2 plain causal mask (CLIP text). qkv fp32 [B,T,3*H*D] */
int avi_attn_train_fwd(const float* qkv, float* out, float* P, int32_t B, int32_t T, int32_t H, int32_t D, float scale, int32_t bias_mode,
                       int32_t period, void* stream);
int avi_attn_train_bwd(const float* qkv, const float* P, const float* dout, float* dqkv, float* dS_scratch /* [B,H,T,T] */, int32_t B,
                       int32_t T, int32_t H, int32_t D, float scale, void* stream);
/* TRAIN mode of the same attention (dropout on the probabilities: nn.MultiheadAttention(dropout=0.1) inside nn.TransformerDecoderLayer,
 * faceformer_vert.py transformer_decoder; HF Wav2Vec2Attention attention_dropout): pmask [B,H,T,T] fp32 is the draw, Bernoulli(1-p)/(1-p),
 * an INPUT (parity stays testable). P keeps the softmax itself; out = (P * pmask) V. pmask == NULL: the calls above. */
int avi_attn_train_fwd_drop(const float* qkv, float* out, float* P, const float* pmask, int32_t B, int32_t T, int32_t H, int32_t D,
                            float scale, int32_t bias_mode, int32_t period, void* stream);
int avi_attn_train_bwd_drop(const float* qkv, const float* P, const float* pmask, const float* dout, float* dqkv, float* dS_scratch,
                            int32_t B, int32_t T, int32_t H, int32_t D, float scale, void* stream);
/* nn.Dropout with the draw as an input: y = x * mask (+ residual when not NULL); its own backward (dx = dy * mask). n % 4 == 0.
 * Replaces the dropout sites of Wav2Vec2EncoderLayer / Wav2Vec2FeedForward / PeriodicPositionalEncoding / nn.TransformerDecoderLayer
 * (dropout1-3, the feed-forward dropout) in the faceformer_vert training step (models/faceformer_vert.py:360-371,437-454). */
int avi_mask_mul_add(const float* x, const float* mask, const float* residual, float* y, int64_t n, void* stream);
/* SpecAugment along time (models/lib/wav2vec.py:120-131): rows [rows, C] whose row_mask byte is set are replaced by masked_spec_embed
 * (in place); backward: g_embed = sum of the masked rows of dx, which are then zeroed (fixed summation order). */
int avi_spec_augment_fwd(float* x, const uint8_t* row_mask, const float* embed, int64_t rows, int32_t C, void* stream);
int avi_spec_augment_bwd(float* dx, const uint8_t* row_mask, float* g_embed, int64_t rows, int32_t C, void* stream);
/* The draws themselves, on the device (csrc/train_draw.cu), so that a TRAIN-mode step replays from one CUDA graph with fresh draws:
 * counter-based Philox4x32-10 keyed by state = uint32[4] {seed_lo, seed_hi, step, 0} in device memory.
 * avi_dropout_masks: element i of out[n] (n % 4 == 0) = word i % 4 of philox(ctr = (i/4 lo, i/4 hi, step, stream_id), key = seed);
 *   u = (word >> 8) * 2^-24; out = u >= p ? 1/(1-p) : 0 - every nn.Dropout site of the step in one flat buffer, one launch
 *   (replaces torch.nn.functional.dropout at each site of HF Wav2Vec2EncoderLayer / PeriodicPositionalEncoding / nn.TransformerDecoderLayer).
 * avi_layerdrop_spec_draw: LayerDrop (Wav2Vec2Encoder.forward: skip layer l when u_l < layerdrop; u_l = word 0 of
 *   philox((l, 0, step, 0x4C440000))) written as keep_flags[l] in {0, 1} and, when blend != NULL, blend[0][l][0..rows) = keep,
 *   blend[1][l][0..rows) = 1 - keep; SpecAugment along time (models/lib/wav2vec.py:16-63,120-131, min_masks = 2): per clip
 *   n = max(min_spans, floor(span_rate + u)) spans of span_len rows starting at (word * (T - span_len + 1)) >> 32, spec[B*T] bytes.
 * avi_draw_bump_step: state[2] += 1 (stream-ordered after the draws). */
int avi_dropout_masks(float* out, int64_t n, float p, const uint32_t* state, uint32_t stream_id, void* stream);
int avi_layerdrop_spec_draw(float* blend, float* keep_flags, int32_t n_layers, int64_t rows, float layerdrop, uint8_t* spec, int32_t B,
                            int32_t T, int32_t span_len, float span_rate, int32_t min_spans, const uint32_t* state, void* stream);
int avi_draw_bump_step(uint32_t* state, void* stream);
/* positional conv: dW [C, C/groups, k] from x [B,T,C] and d(pre-activation) [B,T,C]; weight-norm chain rule (dim = 2) */
int avi_posconv_dw(const float* x, const float* dpc, float* dw, int32_t B, int32_t T, int32_t C, int32_t groups, int32_t k, void* stream);
/* transposed unfold of the positional conv input for a block of CW channels starting at c0:
 * out[(j*CW + ci), b*Tq + t] = xpad[b, t + j, c0 + ci] (zero for T <= t < Tq); dW of the block is then one GEMM over (clip, frame) */
int avi_posconv_unfold_t(const float* x, void* out, int32_t out_dtype, int32_t B, int32_t T, int32_t Tq, int32_t C, int32_t c0, int32_t CW,
                         int32_t k, void* stream);
int avi_weightnorm_bwd(const float* v, const float* g, const float* dw, float* dv, float* dg, int32_t n_rows /* C * C/groups */, int32_t k,
                       void* stream);
/* loss = mean((out - gt)^2) * loss_scale (fp64 scalar on the device) ; dout = d loss / d out (dense [rows, C]) */
int avi_mse_loss_grad(const float* out, const float* gt, float* dout, double* loss, int64_t rows, int32_t C, int64_t out_ld, int64_t gt_ld,
                      float loss_scale, void* stream);
/* torch.optim.Adam step on a flat fp32 buffer; grad_scale multiplies the gradient first (1 / world_size after a sum all-reduce);
 * p_bf16 (may be NULL) receives the bf16 copy of the updated parameters (the GEMM operands of the next step) */
int avi_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2, float eps,
                  int32_t step, float grad_scale, void* stream);
int avi_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream);
/* align_corners linear resample over time only (models/lib/wav2vec.py:67-73), fp32 output [B*T_out, C] */
int avi_w2v_lerp(const void* in, int32_t in_dtype, int64_t in_batch_stride, float* out, int32_t B, int32_t T_in, int32_t T_out, int32_t C,
                 void* stream);

/* ---------------------------------------------------------------------------------------------------------------------------------
 * Diffusion-prior TRAINING step: train_diffusion_prior.py:422-499, models/diffusion_prior.py:369-456 (p_losses /
 * forward) over dalle2_pytorch's Attention / FeedForward / LayerNorm. Dense contractions go through avi_gemm_*; these are the rest.
 * All fp32, dim 128 / 3 tokens / 8 heads x 64 (the configuration train_diffusion_prior.py:966-991 builds).
 * --------------------------------------------------------------------------------------------------------------------------------- */
/* VersatileDiffusionPriorNetwork.forward :258-306 fused with NoiseScheduler.q_sample (p_losses :372):
 *   tokens[b] = [ keep_brain[b] ? brain[b] : null_brain ; temb[b] ; (keep_image[b] ? sqrt_ac[t_b] x0[b] + sqrt_1mac[t_b] noise[b] : null_image)
 *                 + learned_query ]      (keep_* are 0/1 floats: the classifier-free-guidance masks of :258-281 as INPUTS)
 * x_noisy (may be NULL) receives the q_sample result [B, dim]. */
int avi_prior_tokens_fwd(const float* brain, const float* null_brain, const float* keep_brain, const float* x0, const float* noise,
                         const float* sqrt_ac, const float* sqrt_1mac, const int32_t* times, const float* null_image,
                         const float* keep_image, const float* learned_query, const float* temb, float* tokens, float* x_noisy, int32_t B,
                         int32_t dim, void* stream);
/* dtokens [B,3,dim] -> dbrain [B,dim], dtemb [B,dim]; dnull_brain / dnull_image / dlearned_query [dim] are ACCUMULATED (+=) */
int avi_prior_tokens_bwd(const float* dtokens, const float* keep_brain, const float* keep_image, float* dbrain, float* dtemb,
                         float* dnull_brain, float* dnull_image, float* dlearned_query, int32_t B, int32_t dim, void* stream);
/* dalle2_pytorch.Attention core (cosine_sim, scale 16, causal=False) for n_tokens = 3 queries against null + 3 keys:
 * q [3B, heads*dim_head] and kv [3B, 2*dim_head] are the raw to_q / to_kv projections, null_kv [2, dim_head], rotary [3, 16, 2]
 * (cos, sin of position * freq), rel_bias [heads, 3, 4] -> out [3B, heads*dim_head] (before to_out), P [B, heads, 3, 4] saved. */
int avi_prior_attn_fwd(const float* q, const float* kv, const float* null_kv, const float* rotary, const float* rel_bias, float* out,
                       float* P, int32_t B, int32_t n_tokens, int32_t heads, int32_t dim_head, void* stream);
/* -> dq, dkv (shapes of q, kv); per-sample partials dnull [B, 2*dim_head] and dS [B, heads*3*4] (fold with avi_colsum: deterministic) */
int avi_prior_attn_bwd(const float* q, const float* kv, const float* null_kv, const float* rotary, const float* P, const float* dout,
                       float* dq, float* dkv, float* dnull_per_sample, float* dS_per_sample, int32_t B, int32_t n_tokens, int32_t heads,
                       int32_t dim_head, void* stream);
/* SwiGLU of dalle2_pytorch.FeedForward: h [rows, 2*inner] = (a | gate) -> y = a * silu(gate) ; dh from (h, dy) */
int avi_swiglu_fwd(const float* h, float* y, int64_t rows, int32_t inner, void* stream);
int avi_swiglu_bwd(const float* h, const float* dy, float* dh, int64_t rows, int32_t inner, void* stream);
/* out = x / stat per row; mode 0: stat = row maximum (LayerNorm(stable=True), maximum detached), mode 1: stat = max(l2 norm, 1e-12)
 * (F.normalize). Backward: mode 0 dx = dy / stat ; mode 1 dx = (dy - y (y . dy)) / stat with y the forward output. */
int avi_rows_stat_div(const float* x, float* out, float* stat, int64_t rows, int32_t C, int32_t mode, void* stream);
int avi_rows_stat_div_bwd(const float* y, const float* dy, const float* stat, float* dx, int64_t rows, int32_t C, int32_t mode, void* stream);
int avi_mul_f32(const float* a, const float* b, float* y, int64_t n, void* stream);   /* dropout with the (pre-scaled) mask as an input */
int avi_scale_f32(const float* a, float alpha, float* y, int64_t n, void* stream);     /* y = alpha * a (image_embed * image_embed_scale, :453) */
/* soft_clip_loss (train_diffusion_prior.py:125-133) from the raw products pt = preds targs^T and tt = targs targs^T [B,B]:
 * loss (fp64 device scalar) and dsim = d loss / d pt [B,B]; lse_scratch holds 3*B floats. */
int avi_soft_clip_loss_grad(const float* pt, const float* tt, float* lse_scratch, float* dsim, double* loss, int32_t B, float temp,
                            void* stream);
/* the same update for a whole parameter list in ONE launch: a device-resident table of n_entries records */
typedef struct AviAdamwEntry {
  float* p;            /* parameter, updated in place */
  const float* g;      /* gradient                    */
  float* m;            /* first moment                */
  float* v;            /* second moment               */
  int64_t n;           /* elements                    */
  float weight_decay;  /* 0 for the exempt group      */
  int32_t reserved;
} AviAdamwEntry;
int avi_adamw_multi(const AviAdamwEntry* table_dev, int32_t n_entries, float lr, float beta1, float beta2, float eps, int32_t step,
                    float grad_scale, void* stream);
/* torch.optim.AdamW (decoupled weight decay), one parameter tensor per call */
int avi_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int32_t step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVI_B200_H_ */
