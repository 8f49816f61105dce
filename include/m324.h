/* m324.h -- C ABI of libm324.so: the B200-native (sm_100a) kernels behind the Motion324 per-frame motion-estimation
 * hot path, Motion_Latent_Model.forward (reference: model/Pcd_motion.py:450-598).
 *
 * The reference has NO native interface for this path: it is pure PyTorch calling library kernels.  The
 * boundary a maintainer binds against is therefore the set of operator call sites listed per function below; each entry
 * point replaces the library calls at those sites.  INTEGRATION.md shows the ctypes binding (the reference is Python).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (outputs and workspaces included); the library allocates
 *     nothing and keeps no state besides one-time kernel attribute setup;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous, capture-safe (no sync, no malloc);
 *   - return value: 0 = ok, negative = error (M324_ERR_*); m324_last_error() returns a thread-local message;
 *     nothing throws across the ABI and there is NO CPU fallback: unsupported shapes are errors;
 *   - "f16" pointers are IEEE binary16 (tensor-core operands); all statistics / residuals / outputs are fp32;
 *   - "hi|lo split": a value x is stored as h = fp16(x) at column c and fp16(x - h) at column c + lo_off; GEMMs run with
 *     passes = 3 on such operands (A_hi.W_hi + A_lo.W_hi + A_hi.W_lo), see DESIGN.md "Precision".
 */
#ifndef M324_H_
#define M324_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M324_OK 0
#define M324_ERR_INVALID (-1)
#define M324_ERR_CUDA (-2)
#define M324_ERR_DRIVER (-3)
#define M324_ERR_UNSUPPORTED (-4)

int m324_version(void);
const char* m324_last_error(void);
/* Number of CUDA kernels this library has launched in this process (every launch is counted, e.g. an attention call that
 * splits its tail wave launches two): the figure bench.py reports as gpu_launches. */
int64_t m324_launch_count(void);
/* Performance-tuning knobs (results stay within tolerance): knob 0 = attention work-item shape (0 = pair kernel with the
 * tail split when a workspace is lent, 1 = pair kernel, never split, 2 = in-CTA K/V-split kernel); knob 1 = 1 disables the
 * MUFU turn-taking of the attention softmax warps; knob 2 = 1 disables programmatic dependent launch; knob 3 = experiment
 * bits of the attention backward (0 in production); other knobs reserved. */
int m324_set_tuning(int32_t knob, int32_t value);
/* 0 when the current device is sm_100 (B200); M324_ERR_UNSUPPORTED otherwise. */
int m324_check_device(void);

/* C[M,N] = epilogue(A[M,K] . W[N,K]^T): every nn.Linear on the path -- transformer.py:124-126 (to_q/k/v), :200 (to_qkv),
 * :217 / :142 (fc), :73-78 (MLP); Pcd_motion.py:186 (point_embed.mlp), :459/:551 (point_normal_rgb_proj), :561
 * (shared_mlp_output.1); DINOv2 qkv/proj/fc1/fc2/patch_embed behind image_encoder/dinov2.py:99.
 * Fused epilogue: per-head RMS q/k-norm (transformer.py:30-42, 130-132, 205-207), bias, GELU(erf), LayerScale gamma,
 * fp32 residual add (transformer.py:375-376, 421-422). */
typedef struct {
  const void* A; int64_t lda;       /* f16 [M, K] (or [M, a_lo_off + K] if passes == 3) */
  const void* W; int64_t ldw;       /* f16 [N, K] (nn.Linear weight layout) */
  int32_t M, N, K;                  /* K multiple of 64, N multiple of 4 */
  int32_t passes;                   /* 1, or 3 for hi|lo split operands */
  int32_t a_lo_off, w_lo_off;
  int32_t bf16;                     /* operands are bfloat16 instead of binary16 */
  const float* bias;                /* [N] or NULL */
  const float* gamma;               /* [N] or NULL */
  const float* resid; int64_t ldr; int32_t resid_mod; int64_t resid_div;
                                    /* fp32 residual; resid_mod > 0: row -> (row / resid_div) * resid_mod + row % resid_mod
                                       (resid_div == 0: row % resid_mod) */
  float* out32; int64_t ldo32;      /* fp32 output or NULL (may alias resid) */
  void* out16; int64_t ldo16;       /* f16 output or NULL */
  int32_t out16_lo_off;             /* > 0: write the hi|lo split */
  int32_t act;                      /* 0 none, 1 GELU(erf) */
  const float* qn_w; const float* kn_w; float qk_eps; int32_t qk_cols;
  int32_t force_bn128;
  /* ---- training (backward) extensions; zero / NULL on the forward-only path ----
   * The backward of every nn.Linear on the path (autograd of transformer.py:73-78,124-126,200,217 and Pcd_motion.py:186,
   * 459,551-553,561 under train.py:157-170) is two more GEMMs: dX = dY . W (A = dY, W = the transposed weight) and
   * dW += alpha * dY^T . X (tn = 1, straight from the row-major activations, split over K, accumulated in fp32).
   * Gradients travel as f16 in units of 1 / alpha (alpha = dLoss * 2 w / n, so that the seed is pred - target, O(0.1)). */
  int32_t tn;                       /* 1: C[m,n] = sum_k A[k,m] W[k,n]; A is [K, M] (lda), W is [K, N] (ldw); K arbitrary */
  int32_t ksplit;                   /* > 1: split the K range over this many CTAs per tile (needs accumulate = 1) */
  int32_t accumulate;               /* 1: out32 += result (fp32 adds in L2 by the TMA unit; order not fixed) */
  void* aux16; int64_t ldaux;       /* [M, N] f16 side tensor */
  int32_t aux_mode;                 /* 1: store the pre-activation there; 2: multiply the result by gelu'(aux) */
  int32_t out16_bf16;               /* out16 elements are bfloat16 */
  float out_scale;                  /* != 0: result *= out_scale before it is stored / accumulated */
  float* qk_rstd; int64_t ld_rstd;  /* [M, 2 * qk_cols / 64]: reciprocal RMS of every normalised q / k head */
  const float* head_w; float* head_part;   /* fused 3-channel output head (Pcd_motion.py:340, 561: shared_mlp_output.3 folded into the epilogue of
                                              shared_mlp_output.1): head_w [3, N] fp32; head_part [M, N / 64, 4] fp32 receives per row and 64-column
                                              group the partial dot products of the epilogue result with the three rows (out32 / out16 may then be
                                              NULL); m324_head3_from_partials finishes.  Both NULL = off */
} m324_gemm_args;
int m324_gemm(const m324_gemm_args* args, void* stream);

/* xformers.ops.memory_efficient_attention(q, k, v, attn_bias=None, p=0.0) -- transformer.py:134-139, 209-214; layout
 * [B, L, H, 64] with arbitrary row strides (v is a strided view of the packed qkv, transformer.py:200-202). */
typedef struct {
  const void* q; int64_t q_ld; int64_t q_rows;
  const void* k; int64_t k_ld;
  const void* v; int64_t v_ld; int64_t kv_rows;
  int32_t B, H, Lq, Lk;
  int64_t q_batch_rows, kv_batch_rows;   /* 0 = operand shared by all batches */
  int32_t q_batch_div;                   /* q rows of batch b start at (b / q_batch_div) * q_batch_rows; >= 1 */
  void* out; int64_t o_ld;               /* f16 [B*Lq, >= H*64] */
  float scale;                           /* Dh^-0.5 */
  float* lse; int64_t lse_ld;            /* training: log2-domain log-sum-exp per (out row, head), fp32 [B*Lq, >= H]; NULL = off */
  void* workspace; int64_t workspace_bytes;   /* optional caller-owned scratch (m324_attention_workspace_bytes(), 16-byte aligned):
                                                 lets a launch whose last wave of work items is partly filled split those items
                                                 over K/V ranges and merge them (a second small kernel).  NULL = never split */
  int32_t partial_parts, partial_index;       /* 0, 0 = a complete attention.  partial_parts = P > 0: this call covers ONE of P disjoint
                                                 K/V ranges of the same queries (k / v / Lk describe that range only); every query row's
                                                 unnormalised result goes to the workspace (m324_attention_partial_bytes) and
                                                 m324_attention_merge combines the P calls.  Lets a caller start on the keys it already has
                                                 while the others are still in flight (frame-sharded clips: local keys first, the
                                                 all-gathered ones after) */
} m324_attn_args;
int m324_attention(const m324_attn_args* args, void* stream);
int64_t m324_attention_workspace_bytes(void);
/* Host-only: the work decomposition m324_attention would use for these args on a device with sm_count SMs (no pointer but
 * `workspace` is inspected, nothing is launched).  plan[0..7] = { Q-tile pairs per (batch, head), frames per CTA (frame loop,
 * 1 = off), work items that walk all their K/V tiles, K/V ranges per split item, workspace slots, grid size of the attention
 * kernel, grid size of the merge kernel (0 = none), reserved }. */
int m324_attention_plan(const m324_attn_args* args, int32_t sm_count, int32_t* plan);
int64_t m324_attention_partial_bytes(int32_t B, int32_t H, int32_t Lq, int32_t parts);
/* Log-sum-exp merge of the partial_parts partial calls into out (and lse): reads B, H, Lq, out, o_ld, lse, lse_ld, workspace,
 * workspace_bytes, partial_parts of args. */
int m324_attention_merge(const m324_attn_args* args, void* stream);

/* Backward of the same call (the BwOp of transformer.py:134-139, 209-214).  Operand addressing as in the forward; dO f16
 * [B*Lq, do_ld]; lse from the forward, D from m324_attn_dot; dQ / dK / dV fp32 addressed like q / k / v.  dQ is ACCUMULATED
 * (zero it first; a query operand shared by several batches receives the sum over them), dK / dV are written. */
typedef struct {
  const void* q; int64_t q_ld; int64_t q_rows;
  const void* k; int64_t k_ld;
  const void* v; int64_t v_ld; int64_t kv_rows;
  int32_t B, H, Lq, Lk;
  int64_t q_batch_rows, kv_batch_rows;
  int32_t q_batch_div;
  const void* dO; int64_t do_ld;
  const float* lse; int64_t lse_ld;
  const float* D; int64_t d_ld;
  float* dQ; int64_t dq_ld;
  float* dK; int64_t dk_ld;
  float* dV; int64_t dv_ld;
  float scale;
} m324_attn_bwd_args;
int m324_attention_bwd(const m324_attn_bwd_args* args, void* stream);

/* nn.LayerNorm (transformer.py:345-357, 400, 411; Pcd_motion.py:326, 337) -> f16 GEMM operand and/or fp32. */
int m324_layernorm(const float* x, int64_t ldx, const float* w, const float* b, float eps, int64_t rows, int32_t cols,
                   int32_t src_rpg, int64_t src_gstride, int64_t src_goff, void* out16, int64_t ldo16, int32_t lo_off,
                   float* out32, int64_t ldo32, void* stream);
/* PointEmbed.embed (Pcd_motion.py:177-187) -> [n, 64] hi|lo operand. */
int m324_point_embed_features(const float* xyz, int32_t n, void* out, int64_t ldo, int32_t lo_off, void* stream);
/* torch.cat([emb, normal, rgb]) (Pcd_motion.py:459, 551-553) -> columns [col0, col0+6), zero K padding. */
int m324_point_extra_features(const float* normal, const float* rgb, int32_t n, void* out, int64_t ldo, int32_t col0,
                              int32_t kpad, int32_t lo_off, void* stream);
/* Pcd_motion.py:470-472 + image_encoder/dinov2.py:78-80 + patch-embed im2col. */
int m324_preprocess_frames(const float* video, int32_t F, int32_t Hin, int32_t Win, int32_t S, void* patches, int64_t ldp,
                           int32_t kpad, void* stream);
/* DINOv2 prepare_tokens (cls + interpolated position table). */
int m324_dino_assemble(const float* patch, const float* cls, const float* pos, int32_t F, int32_t np, int32_t C, float* x,
                       void* stream);
/* DINOv2 final norm + Pcd_motion.py:489-509 (pos_embed add, pos_drop, token concat, transformer_input_layernorm).
 * drop_p > 0 (train() only): nn.Dropout(p) of Pcd_motion.py:369-370,490 on the image tokens, keep mask from a counter hash of
 * (seed, element index).  pre_ln_out (training, may be NULL): the concatenated tokens before transformer_input_layernorm,
 * [B*T*(4+ntok+npatch), C] fp32, kept for the backward pass. */
int m324_assemble_tokens(const float* dino_x, const float* dino_nw, const float* dino_nb, float dino_eps,
                         const float* pos_embed, const float* sp0, const float* sprest, const float* mesh_feat,
                         const float* ln_w, float ln_eps, int32_t B, int32_t T, int32_t ntok, int32_t npatch, int32_t C,
                         float* out, float drop_p, uint64_t seed, float* pre_ln_out, void* stream);
/* shared_mlp_output.3 (Pcd_motion.py:340, 561) + squared-error partials of MSELossComputer (model/loss.py:59-61).
 * pre_gelu = 1 (training): h holds the pre-activation of shared_mlp_output.2 (kept for the backward); GELU(erf) is applied here. */
int m324_head3_mse(const float* h, int64_t ldh, const float* w3, const float* b3, int64_t rows, int32_t C, float* out,
                   const float* target, float* partials, int32_t* n_partials, int32_t pre_gelu, void* stream);
int m324_mse_finalize(const float* partials, int32_t n, double count, float weight, float* loss, void* stream);
/* Second half of the fused head: out[r, c] = sum_g part[r, g, c] (index order) + b3[c]; squared-error partial sums against target like
 * m324_head3_mse (target / partials may be NULL).  part [rows, groups, 4] fp32 as written by m324_gemm (groups = N / 64). */
int m324_head3_from_partials(const float* part, int32_t groups, const float* b3, int64_t rows, float* out, const float* target, float* partials,
                             int32_t* n_partials, void* stream);
/* F.mse_loss * coord_mse_loss_weight (model/loss.py:59-61): loss[0] = mse, loss[1] = weight * mse. */
int m324_mse_loss(const float* pred, const float* target, int64_t n, float weight, float* partials, float* loss, void* stream);
/* fp32 -> f16 parameter / operand conversion with zero K padding and optional hi|lo split. */
int m324_cast_pad_f16(const float* src, int64_t lds, int32_t rows, int32_t cols, void* dst, int64_t ldo, int32_t kpad,
                      int32_t lo_off, void* stream);

/* smooth_trajectories(method = 'threshold' | 'gaussian' | 'combined') of utils/inference_utils.py:99-145
 * (scipy.ndimage.gaussian_filter1d, mode='nearest', truncate 4; fp64 accumulation).  trajs/out [B,T,N,3] fp32, out != trajs. */
int m324_smooth_trajectories(const float* trajs, float* out, int32_t B, int32_t T, int32_t N, float motion_threshold, float sigma,
                             int32_t do_threshold, int32_t do_gaussian, void* stream);

/* The other two methods of the same function: mode 1 = 'savgol' (utils/inference_utils.py:148-163: scipy.signal.savgol_filter,
 * mode='nearest'; taps_host = the ntaps (odd, <= 17) FIR coefficients, a HOST array copied by value into the launch), mode 2 =
 * 'oneeuro' (:58-96, 165-175: OneEuroFilter(mincutoff, beta, dcutoff = 1) per vertex and axis, fp32 like NumPy's float32 scalars).
 * No threshold pass (the reference applies it for 'threshold' / 'combined' only).  trajs/out [B,T,N,3] fp32, out != trajs. */
int m324_filter_trajectories(const float* trajs, float* out, int32_t B, int32_t T, int32_t N, int32_t mode, const double* taps_host,
                             int32_t ntaps, float mincutoff, float beta, void* stream);

/* ---- backward pass (what torch.autograd runs under train.py:157-170 for this model) -----------------------
 * Activation gradients are carried in units of 1/alpha, alpha = dLoss * 2 * coord_mse_loss_weight / n (the seed is
 * pred - target): f16 between GEMMs, fp32 on the residual stream.  Parameter gradients are ACCUMULATED (+=) into fp32
 * buffers in true units (alpha applied), like autograd's .grad. */
/* nn.LayerNorm backward (transformer.py:345-357,400,411; Pcd_motion.py:326,337).  mean / rstd are recomputed from the saved
 * input x.  src_* : the forward's gathered-row mapping (applies to x, dres and dx; dy is compact).  dx = dres + LN'(dy). */
int m324_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* w, float eps, int64_t rows, int32_t cols,
                       int32_t src_rpg, int64_t src_gstride, int64_t src_goff, const float* dres, int64_t lddres, float* dx32,
                       int64_t lddx32, void* dx16, int64_t lddx16, float* dgamma, float* dbeta, float alpha, void* stream);
/* RMSNorm backward of the q / k heads (transformer.py:30-42,130-132,205-207) + fp32 -> f16 of the attention gradients:
 * d_in fp32 [rows, cols] = dQ | dK | dV, y16 = the forward's normalised q / k, rstd from m324_gemm (qk_rstd). */
int m324_qknorm_bwd(const float* d_in, int64_t ld_in, const void* y16, int64_t ldy, const float* rstd, int64_t ld_rstd, const float* wq,
                    const float* wk, int32_t q_cols, int32_t norm_cols, int32_t cols, int64_t rows, void* out16, int64_t ldo, float* dwq,
                    float* dwk, float alpha, void* stream);
/* shared_mlp_output.2-3 + MSE backward (Pcd_motion.py:340,561; model/loss.py:59-61): u = pre-GELU input of the 768->3 layer. */
int m324_head_bwd(const float* pred, const float* target, const float* u, int64_t ldu, const float* w3, int64_t rows, int32_t C, void* du16,
                  int64_t lddu, float* dw3, float* db3, float alpha, void* stream);
/* bias gradient: db[c] += alpha * sum_rows dy16[row, c] */
int m324_colsum(const void* dy16, int64_t ld, int64_t rows, int32_t cols, float* db, float alpha, void* stream);
/* gradient of an operand that the forward broadcast to `ngroups` frames (Pcd_motion.py:495-507, 539-560): sum over groups */
int m324_sum_groups(const float* in, int64_t ld_in, int32_t ngroups, int64_t group_stride, int32_t rpg, int64_t in_gstride, int64_t in_goff,
                    int64_t rows, int32_t cols, float scale, int32_t accumulate, float* out32, int64_t ldo32, void* out16, int64_t ldo16,
                    void* stream);
/* fp32 weight [N, K] -> f16 W^T [K, npad] (the operand of dX = dY . W) */
int m324_cast_transpose_f16(const float* src, int64_t lds, int32_t N, int32_t K, void* dst, int64_t ldo, int32_t npad, void* stream);
/* out[r, c] (+)= scale * in[r, c], c < cols, any row strides: un-pads a weight gradient computed at the K-padded operand width
 * (point_embed.mlp 51 -> 64, point_normal_rgb_proj 774 -> 832) and rescales gradients in place (in == out). */
int m324_add_block(const float* in, int64_t ld_in, int64_t rows, int32_t cols, float scale, int32_t accumulate, float* out, int64_t ldo,
                   void* stream);
/* ---- data-prep gathers, SURVEY.md 8(f4) ------------------------------------------------------------------------------
 * dataset/dataset_utils.py:44-136 track_with_normal_rgb after the host-side surface sampling: S sampled points (face index +
 * barycentric coordinates, float64 as trimesh returns them) tracked through T frames.
 *   points[t, s]  = sum_c bary[s, c] * vertex_frames[t, faces[face_indices[s], c]]                 (:112-114)
 *   normals[t, s] = normalise(sum_c bary[s, c] * vertex_normals[t, faces[face_indices[s], c]])     (:116-127), zero norm kept
 * vertex_frames / vertex_normals: [T, V, 3] fp32 (is_f64 = 0) or fp64 (is_f64 = 1); vertex_normals / normals may both be NULL.
 * float64 arithmetic in NumPy's operation order, one rounding to fp32 at the end (:131-132).  err_flag: device int32, set
 * non-zero if a face or vertex index is out of range (NumPy would raise IndexError); the caller reads it. */
int m324_track_points(const void* vertex_frames, const void* vertex_normals, int32_t is_f64, int32_t T, int64_t V, const int64_t* faces,
                      int64_t F, const int64_t* face_indices, const double* barycentric, int32_t S, float* points, float* normals,
                      int32_t* err_flag, void* stream);
/* dataset/dataset_utils.py:85-99 + 19-41 sample_texture_color_vectorized: uv = sum_c bary * face_uvs[face_indices] (float64),
 * x = clip(int(u * (W - 1))), y = clip(int((1 - v) * (H - 1))), rgb = texture[y, x] / 255.  face_uvs [F, 3, 2] fp64, texture
 * [H, W, 3] uint8, rgb [S, 3] fp32; texel_yx (optional) [S, 2] int64 receives the gathered (y, x): bit-exact integer work. */
int m324_sample_texture_colors(const double* face_uvs, int64_t F, const int64_t* face_indices, const double* barycentric, int32_t S,
                               const uint8_t* texture, int32_t H, int32_t W, float* rgb, int64_t* texel_yx, int32_t* err_flag,
                               void* stream);
/* buf[0..n) *= (*scalar_a + coeff_b * *scalar_b), scalars read from DEVICE memory (either may be NULL): the upstream gradient of
 * ``(loss / grad_accum_steps).backward()`` (train.py:159-166) applied to the flat gradient buffer without a host sync. */
int m324_scale_by_device_scalars(float* buf, int64_t n, const float* scalar_a, const float* scalar_b, float coeff_b, void* stream);
/* utils/mesh_processing.py:130-191 sample_pointcloud_with_albedo, its per-sample Python loop (:174-182, called from
 * scripts/inference_with_video_mesh.py:107-111): barycentric coordinates re-derived from the sampled point and its triangle
 * (:107-127), uv = sum_c w_c * (uv[faces[f, c]] mod 1), x = int(clip(u * W, 0, W - 1)), y = int(clip((1 - v) * H, 0, H - 1)),
 * rgb = float32(texture[y, x]) / 255.  vertices [V, 3], uv [V, 2], points [S, 3]: fp64; faces [F, 3], face_indices [S]: int64;
 * texel_yx (optional) [S, 2] int64: bit-exact integer work.  err_flag: 1 = face index, 2 = vertex index out of range. */
int m324_sample_albedo(const double* vertices, int64_t V, const int64_t* faces, int64_t F, const double* uv, const int64_t* face_indices,
                       const double* points, int32_t S, const uint8_t* texture, int32_t H, int32_t W, float* rgb, int64_t* texel_yx,
                       int32_t* err_flag, void* stream);
/* D[row, h] = sum_d dO[row, 64h+d] * O[row, 64h+d]: the row term of the softmax backward */
int m324_attn_dot(const void* dO, int64_t lddo, const void* O, int64_t ldo, int64_t rows, int32_t H, float* D, int64_t ldd, void* stream);

/* evaluation/evaluation_pcd.py:575-588 (compute_chamfer_distance) and :591-609 (compute_fscore), i.e. the
 * two scipy.spatial.cKDTree builds + k=1 queries per frame (:884-885, 50 000 x 50 000 points).  Exact brute-force nearest
 * neighbours in float64 (the reference's dtype), ties -> smallest index; `frames` independent frames per call.
 *   points1 [frames, n1, 3], points2 [frames, n2, 3]: fp32 (is_f64 = 0) or fp64 (is_f64 = 1)
 *   dist1 / idx1 [frames, n2] = tree1.query(points2);  dist2 / idx2 [frames, n1] = tree2.query(points1); idx may be NULL */
int m324_chamfer_nn(const void* points1, int32_t n1, const void* points2, int32_t n2, int32_t frames, int32_t is_f64,
                    double* dist1, int32_t* idx1, double* dist2, int32_t* idx2, void* stream);
/* out [frames, 4] = { mean(dist1) + mean(dist2), F-score = 2PR/(P+R) (0 when P+R = 0), P = mean(dist1 < threshold),
 * R = mean(dist2 < threshold) }; fixed reduction order (bit-reproducible). */
int m324_chamfer_reduce(const double* dist1, int32_t n2, const double* dist2, int32_t n1, int32_t frames, double threshold,
                        double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* M324_H_ */
