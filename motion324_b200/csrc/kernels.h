// Internal C++ interface between the kernels and the C-ABI layer (capi.cu).  Raw device pointers + sizes only.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace m324 {

const char* last_error();
long long launch_count();
int get_tuning(int knob);
void set_tuning(int knob, int value);

// ---- tcgen05 GEMM (gemm.cu) -------------------------------------------------------------------------------
struct GemmArgs {
  const __half* A; long lda;   // [M, K] (or [M, a_lo_off + K] when passes == 3), K-major
  const __half* W; long ldw;   // [N, K] (nn.Linear weight layout), K-major
  int M, N, K;                 // K = reduction length per pass, multiple of 64
  int passes;                  // 1, or 3 = split-fp16 (hi.hi + lo.hi + hi.lo)
  int a_lo_off, w_lo_off;      // column offset of the lo halves when passes == 3
  int bf16;                    // operands are bf16 instead of fp16
  const float* bias;           // [N] or null
  const float* gamma;          // [N] or null (LayerScale)
  const float* resid; long ldr; int resid_mod; long resid_div;  // fp32 residual; row -> (row / resid_div) * resid_mod + row % resid_mod when resid_mod > 0
  float* out32; long ldo32;    // fp32 output or null
  __half* out16; long ldo16;   // fp16 output or null
  int out16_lo_off;            // > 0: also write the fp16 remainder (x - half(x)) at this column offset
  int act;                     // 0 none, 1 GELU(erf)
  const float* qn_w; const float* kn_w; float qk_eps; int qk_cols;  // per-64-col RMSNorm on cols [0,qk_cols) / [qk_cols,2qk_cols)
  int force_bn128;             // testing: force the 128x128 tile
  // ---- training (backward) extensions; all zero / null on the forward-only path ----
  int tn;                      // 1: C[m,n] = sum_k A[k,m] * W[k,n]; A is [K, M] (lda), W is [K, N] (ldw), both MN-major: weight
                               //    gradients dW = dY^T X straight from the row-major activations, no transposes; K arbitrary
  int ksplit;                  // > 1: the K range is split over this many CTAs per output tile (needs accumulate)
  int accumulate;              // 1: out32 += result (TMA reduce-add; how gradients accumulate and split-K partials meet)
  __half* aux16; long ldaux;   // [M, N] fp16 side tensor
  int aux_mode;                // 1: store the pre-activation value there (forward, training)   2: multiply by gelu'(aux) (backward)
  int out16_bf16;              // out16 elements are bfloat16
  float out_scale;             // != 0: result *= out_scale before it is stored / accumulated
  float* qk_rstd; long ld_rstd;  // forward, training: [M, 2 * qk_cols / 64] reciprocal RMS of every normalised q / k head
  int fast_resid;                // set by gemm(): the non-in-place residual is prefetched by TMA (2-CTA kernel, full epilogue)
  const float* head_w; float* head_part;   // fused 3-channel head (shared_mlp_output.3): head_w [3, N] fp32; head_part [M, N / 64, 4] fp32 receives,
                                            // per row and 64-column group, the partial dot products of the epilogue result with the three rows
};
int head3_from_partials(const float* part, int groups, const float* b3, long rows, float* out, const float* target, float* partials,
                        int* n_partials, cudaStream_t stream);
int gemm(const GemmArgs& a, cudaStream_t stream);

// ---- tcgen05 flash attention forward (attention.cu) ------------------------------------------------------------
// softmax(q k^T * scale) v per (batch, head); head dim 64; q/k/v fp16 with heads as 64-column groups of a row.
struct AttnArgs {
  const __half* q; long q_ld; long q_rows;   // q_rows: rows addressable from q (tensor-map bound)
  const __half* k; long k_ld;
  const __half* v; long v_ld; long kv_rows;  // rows addressable from k and v
  int B, H, Lq, Lk;
  long q_batch_rows, kv_batch_rows;          // row offset between consecutive batches (0 = shared by all batches)
  int q_batch_div;                           // q rows of batch b start at (b / q_batch_div) * q_batch_rows (>= 1)
  __half* out; long o_ld;                    // out row = b * Lq + l, head h at columns [64h, 64h+64)
  float scale;
  int tune_event, tune_skew;                 // m324_set_tuning knobs 0 / 1: work-item shape (0 auto, 1 pair, 2 split); reserved
  float* lse; long lse_ld;                   // training: log2-domain log-sum-exp per (out row, head), [B*Lq, >= H] fp32, or null
  float* ws; long ws_bytes;                  // optional scratch (attention_workspace_bytes()) for the tail split; null = off
  int partial_parts, partial_index;          // > 0: this launch covers one of `partial_parts` K/V ranges of every query row: all
                                             // work items leave (O, m, l) in ws (slot = item * parts + index); attention_merge()
                                             // combines them.  ws must hold attention_partial_bytes(B, H, Lq, parts).
  int n_qt, items_whole, split_parts, split_slots, frame_loop;   // set by attention(): work-item decomposition (see attn_kernel)
  int item_loop;                             // 1: attn_items_kernel (persistent CTAs over contiguous chunks of short work items)
};
int attention(const AttnArgs& a, cudaStream_t stream);
int attention_plan(AttnArgs& a, int sms, int* grid_x, long* merge_rows);   // host-only work decomposition of attention()
// Combine the partial results of `partial_parts` attention() launches over disjoint K/V ranges (log-sum-exp merge) into
// out (and lse).  Only B, H, Lq, out, o_ld, lse, lse_ld, ws, partial_parts of the args are read.
int attention_merge(const AttnArgs& a, cudaStream_t stream);
long attention_partial_bytes(int B, int H, int Lq, int parts);
long attention_workspace_bytes();

// ---- tcgen05 flash attention backward (attention_bwd.cu) ---------------------------------------------------------
// Same operand addressing as the forward.  dO fp16 [B*Lq, do_ld]; lse / D fp32 [B*Lq, ld] (forward lse, attn_dot);
// dQ / dK / dV fp32 addressed like q / k / v.  dQ is ACCUMULATED (TMA reduce-add: zero it first; a query operand shared by
// several batches receives the sum), dK / dV are written.
struct AttnBwdArgs {
  const __half* q; long q_ld; long q_rows;
  const __half* k; long k_ld;
  const __half* v; long v_ld; long kv_rows;
  int B, H, Lq, Lk;
  long q_batch_rows, kv_batch_rows;
  int q_batch_div;
  const __half* dO; long do_ld;
  const float* lse; long lse_ld;
  const float* D; long d_ld;
  float* dQ; long dq_ld;
  float* dK; long dk_ld;
  float* dV; long dv_ld;
  float scale;
  int tune;                                  // m324_set_tuning knob 3 (experiments only): bit 0 = drop the dQ reduce-add, bit 1 = no exponentials
};
int attention_bwd(const AttnBwdArgs& a, cudaStream_t stream);

// ---- HBM-bound kernels (pointwise.cu) ------------------------------------------------------------------------
// LayerNorm over the last dim (fp32 in) -> fp16 out (optionally hi|lo split) and/or fp32 out.
// Source rows may be gathered: src_rpg > 0 -> src row = (r / src_rpg) * src_gstride + src_goff + r % src_rpg (this is the
// exact token slice [:, :, 4:4+tokens] of Pcd_motion.py:520 feeding the decoder's norm_kv).
int layernorm(const float* x, long ldx, const float* w, const float* b, float eps, long rows, int cols, int src_rpg,
              long src_gstride, long src_goff, __half* out16, long ldo16, int lo_off, float* out32, long ldo32,
              cudaStream_t stream);
// PointEmbed features (Pcd_motion.py:177-187): row = [sin(x.basis) (24), cos(x.basis) (24), x (3), 0-pad to 64] as hi|lo fp16.
int point_embed_features(const float* xyz, int n, __half* out, long ldo, int lo_off, cudaStream_t stream);
// Fill columns [768, 768+6) of the point-feature operand with (normal, rgb) as hi|lo fp16; zero the K padding.
int point_extra_features(const float* normal, const float* rgb, int n, __half* out, long ldo, int col0, int kpad,
                         int lo_off, cudaStream_t stream);
// rgb_video [F, Hin, Win, 3] fp32 in [0,1] -> bilinear resize to SxS (align_corners=False) -> ImageNet normalise ->
// im2col patches [F*hp*hp, kpad] fp16 (k = c*196 + py*14 + px), zero padded to kpad.
int preprocess_frames(const float* video, int F, int Hin, int Win, int S, __half* patches, long ldp, int kpad,
                      cudaStream_t stream);
// DINOv2 token assembly: x[f, 0] = cls + pos[0]; x[f, 1+i] = patch[f*np + i] + pos[1+i]   (fp32)
int dino_assemble(const float* patch, const float* cls, const float* pos, int F, int np, int C, float* x, cudaStream_t stream);
// Final DINO LayerNorm fused with trunk token assembly + transformer_input_layernorm (Pcd_motion.py:489-509):
//   tokens[b,t,0:4] = special, [4:4+M] = mesh_feat[b], [4+M: ] = LN_dino(x[f,1:]) + pos_embed[t]; then LN(no bias) -> fp32.
int assemble_tokens(const float* dino_x, const float* dino_nw, const float* dino_nb, float dino_eps, const float* pos_embed,
                    const float* sp0, const float* sprest, const float* mesh_feat, const float* ln_w, float ln_eps,
                    int B, int T, int ntok, int npatch, int C, float* out, float drop_p, unsigned long long seed,
                    float* pre_out, cudaStream_t stream);
// out[r, 0:3] = h[r, :] . W3^T + b3 (fp32), optional squared-error partial sums against target (per-block partials).
int head3_mse(const float* h, long ldh, const float* w3, const float* b3, long rows, int C, float* out,
              const float* target, float* partials, int* n_partials, int pre_gelu, cudaStream_t stream);
// Deterministic final reduce: loss[0] = mean sq err, loss[1] = weight * loss[0].
int mse_finalize(const float* partials, int n, double count, float weight, float* loss, cudaStream_t stream);
// Standalone MSE (model/loss.py:59-61) over n floats.
int mse_loss(const float* pred, const float* target, long n, float weight, float* partials, float* loss, cudaStream_t stream);
// fp32 -> fp16 cast of a [rows, cols] matrix into a [rows, ldo] buffer (zero K-padding), optional hi|lo split.
int cast_pad_f16(const float* src, long lds, int rows, int cols, __half* dst, long ldo, int kpad, int lo_off,
                 cudaStream_t stream);

// smooth_trajectories (utils/inference_utils.py:99-145), methods threshold / gaussian / combined.
int smooth_trajectories(const float* trajs, float* out, int B, int T, int N, float motion_threshold, float sigma, int do_threshold,
                        int do_gaussian, cudaStream_t stream);

// ---- backward pass, HBM-bound kernels (backward.cu); gradients of activations are in units of 1/alpha ------------------
int layernorm_bwd(const float* dy, long lddy, const float* x, long ldx, const float* w, float eps, long rows, int cols, int src_rpg,
                  long src_gstride, long src_goff, const float* dres, long lddres, float* dx32, long lddx32, __half* dx16,
                  long lddx16, float* dgamma, float* dbeta, float alpha, cudaStream_t stream);
int qknorm_bwd(const float* d_in, long ld_in, const __half* y16, long ldy, const float* rstd, long ld_rstd, const float* wq,
               const float* wk, int q_cols, int norm_cols, int cols, long rows, __half* out16, long ldo, float* dwq, float* dwk,
               float alpha, cudaStream_t stream);
int head_bwd(const float* pred, const float* target, const float* u, long ldu, const float* w3, long rows, int C, __half* du16,
             long lddu, float* dw3, float* db3, float alpha, cudaStream_t stream);
int colsum(const __half* dy, long ld, long rows, int cols, float* db, float alpha, cudaStream_t stream);
int sum_groups(const float* in, long ld_in, int ngroups, long group_stride, int rpg, long in_gstride, long in_goff, long rows,
               int cols, float scale, int accumulate, float* out32, long ldo32, __half* out16, long ldo16, cudaStream_t stream);
int cast_transpose_f16(const float* src, long lds, int N, int K, __half* dst, long ldo, int npad, cudaStream_t stream);
int attn_dot(const __half* dO, long lddo, const __half* O, long ldo, long rows, int H, float* D, long ldd, cudaStream_t stream);
int add_block(const float* in, long ld_in, long rows, int cols, float scale, int accumulate, float* out, long ldo, cudaStream_t stream);

// ---- data-prep gathers (dataprep.cu): dataset/dataset_utils.py:44-136, 19-41 --------------------------------------
// Barycentric tracking of S sampled surface points through T frames (+ interpolated, normalised vertex normals) and the
// UV-texture colour lookup.  verts / vnormals [T, V, 3] fp32 (f64 = 0) or fp64 (f64 = 1); faces [F, 3] int64; face_idx [S]
// int64; bary [S, 3] fp64; err: device int set non-zero on an out-of-range face / vertex index (checked by the caller).
int track_points(const void* verts, const void* vnormals, int f64, int T, long V, const long* faces, long F, const long* face_idx,
                 const double* bary, int S, float* points, float* normals, int* err, cudaStream_t stream);
// face_uvs [F, 3, 2] fp64; tex [H, W, 3] uint8; rgb [S, 3] fp32 in [0, 1]; texel (optional) [S, 2] int64 = (y, x) gathered.
int filter_trajectories(const float* trajs, float* out, int B, int T, int N, int mode, const double* taps_host, int ntaps, float mincutoff,
                        float beta, cudaStream_t stream);
int scale_by_device_scalars(float* buf, long n, const float* sa, const float* sb, float cb, cudaStream_t stream);
int sample_albedo(const double* verts, long V, const long* faces, long F, const double* uv, const long* face_idx, const double* points, int S,
                  const unsigned char* tex, int H, int W, float* rgb, long* texel, int* err, cudaStream_t stream);
int sample_texture(const double* face_uvs, long F, const long* face_idx, const double* bary, int S, const unsigned char* tex, int H, int W,
                   float* rgb, long* texel, int* err, cudaStream_t stream);

// ---- point-cloud evaluation metrics (chamfer.cu) ---------------------------------------------------------------
// Bidirectional exact nearest neighbours (float64 arithmetic) for `frames` independent frames: p1 [frames, n1, 3],
// p2 [frames, n2, 3] (fp32 or fp64).  dist1 / idx1 [frames, n2]: for every point of p2 its nearest point of p1;
// dist2 / idx2 [frames, n1]: the other direction.  idx pointers may be null.
int chamfer_nn(const void* p1, int n1, const void* p2, int n2, int frames, int is_f64, double* dist1, int* idx1, double* dist2,
               int* idx2, cudaStream_t stream);
// out [frames, 4] = { chamfer = mean(dist1) + mean(dist2), F-score, precision, recall } at `threshold`.
int chamfer_reduce(const double* dist1, int n2, const double* dist2, int n1, int frames, double threshold, double* out,
                   cudaStream_t stream);

}  // namespace m324
