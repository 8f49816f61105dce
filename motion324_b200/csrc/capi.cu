// extern "C" surface of libm324.so (declared in include/m324.h): plain pointers and sizes in, int status out.
#include "../../include/m324.h"

#include "common.cuh"
#include "kernels.h"

using namespace m324;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int m324_version(void) { return 100; }

const char* m324_last_error(void) { return m324::last_error(); }

int64_t m324_launch_count(void) { return launch_count(); }

int m324_set_tuning(int32_t knob, int32_t value) {
  set_tuning(knob, value);
  return M324_OK;
}

int m324_check_device(void) {
  int dev = 0;
  M324_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  M324_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  M324_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("device compute capability %d.%d is not sm_100 (B200); libm324 has no other code path", major, minor);
    return M324_ERR_UNSUPPORTED;
  }
  return M324_OK;
}

int m324_gemm(const m324_gemm_args* a, void* stream) {
  M324_REQUIRE(a != nullptr, "m324_gemm: null args");
  GemmArgs g;
  g.A = static_cast<const __half*>(a->A); g.lda = a->lda;
  g.W = static_cast<const __half*>(a->W); g.ldw = a->ldw;
  g.M = a->M; g.N = a->N; g.K = a->K; g.passes = a->passes; g.a_lo_off = a->a_lo_off; g.w_lo_off = a->w_lo_off;
  g.bf16 = a->bf16; g.bias = a->bias; g.gamma = a->gamma; g.resid = a->resid; g.ldr = a->ldr; g.resid_mod = a->resid_mod; g.resid_div = a->resid_div;
  g.out32 = a->out32; g.ldo32 = a->ldo32; g.out16 = static_cast<__half*>(a->out16); g.ldo16 = a->ldo16;
  g.out16_lo_off = a->out16_lo_off; g.act = a->act; g.qn_w = a->qn_w; g.kn_w = a->kn_w; g.qk_eps = a->qk_eps;
  g.qk_cols = a->qk_cols; g.force_bn128 = a->force_bn128;
  g.tn = a->tn; g.ksplit = a->ksplit; g.accumulate = a->accumulate;
  g.aux16 = static_cast<__half*>(a->aux16); g.ldaux = a->ldaux; g.aux_mode = a->aux_mode; g.out16_bf16 = a->out16_bf16; g.out_scale = a->out_scale;
  g.qk_rstd = a->qk_rstd; g.ld_rstd = a->ld_rstd;
  g.head_w = a->head_w; g.head_part = a->head_part;
  return gemm(g, S(stream));
}

int m324_attention(const m324_attn_args* a, void* stream) {
  M324_REQUIRE(a != nullptr, "m324_attention: null args");
  AttnArgs t = {};
  t.q = static_cast<const __half*>(a->q); t.q_ld = a->q_ld; t.q_rows = a->q_rows;
  t.k = static_cast<const __half*>(a->k); t.k_ld = a->k_ld;
  t.v = static_cast<const __half*>(a->v); t.v_ld = a->v_ld; t.kv_rows = a->kv_rows;
  t.B = a->B; t.H = a->H; t.Lq = a->Lq; t.Lk = a->Lk; t.q_batch_rows = a->q_batch_rows; t.kv_batch_rows = a->kv_batch_rows; t.q_batch_div = a->q_batch_div;
  t.out = static_cast<__half*>(a->out); t.o_ld = a->o_ld; t.scale = a->scale;
  t.tune_event = get_tuning(0); t.tune_skew = get_tuning(1);
  t.lse = a->lse; t.lse_ld = a->lse_ld;
  t.ws = static_cast<float*>(a->workspace); t.ws_bytes = a->workspace_bytes;
  t.partial_parts = a->partial_parts; t.partial_index = a->partial_index;
  M324_REQUIRE(t.ws == nullptr || (reinterpret_cast<uintptr_t>(t.ws) & 15) == 0, "m324_attention: workspace must be 16-byte aligned");
  return attention(t, S(stream));
}

int m324_attention_plan(const m324_attn_args* a, int32_t sm_count, int32_t* plan) {
  M324_REQUIRE(a != nullptr && plan != nullptr, "m324_attention_plan: null args");
  AttnArgs t = {};
  t.B = a->B; t.H = a->H; t.Lq = a->Lq; t.Lk = a->Lk; t.q_batch_rows = a->q_batch_rows; t.kv_batch_rows = a->kv_batch_rows; t.q_batch_div = a->q_batch_div;
  t.tune_event = get_tuning(0); t.tune_skew = get_tuning(1);
  t.lse = a->lse;
  t.ws = static_cast<float*>(a->workspace); t.ws_bytes = a->workspace_bytes;
  t.partial_parts = a->partial_parts; t.partial_index = a->partial_index;
  int grid = 0;
  long merge_rows = 0;
  const int e = attention_plan(t, sm_count, &grid, &merge_rows);
  if (e) return e;
  plan[0] = t.n_qt; plan[1] = t.frame_loop; plan[2] = t.items_whole; plan[3] = t.split_parts; plan[4] = t.split_slots; plan[5] = grid;
  plan[6] = static_cast<int32_t>((merge_rows + 7) / 8); plan[7] = t.item_loop;
  return M324_OK;
}

int64_t m324_attention_workspace_bytes(void) { return attention_workspace_bytes(); }

int64_t m324_attention_partial_bytes(int32_t B, int32_t H, int32_t Lq, int32_t parts) { return attention_partial_bytes(B, H, Lq, parts); }

int m324_attention_merge(const m324_attn_args* a, void* stream) {
  M324_REQUIRE(a != nullptr, "m324_attention_merge: null args");
  AttnArgs t = {};
  t.B = a->B; t.H = a->H; t.Lq = a->Lq; t.out = static_cast<__half*>(a->out); t.o_ld = a->o_ld; t.lse = a->lse; t.lse_ld = a->lse_ld;
  t.ws = static_cast<float*>(a->workspace); t.ws_bytes = a->workspace_bytes; t.partial_parts = a->partial_parts;
  return attention_merge(t, S(stream));
}

int m324_layernorm(const float* x, int64_t ldx, const float* w, const float* b, float eps, int64_t rows, int32_t cols,
                   int32_t src_rpg, int64_t src_gstride, int64_t src_goff, void* out16, int64_t ldo16, int32_t lo_off,
                   float* out32, int64_t ldo32, void* stream) {
  return layernorm(x, ldx, w, b, eps, rows, cols, src_rpg, src_gstride, src_goff, static_cast<__half*>(out16), ldo16, lo_off,
                   out32, ldo32, S(stream));
}

int m324_point_embed_features(const float* xyz, int32_t n, void* out, int64_t ldo, int32_t lo_off, void* stream) {
  return point_embed_features(xyz, n, static_cast<__half*>(out), ldo, lo_off, S(stream));
}

int m324_point_extra_features(const float* normal, const float* rgb, int32_t n, void* out, int64_t ldo, int32_t col0,
                              int32_t kpad, int32_t lo_off, void* stream) {
  return point_extra_features(normal, rgb, n, static_cast<__half*>(out), ldo, col0, kpad, lo_off, S(stream));
}

int m324_preprocess_frames(const float* video, int32_t F, int32_t Hin, int32_t Win, int32_t Sz, void* patches, int64_t ldp,
                           int32_t kpad, void* stream) {
  return preprocess_frames(video, F, Hin, Win, Sz, static_cast<__half*>(patches), ldp, kpad, S(stream));
}

int m324_dino_assemble(const float* patch, const float* cls, const float* pos, int32_t F, int32_t np, int32_t C, float* x,
                       void* stream) {
  return dino_assemble(patch, cls, pos, F, np, C, x, S(stream));
}

int m324_assemble_tokens(const float* dino_x, const float* dino_nw, const float* dino_nb, float dino_eps,
                         const float* pos_embed, const float* sp0, const float* sprest, const float* mesh_feat,
                         const float* ln_w, float ln_eps, int32_t B, int32_t T, int32_t ntok, int32_t npatch, int32_t C,
                         float* out, float drop_p, uint64_t seed, float* pre_ln_out, void* stream) {
  return assemble_tokens(dino_x, dino_nw, dino_nb, dino_eps, pos_embed, sp0, sprest, mesh_feat, ln_w, ln_eps, B, T, ntok,
                         npatch, C, out, drop_p, seed, pre_ln_out, S(stream));
}

int m324_head3_mse(const float* h, int64_t ldh, const float* w3, const float* b3, int64_t rows, int32_t C, float* out,
                   const float* target, float* partials, int32_t* n_partials, int32_t pre_gelu, void* stream) {
  int n = 0;
  int e = head3_mse(h, ldh, w3, b3, rows, C, out, target, partials, &n, pre_gelu, S(stream));
  if (n_partials) *n_partials = n;
  return e;
}

int m324_mse_finalize(const float* partials, int32_t n, double count, float weight, float* loss, void* stream) {
  return mse_finalize(partials, n, count, weight, loss, S(stream));
}

int m324_mse_loss(const float* pred, const float* target, int64_t n, float weight, float* partials, float* loss, void* stream) {
  return mse_loss(pred, target, n, weight, partials, loss, S(stream));
}

int m324_cast_pad_f16(const float* src, int64_t lds, int32_t rows, int32_t cols, void* dst, int64_t ldo, int32_t kpad,
                      int32_t lo_off, void* stream) {
  return cast_pad_f16(src, lds, rows, cols, static_cast<__half*>(dst), ldo, kpad, lo_off, S(stream));
}

int m324_smooth_trajectories(const float* trajs, float* out, int32_t B, int32_t T, int32_t N, float motion_threshold, float sigma,
                             int32_t do_threshold, int32_t do_gaussian, void* stream) {
  return smooth_trajectories(trajs, out, B, T, N, motion_threshold, sigma, do_threshold, do_gaussian, S(stream));
}


int m324_chamfer_nn(const void* points1, int32_t n1, const void* points2, int32_t n2, int32_t frames, int32_t is_f64,
                    double* dist1, int32_t* idx1, double* dist2, int32_t* idx2, void* stream) {
  return chamfer_nn(points1, n1, points2, n2, frames, is_f64, dist1, idx1, dist2, idx2, S(stream));
}

int m324_chamfer_reduce(const double* dist1, int32_t n2, const double* dist2, int32_t n1, int32_t frames, double threshold,
                        double* out, void* stream) {
  return chamfer_reduce(dist1, n2, dist2, n1, frames, threshold, out, S(stream));
}


int m324_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* w, float eps, int64_t rows, int32_t cols,
                       int32_t src_rpg, int64_t src_gstride, int64_t src_goff, const float* dres, int64_t lddres, float* dx32,
                       int64_t lddx32, void* dx16, int64_t lddx16, float* dgamma, float* dbeta, float alpha, void* stream) {
  return layernorm_bwd(dy, lddy, x, ldx, w, eps, rows, cols, src_rpg, src_gstride, src_goff, dres, lddres, dx32, lddx32,
                       static_cast<__half*>(dx16), lddx16, dgamma, dbeta, alpha, S(stream));
}

int m324_qknorm_bwd(const float* d_in, int64_t ld_in, const void* y16, int64_t ldy, const float* rstd, int64_t ld_rstd, const float* wq,
                    const float* wk, int32_t q_cols, int32_t norm_cols, int32_t cols, int64_t rows, void* out16, int64_t ldo, float* dwq,
                    float* dwk, float alpha, void* stream) {
  return qknorm_bwd(d_in, ld_in, static_cast<const __half*>(y16), ldy, rstd, ld_rstd, wq, wk, q_cols, norm_cols, cols, rows,
                    static_cast<__half*>(out16), ldo, dwq, dwk, alpha, S(stream));
}

int m324_head_bwd(const float* pred, const float* target, const float* u, int64_t ldu, const float* w3, int64_t rows, int32_t C, void* du16,
                  int64_t lddu, float* dw3, float* db3, float alpha, void* stream) {
  return head_bwd(pred, target, u, ldu, w3, rows, C, static_cast<__half*>(du16), lddu, dw3, db3, alpha, S(stream));
}

int m324_colsum(const void* dy16, int64_t ld, int64_t rows, int32_t cols, float* db, float alpha, void* stream) {
  return colsum(static_cast<const __half*>(dy16), ld, rows, cols, db, alpha, S(stream));
}

int m324_sum_groups(const float* in, int64_t ld_in, int32_t ngroups, int64_t group_stride, int32_t rpg, int64_t in_gstride, int64_t in_goff,
                    int64_t rows, int32_t cols, float scale, int32_t accumulate, float* out32, int64_t ldo32, void* out16, int64_t ldo16,
                    void* stream) {
  return sum_groups(in, ld_in, ngroups, group_stride, rpg, in_gstride, in_goff, rows, cols, scale, accumulate, out32, ldo32,
                    static_cast<__half*>(out16), ldo16, S(stream));
}

int m324_cast_transpose_f16(const float* src, int64_t lds, int32_t N, int32_t K, void* dst, int64_t ldo, int32_t npad, void* stream) {
  return cast_transpose_f16(src, lds, N, K, static_cast<__half*>(dst), ldo, npad, S(stream));
}

int m324_attn_dot(const void* dO, int64_t lddo, const void* O, int64_t ldo, int64_t rows, int32_t H, float* D, int64_t ldd, void* stream) {
  return attn_dot(static_cast<const __half*>(dO), lddo, static_cast<const __half*>(O), ldo, rows, H, D, ldd, S(stream));
}


int m324_add_block(const float* in, int64_t ld_in, int64_t rows, int32_t cols, float scale, int32_t accumulate, float* out, int64_t ldo,
                   void* stream) {
  return add_block(in, ld_in, rows, cols, scale, accumulate, out, ldo, S(stream));
}

int m324_track_points(const void* vertex_frames, const void* vertex_normals, int32_t is_f64, int32_t T, int64_t V, const int64_t* faces,
                      int64_t F, const int64_t* face_indices, const double* barycentric, int32_t n_samples, float* points, float* normals,
                      int32_t* err_flag, void* stream) {
  return track_points(vertex_frames, vertex_normals, is_f64, T, V, reinterpret_cast<const long*>(faces), F,
                      reinterpret_cast<const long*>(face_indices), barycentric, n_samples, points, normals, err_flag, S(stream));
}

int m324_sample_texture_colors(const double* face_uvs, int64_t F, const int64_t* face_indices, const double* barycentric, int32_t n_samples,
                               const uint8_t* texture, int32_t H, int32_t W, float* rgb, int64_t* texel_yx, int32_t* err_flag,
                               void* stream) {
  return sample_texture(face_uvs, F, reinterpret_cast<const long*>(face_indices), barycentric, n_samples, texture, H, W, rgb,
                        reinterpret_cast<long*>(texel_yx), err_flag, S(stream));
}

int m324_head3_from_partials(const float* part, int32_t groups, const float* b3, int64_t rows, float* out, const float* target, float* partials,
                             int32_t* n_partials, void* stream) {
  return head3_from_partials(part, groups, b3, rows, out, target, partials, n_partials, S(stream));
}

int m324_filter_trajectories(const float* trajs, float* out, int32_t B, int32_t T, int32_t N, int32_t mode, const double* taps_host,
                             int32_t ntaps, float mincutoff, float beta, void* stream) {
  return filter_trajectories(trajs, out, B, T, N, mode, taps_host, ntaps, mincutoff, beta, S(stream));
}

int m324_scale_by_device_scalars(float* buf, int64_t n, const float* scalar_a, const float* scalar_b, float coeff_b, void* stream) {
  return scale_by_device_scalars(buf, n, scalar_a, scalar_b, coeff_b, S(stream));
}

int m324_sample_albedo(const double* vertices, int64_t V, const int64_t* faces, int64_t F, const double* uv, const int64_t* face_indices,
                       const double* points, int32_t n_samples, const uint8_t* texture, int32_t H, int32_t W, float* rgb, int64_t* texel_yx,
                       int32_t* err_flag, void* stream) {
  return sample_albedo(vertices, V, reinterpret_cast<const long*>(faces), F, uv, reinterpret_cast<const long*>(face_indices), points,
                       n_samples, texture, H, W, rgb, reinterpret_cast<long*>(texel_yx), err_flag, S(stream));
}

int m324_attention_bwd(const m324_attn_bwd_args* a, void* stream) {
  M324_REQUIRE(a != nullptr, "m324_attention_bwd: null args");
  AttnBwdArgs t;
  t.q = static_cast<const __half*>(a->q); t.q_ld = a->q_ld; t.q_rows = a->q_rows;
  t.k = static_cast<const __half*>(a->k); t.k_ld = a->k_ld;
  t.v = static_cast<const __half*>(a->v); t.v_ld = a->v_ld; t.kv_rows = a->kv_rows;
  t.B = a->B; t.H = a->H; t.Lq = a->Lq; t.Lk = a->Lk; t.q_batch_rows = a->q_batch_rows; t.kv_batch_rows = a->kv_batch_rows; t.q_batch_div = a->q_batch_div;
  t.dO = static_cast<const __half*>(a->dO); t.do_ld = a->do_ld; t.lse = a->lse; t.lse_ld = a->lse_ld; t.D = a->D; t.d_ld = a->d_ld;
  t.dQ = a->dQ; t.dq_ld = a->dq_ld; t.dK = a->dK; t.dk_ld = a->dk_ld; t.dV = a->dV; t.dv_ld = a->dv_ld; t.scale = a->scale;
  t.tune = get_tuning(3);
  return attention_bwd(t, S(stream));
}

}
