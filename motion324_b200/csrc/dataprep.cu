// Data-prep gathers on the GPU: the on-disk -> tensor step in front of Motion_Latent_Model.forward.
// Replaces the per-object NumPy work of the reference's dataset (dataset/dataset_utils.py:44-136 track_with_normal_rgb,
// :19-41 sample_texture_color_vectorized): given the sampled faces and their barycentric coordinates (sampling itself is
// trimesh's RNG and stays on the host),
//   points[t, s]  = sum_c bary[s, c] * vertices[t, faces[face_idx[s], c]]                       (dataset_utils.py:112-114)
//   normals[t, s] = normalise(sum_c bary[s, c] * vertex_normals[t, faces[face_idx[s], c]])      (:116-127; zero norm -> / 1)
//   rgb[s]        = texture[y, x] / 255,  uv = sum_c bary[s, c] * face_uvs[face_idx[s], c],
//                   x = clip(int(u * (W - 1))), y = clip(int((1 - v) * (H - 1)))                  (:88-98, 33-41)
// The reference computes in float64 (NumPy) and casts to float32 at the end (:131-133); the kernels do the same, sums in
// index order without FMA contraction, so the INTEGER texel indices and the gathered bytes are bit-exact and the float
// outputs equal the float64 -> float32 rounding of the NumPy result.  HBM-bound: one thread per (frame, sample), the three
// vertex rows are 12-byte gathers (L2-resident: a frame's vertex array is a few hundred KB), outputs are coalesced.
#include "common.cuh"
#include "kernels.h"

namespace m324 {
namespace {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

template <typename TV>
__global__ void __launch_bounds__(256) track_points_kernel(const TV* __restrict__ verts, const TV* __restrict__ vnormals, int T, long V,
                                                           const long* __restrict__ faces, long F, const long* __restrict__ face_idx,
                                                           const double* __restrict__ bary, int S, float* __restrict__ points,
                                                           float* __restrict__ normals, int* __restrict__ err) {
  pdl_trigger();
  pdl_wait();
  const long total = static_cast<long>(T) * S;
  for (long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(idx / S), s = static_cast<int>(idx % S);
    const long f = face_idx[s];
    if (f < 0 || f >= F) { atomicExch(err, 1); continue; }
    long vi[3];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      vi[c] = faces[3 * f + c];
      ok = ok && vi[c] >= 0 && vi[c] < V;
    }
    if (!ok) { atomicExch(err, 2); continue; }
    const double b0 = bary[3L * s], b1 = bary[3L * s + 1], b2 = bary[3L * s + 2];
    const TV* vt = verts + static_cast<long>(t) * V * 3;
    float* po = points + idx * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // trimesh.triangles.barycentric_to_points: (triangles * barycentric[:, :, None]).sum(axis=1)
      const double v = dadd(dadd(dmul(static_cast<double>(vt[3 * vi[0] + k]), b0), dmul(static_cast<double>(vt[3 * vi[1] + k]), b1)),
                            dmul(static_cast<double>(vt[3 * vi[2] + k]), b2));
      po[k] = static_cast<float>(v);
    }
    if (vnormals != nullptr) {
      const TV* nt = vnormals + static_cast<long>(t) * V * 3;
      double n[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)   // np.einsum('ij,ijk->ik'): sum over the three corners in index order
        n[k] = dadd(dadd(dmul(b0, static_cast<double>(nt[3 * vi[0] + k])), dmul(b1, static_cast<double>(nt[3 * vi[1] + k]))),
                    dmul(b2, static_cast<double>(nt[3 * vi[2] + k])));
      double nrm = sqrt(dadd(dadd(dmul(n[0], n[0]), dmul(n[1], n[1])), dmul(n[2], n[2])));   // np.linalg.norm(axis=1)
      if (nrm == 0.0) nrm = 1.0;
      float* no = normals + idx * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) no[k] = static_cast<float>(n[k] / nrm);
    }
  }
}

__global__ void __launch_bounds__(256) sample_texture_kernel(const double* __restrict__ face_uvs, long F, const long* __restrict__ face_idx,
                                                             const double* __restrict__ bary, int S, const unsigned char* __restrict__ tex,
                                                             int H, int W, float* __restrict__ rgb, long* __restrict__ texel,
                                                             int* __restrict__ err) {
  pdl_trigger();
  pdl_wait();
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    const long f = face_idx[s];
    if (f < 0 || f >= F) { atomicExch(err, 1); continue; }
    const double b0 = bary[3L * s], b1 = bary[3L * s + 1], b2 = bary[3L * s + 2];
    const double* uv = face_uvs + 6 * f;
    const double u = dadd(dadd(dmul(b0, uv[0]), dmul(b1, uv[2])), dmul(b2, uv[4]));
    const double v = dadd(dadd(dmul(b0, uv[1]), dmul(b1, uv[3])), dmul(b2, uv[5]));
    // .astype(int): truncation toward zero of the float64 product, then np.clip
    long x = static_cast<long>(dmul(u, static_cast<double>(W - 1)));
    long y = static_cast<long>(dmul(dadd(1.0, -v), static_cast<double>(H - 1)));
    x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
    y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
    if (texel != nullptr) { texel[2L * s] = y; texel[2L * s + 1] = x; }
    const unsigned char* px = tex + (static_cast<long>(y) * W + x) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[3L * s + k] = static_cast<float>(static_cast<double>(px[k]) / 255.0);
  }
}

// utils/mesh_processing.py:130-191 sample_pointcloud_with_albedo, the per-sample Python loop (:174-182): barycentric
// coordinates of the sampled point RE-DERIVED from the point and its triangle (barycentric_coords, :107-127: dot products,
// denom == 0 -> 1/3 each), uv = sum_c w_c * (uv_c mod 1), u = int(clip(uv.x * W, 0, W - 1)), v = int(clip((1 - uv.y) * H, 0,
// H - 1)), colour = float32(texel) / float32(255) (the reference divides a float32 array).  float64 like NumPy, no FMA
// contraction, dot products summed in index order: the texel indices are bit-exact.
__device__ __forceinline__ double ddot3(const double* a, const double* b) {
  return dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2]));
}
__device__ __forceinline__ double pymod1(double x) {   // Python / NumPy float modulo by 1.0: result in [0, 1)
  double r = fmod(x, 1.0);
  if (r != 0.0 && r < 0.0) r = dadd(r, 1.0);
  return r;
}
__global__ void __launch_bounds__(256) sample_albedo_kernel(const double* __restrict__ verts, long V, const long* __restrict__ faces, long F,
                                                            const double* __restrict__ uv, const long* __restrict__ face_idx,
                                                            const double* __restrict__ points, int S, const unsigned char* __restrict__ tex,
                                                            int H, int W, float* __restrict__ rgb, long* __restrict__ texel,
                                                            int* __restrict__ err) {
  pdl_trigger();
  pdl_wait();
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    const long f = face_idx[s];
    if (f < 0 || f >= F) { atomicExch(err, 1); continue; }
    long vi[3];
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      vi[c] = faces[3 * f + c];
      ok = ok && vi[c] >= 0 && vi[c] < V;
    }
    if (!ok) { atomicExch(err, 2); continue; }
    double a[3], v0[3], v1[3], v2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      a[k] = verts[3 * vi[0] + k];
      v0[k] = dadd(verts[3 * vi[1] + k], -a[k]);
      v1[k] = dadd(verts[3 * vi[2] + k], -a[k]);
      v2[k] = dadd(points[3L * s + k], -a[k]);
    }
    const double d00 = ddot3(v0, v0), d01 = ddot3(v0, v1), d11 = ddot3(v1, v1), d20 = ddot3(v2, v0), d21 = ddot3(v2, v1);
    const double denom = dadd(dmul(d00, d11), -dmul(d01, d01));
    double w0, w1, w2;
    if (denom == 0.0) {
      w0 = w1 = w2 = 1.0 / 3.0;
    } else {
      w1 = __ddiv_rn(dadd(dmul(d11, d20), -dmul(d01, d21)), denom);
      w2 = __ddiv_rn(dadd(dmul(d00, d21), -dmul(d01, d20)), denom);
      w0 = dadd(dadd(1.0, -w1), -w2);
    }
    double us = 0.0, vs = 0.0;
    {
      const double wc[3] = {w0, w1, w2};
      double uu[3], vv[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uu[c] = dmul(wc[c], pymod1(uv[2 * vi[c]]));
        vv[c] = dmul(wc[c], pymod1(uv[2 * vi[c] + 1]));
      }
      us = dadd(dadd(uu[0], uu[1]), uu[2]);
      vs = dadd(dadd(vv[0], vv[1]), vv[2]);
    }
    double fx = dmul(us, static_cast<double>(W)), fy = dmul(dadd(1.0, -vs), static_cast<double>(H));
    fx = fx < 0.0 ? 0.0 : (fx > W - 1 ? static_cast<double>(W - 1) : fx);      // np.clip, then int(): truncation
    fy = fy < 0.0 ? 0.0 : (fy > H - 1 ? static_cast<double>(H - 1) : fy);
    const long x = static_cast<long>(fx), y = static_cast<long>(fy);
    if (texel != nullptr) { texel[2L * s] = y; texel[2L * s + 1] = x; }
    const unsigned char* px = tex + (y * W + x) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[3L * s + k] = __fdiv_rn(static_cast<float>(px[k]), 255.0f);
  }
}

int grid_1d(long n, int per_block) {
  long g = (n + per_block - 1) / per_block;
  const long cap = static_cast<long>(sm_count() > 0 ? sm_count() : 148) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int track_points(const void* verts, const void* vnormals, int f64, int T, long V, const long* faces, long F, const long* face_idx,
                 const double* bary, int S, float* points, float* normals, int* err, cudaStream_t stream) {
  M324_REQUIRE(T >= 0 && V > 0 && F > 0 && S >= 0, "track_points: bad sizes T=%d V=%ld F=%ld S=%d", T, V, F, S);
  if (static_cast<long>(T) * S == 0) return M324_OK;     // empty sample set / no frames: nothing to write
  M324_REQUIRE(verts && faces && face_idx && bary && points && err, "track_points: null pointer");
  M324_REQUIRE((vnormals == nullptr) == (normals == nullptr), "track_points: vertex normals and the normals output go together");
  const int grid = grid_1d(static_cast<long>(T) * S, 256);
  if (f64)
    M324_CUDA(launch_pdl(track_points_kernel<double>, dim3(grid), dim3(256), 0, stream, static_cast<const double*>(verts),
                         static_cast<const double*>(vnormals), T, V, faces, F, face_idx, bary, S, points, normals, err));
  else
    M324_CUDA(launch_pdl(track_points_kernel<float>, dim3(grid), dim3(256), 0, stream, static_cast<const float*>(verts),
                         static_cast<const float*>(vnormals), T, V, faces, F, face_idx, bary, S, points, normals, err));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int sample_texture(const double* face_uvs, long F, const long* face_idx, const double* bary, int S, const unsigned char* tex, int H, int W,
                   float* rgb, long* texel, int* err, cudaStream_t stream) {
  M324_REQUIRE(F > 0 && H > 0 && W > 0 && S >= 0, "sample_texture: bad sizes F=%ld H=%d W=%d S=%d", F, H, W, S);
  if (S == 0) return M324_OK;
  M324_REQUIRE(face_uvs && face_idx && bary && tex && rgb && err, "sample_texture: null pointer");
  M324_CUDA(launch_pdl(sample_texture_kernel, dim3(grid_1d(S, 256)), dim3(256), 0, stream, face_uvs, F, face_idx, bary, S, tex, H, W, rgb,
                       texel, err));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int sample_albedo(const double* verts, long V, const long* faces, long F, const double* uv, const long* face_idx, const double* points, int S,
                  const unsigned char* tex, int H, int W, float* rgb, long* texel, int* err, cudaStream_t stream) {
  M324_REQUIRE(V > 0 && F > 0 && H > 0 && W > 0 && S >= 0, "sample_albedo: bad sizes V=%ld F=%ld H=%d W=%d S=%d", V, F, H, W, S);
  if (S == 0) return M324_OK;
  M324_REQUIRE(verts && faces && uv && face_idx && points && tex && rgb && err, "sample_albedo: null pointer");
  M324_CUDA(launch_pdl(sample_albedo_kernel, dim3(grid_1d(S, 256)), dim3(256), 0, stream, verts, V, faces, F, uv, face_idx, points, S, tex, H,
                       W, rgb, texel, err));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

}  // namespace m324
