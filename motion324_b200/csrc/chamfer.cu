// Point-cloud evaluation metrics on the GPU: bidirectional nearest-neighbour distances, Chamfer distance
// and F-score, replacing the two scipy cKDTree builds + queries per frame of the reference
// (evaluation/evaluation_pcd.py:575-588 compute_chamfer_distance, :591-609 compute_fscore; called per frame at :884-885
// on 50 000 x 50 000 sampled points).
//
// The reference works in float64 (numpy points, cKDTree), so the kernel does too: squared distance
// (dx*dx + dy*dy) + dz*dz in IEEE double without FMA contraction, exact argmin (ties -> smallest index), sqrt at the end.
// fp32 inputs (the model's pcd_moved) are widened exactly.  Brute force: one query per thread, targets staged through
// shared memory as SoA doubles and broadcast to the warp; 9 fp64-pipe instructions per (query, target) pair, so the kernel
// is bound by the fp64 pipe, not by HBM (each point is read n/256 times from L2-resident arrays of 1.2 MB).
#include "common.cuh"
#include "kernels.h"

namespace m324 {
namespace {

constexpr int NN_THREADS = 256;
constexpr int NN_TILE = 1024;   // targets per shared-memory tile: 3 x 8 KB

template <typename T>
__global__ void __launch_bounds__(NN_THREADS) nn_kernel(const T* __restrict__ p1, int n1, const T* __restrict__ p2, int n2,
                                                        double* __restrict__ dist1, int* __restrict__ idx1,
                                                        double* __restrict__ dist2, int* __restrict__ idx2) {
  // blockIdx.y = 2 * frame + direction.  direction 0: queries = points2, targets = points1 -> dist1 / idx1 [n2]
  //                                      direction 1: queries = points1, targets = points2 -> dist2 / idx2 [n1]
  __shared__ double sx[NN_TILE], sy[NN_TILE], sz[NN_TILE];
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y >> 1, dir = blockIdx.y & 1;
  const int nq = dir == 0 ? n2 : n1, nt = dir == 0 ? n1 : n2;
  if (static_cast<long>(blockIdx.x) * NN_THREADS >= nq) return;   // whole block out of range (grid.x is sized for max(n1, n2))
  const T* q = (dir == 0 ? p2 + static_cast<long>(f) * n2 * 3 : p1 + static_cast<long>(f) * n1 * 3);
  const T* t = (dir == 0 ? p1 + static_cast<long>(f) * n1 * 3 : p2 + static_cast<long>(f) * n2 * 3);
  const int qi = blockIdx.x * NN_THREADS + threadIdx.x;
  const bool active = qi < nq;
  double qx = 0.0, qy = 0.0, qz = 0.0;
  if (active) {
    qx = static_cast<double>(q[3L * qi]);
    qy = static_cast<double>(q[3L * qi + 1]);
    qz = static_cast<double>(q[3L * qi + 2]);
  }
  double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
  int best_i = -1;
  for (int t0 = 0; t0 < nt; t0 += NN_TILE) {
    const int cnt = min(NN_TILE, nt - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += NN_THREADS) {
      sx[i] = static_cast<double>(t[3L * (t0 + i)]);
      sy[i] = static_cast<double>(t[3L * (t0 + i) + 1]);
      sz[i] = static_cast<double>(t[3L * (t0 + i) + 2]);
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
      const double dx = __dsub_rn(qx, sx[i]), dy = __dsub_rn(qy, sy[i]), dz = __dsub_rn(qz, sz[i]);
      const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (d < best) {   // strict: the first (smallest-index) minimum wins
        best = d;
        best_i = t0 + i;
      }
    }
  }
  if (active) {
    double* dd = dir == 0 ? dist1 + static_cast<long>(f) * n2 : dist2 + static_cast<long>(f) * n1;
    int* ib = dir == 0 ? idx1 : idx2;
    dd[qi] = sqrt(best);
    if (ib) ib[static_cast<long>(f) * nq + qi] = best_i;
  }
}

// Per frame: out[4f..] = { mean(dist1) + mean(dist2), F-score, precision = mean(dist1 < thr), recall = mean(dist2 < thr) }.
// One block per frame, fixed reduction order (bit-reproducible).
constexpr int RED_THREADS = 512;
__global__ void __launch_bounds__(RED_THREADS) chamfer_reduce_kernel(const double* __restrict__ dist1, int n2,
                                                                      const double* __restrict__ dist2, int n1, double thr,
                                                                      double* __restrict__ out) {
  __shared__ double s_sum[2][RED_THREADS];
  __shared__ int s_cnt[2][RED_THREADS];
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.x;
  const double* d1 = dist1 + static_cast<long>(f) * n2;
  const double* d2 = dist2 + static_cast<long>(f) * n1;
  double a1 = 0.0, a2 = 0.0;
  int c1 = 0, c2 = 0;
  for (int i = threadIdx.x; i < n2; i += RED_THREADS) { const double v = d1[i]; a1 += v; c1 += v < thr; }
  for (int i = threadIdx.x; i < n1; i += RED_THREADS) { const double v = d2[i]; a2 += v; c2 += v < thr; }
  s_sum[0][threadIdx.x] = a1; s_sum[1][threadIdx.x] = a2;
  s_cnt[0][threadIdx.x] = c1; s_cnt[1][threadIdx.x] = c2;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_sum[0][threadIdx.x] += s_sum[0][threadIdx.x + o];
      s_sum[1][threadIdx.x] += s_sum[1][threadIdx.x + o];
      s_cnt[0][threadIdx.x] += s_cnt[0][threadIdx.x + o];
      s_cnt[1][threadIdx.x] += s_cnt[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double m1 = s_sum[0][0] / n2, m2 = s_sum[1][0] / n1;
    const double prec = static_cast<double>(s_cnt[0][0]) / n2, rec = static_cast<double>(s_cnt[1][0]) / n1;
    out[4 * f + 0] = m1 + m2;
    out[4 * f + 1] = (prec + rec == 0.0) ? 0.0 : 2.0 * prec * rec / (prec + rec);
    out[4 * f + 2] = prec;
    out[4 * f + 3] = rec;
  }
}

}  // namespace

int chamfer_nn(const void* p1, int n1, const void* p2, int n2, int frames, int is_f64, double* dist1, int* idx1, double* dist2,
               int* idx2, cudaStream_t stream) {
  M324_REQUIRE(p1 && p2 && dist1 && dist2, "chamfer_nn: null pointer");
  M324_REQUIRE(n1 > 0 && n2 > 0 && frames > 0, "chamfer_nn: empty point set (n1=%d n2=%d frames=%d)", n1, n2, frames);
  M324_REQUIRE(frames <= 32767, "chamfer_nn: at most 32767 frames per call");
  const int nmax = n1 > n2 ? n1 : n2;
  dim3 grid((nmax + NN_THREADS - 1) / NN_THREADS, 2 * frames);
  if (is_f64)
    M324_CUDA(launch_pdl(nn_kernel<double>, grid, dim3(NN_THREADS), 0, stream, static_cast<const double*>(p1), n1,
                         static_cast<const double*>(p2), n2, dist1, idx1, dist2, idx2));
  else
    M324_CUDA(launch_pdl(nn_kernel<float>, grid, dim3(NN_THREADS), 0, stream, static_cast<const float*>(p1), n1,
                         static_cast<const float*>(p2), n2, dist1, idx1, dist2, idx2));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int chamfer_reduce(const double* dist1, int n2, const double* dist2, int n1, int frames, double threshold, double* out,
                   cudaStream_t stream) {
  M324_REQUIRE(dist1 && dist2 && out && n1 > 0 && n2 > 0 && frames > 0, "chamfer_reduce: bad arguments");
  M324_CUDA(launch_pdl(chamfer_reduce_kernel, dim3(frames), dim3(RED_THREADS), 0, stream, dist1, n2, dist2, n1, threshold, out));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

}  // namespace m324
