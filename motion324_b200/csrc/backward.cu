// HBM-bound kernels of the backward pass (what torch.autograd runs for Motion_Latent_Model under
// train.py:157-170).  Gradients of activations are carried in units of 1/alpha (alpha = dLoss * 2 w / n, so the seed is
// pred - target): fp16 between GEMMs, fp32 on the residual stream; parameter gradients are accumulated into fp32 buffers
// as alpha * (...), i.e. in true units, with atomics after a per-block reduction.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace m324 {
namespace {

constexpr int kMaxVec = 8;   // rows up to 1024 columns (float4 x 32 lanes x 8)

__device__ __forceinline__ void store4_half(__half* dst, float4 x) {
  const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&h01);
  u.y = *reinterpret_cast<const uint32_t*>(&h23);
  *reinterpret_cast<uint2*>(dst) = u;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward (autograd of transformer.py:345-357,400,411 and Pcd_motion.py:326,337).  One warp per row, rows
// strided over the grid so that every thread keeps a private partial of dgamma / dbeta for its 4-column slices:
//   xhat = (x - mean) * rstd;  g = dy * w;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat))  [+ dres]
//   dgamma += alpha * sum_rows dy * xhat;  dbeta += alpha * sum_rows dy
// mean / rstd are recomputed from the saved x (two-pass, in registers).  Rows of x (and of dx) may be the gathered token
// slice of the forward (src_rpg mapping).  dx32 may alias dres (residual-stream gradient updated in place).
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, long lddy, const float* __restrict__ x,
                                                            long ldx, const float* __restrict__ w, float eps, long rows, int C,
                                                            int src_rpg, long src_gstride, long src_goff,
                                                            const float* dres, long lddres, float* dx32, long lddx32,
                                                            __half* dx16, long lddx16, float* dgamma, float* dbeta, float alpha) {
  __shared__ float red[8][32 * 4 + 4];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nvec = C / 128;
  float4 pg[kMaxVec], pb[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) pg[i] = pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float invC = 1.0f / static_cast<float>(C);
  for (long row = static_cast<long>(blockIdx.x) * 8 + wib; row < rows; row += static_cast<long>(gridDim.x) * 8) {
    const long srow = src_rpg > 0 ? (row / src_rpg) * src_gstride + src_goff + row % src_rpg : row;
    float4 xv[kMaxVec], gv[kMaxVec], rv[kMaxVec];
    float s = 0.f;
    // all loads of the row are issued before the first reduction: the kernel is latency-bound otherwise (one row of one
    // operand = 3 KB in flight per warp)
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        xv[i] = *reinterpret_cast<const float4*>(x + srow * ldx + (lane + 32 * i) * 4);
        gv[i] = *reinterpret_cast<const float4*>(dy + row * lddy + (lane + 32 * i) * 4);
        if (dres) rv[i] = *reinterpret_cast<const float4*>(dres + srow * lddres + (lane + 32 * i) * 4);
      }
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    const float mean = warp_sum(s) * invC;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
        ss += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
      }
    const float rstd = rsqrtf(warp_sum(ss) * invC + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        const int c = (lane + 32 * i) * 4;
        const float4 d4 = gv[i];
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
        xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;          // xhat
        pg[i].x = fmaf(d4.x, xv[i].x, pg[i].x); pg[i].y = fmaf(d4.y, xv[i].y, pg[i].y);
        pg[i].z = fmaf(d4.z, xv[i].z, pg[i].z); pg[i].w = fmaf(d4.w, xv[i].w, pg[i].w);
        pb[i].x += d4.x; pb[i].y += d4.y; pb[i].z += d4.z; pb[i].w += d4.w;
        gv[i] = make_float4(d4.x * w4.x, d4.y * w4.y, d4.z * w4.z, d4.w * w4.w);
        c1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
        c2 += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
      }
    c1 = warp_sum(c1) * invC;
    c2 = warp_sum(c2) * invC;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        const int c = (lane + 32 * i) * 4;
        float4 o;
        o.x = rstd * (gv[i].x - c1 - xv[i].x * c2); o.y = rstd * (gv[i].y - c1 - xv[i].y * c2);
        o.z = rstd * (gv[i].z - c1 - xv[i].z * c2); o.w = rstd * (gv[i].w - c1 - xv[i].w * c2);
        if (dres) {
          const float4 r4 = rv[i];
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        if (dx32) *reinterpret_cast<float4*>(dx32 + srow * lddx32 + c) = o;
        if (dx16) store4_half(dx16 + srow * lddx16 + c, o);
      }
  }
  // block reduction of the per-thread partials, one 128-column slice at a time, then one atomic per column and block
  for (int pass = 0; pass < (dbeta ? 2 : 1); ++pass) {
    float* dst = pass == 0 ? dgamma : dbeta;
    if (dst == nullptr) continue;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        const float4 v = pass == 0 ? pg[i] : pb[i];
        __syncthreads();
        red[wib][lane * 4 + 0] = v.x; red[wib][lane * 4 + 1] = v.y; red[wib][lane * 4 + 2] = v.z; red[wib][lane * 4 + 3] = v.w;
        __syncthreads();
        if (threadIdx.x < 128) {
          float t = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
          atomicAdd(dst + i * 128 + threadIdx.x, alpha * t);
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-head RMSNorm backward (autograd of transformer.py:30-42 as applied at :130-132 / :205-207) + fp32 -> fp16 conversion
// of the attention gradients.  d_in: fp32 [rows, ld_in] (dQ | dK | dV as produced by the attention backward);
// y16: the forward's NORMALISED q / k (fp16) -> xhat = y / w;  rstd: [rows, ld_rstd] saved by the GEMM epilogue.
//   d_raw = rstd * (w * dy - xhat * mean_64(w * dy * xhat));   dw += alpha * sum dy * xhat
// Columns [norm_cols, cols) (the V part) are only converted.  One warp per row, 2 columns per lane and head.
__global__ void __launch_bounds__(256) qknorm_bwd_kernel(const float* __restrict__ d_in, long ld_in, const __half* __restrict__ y16,
                                                         long ldy, const float* __restrict__ rstd, long ld_rstd,
                                                         const float* __restrict__ wq, const float* __restrict__ wk, int q_cols,
                                                         int norm_cols, int cols, long rows, __half* out16, long ldo,
                                                         float* dwq, float* dwk, float alpha) {
  __shared__ float red[8][64];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float2 pq = make_float2(0.f, 0.f), pk = make_float2(0.f, 0.f);
  const float2 wq2 = wq ? make_float2(wq[2 * lane], wq[2 * lane + 1]) : make_float2(1.f, 1.f);
  const float2 wk2 = wk ? make_float2(wk[2 * lane], wk[2 * lane + 1]) : make_float2(1.f, 1.f);
  auto inv = [](float w) { return fabsf(w) > 1e-20f ? 1.0f / w : 0.f; };
  const float2 iwq2 = make_float2(inv(wq2.x), inv(wq2.y)), iwk2 = make_float2(inv(wk2.x), inv(wk2.y));
  for (long row = static_cast<long>(blockIdx.x) * 8 + wib; row < rows; row += static_cast<long>(gridDim.x) * 8) {
#pragma unroll 6   // 12 / 24 / 36 head groups per row: six independent load + warp-reduce chains in flight
    for (int c0 = 0; c0 < cols; c0 += 64) {
      const int c = c0 + 2 * lane;
      const float2 d2 = *reinterpret_cast<const float2*>(d_in + row * ld_in + c);
      float2 o = d2;
      if (c0 < norm_cols) {
        const bool is_q = c0 < q_cols;
        const float2 w2 = is_q ? wq2 : wk2;
        const float2 y2 = __half22float2(*reinterpret_cast<const __half2*>(y16 + row * ldy + c));
        const float2 iw2 = is_q ? iwq2 : iwk2;
        const float2 xh = make_float2(y2.x * iw2.x, y2.y * iw2.y);
        const float r = rstd[row * ld_rstd + (c0 >> 6)];
        const float2 g = make_float2(d2.x * w2.x, d2.y * w2.y);
        const float m = warp_sum(g.x * xh.x + g.y * xh.y) * (1.0f / 64.0f);
        o.x = r * (g.x - xh.x * m);
        o.y = r * (g.y - xh.y * m);
        if (is_q) { pq.x = fmaf(d2.x, xh.x, pq.x); pq.y = fmaf(d2.y, xh.y, pq.y); }
        else { pk.x = fmaf(d2.x, xh.x, pk.x); pk.y = fmaf(d2.y, xh.y, pk.y); }
      }
      *reinterpret_cast<__half2*>(out16 + row * ldo + c) = __floats2half2_rn(o.x, o.y);
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dwq : dwk;
    if (dst == nullptr) continue;
    const float2 v = pass == 0 ? pq : pk;
    __syncthreads();
    red[wib][2 * lane] = v.x;
    red[wib][2 * lane + 1] = v.y;
    __syncthreads();
    if (threadIdx.x < 64) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
      atomicAdd(dst + threadIdx.x, alpha * t);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Output head + MSE backward (autograd of Pcd_motion.py:340,561 and model/loss.py:59-61), one warp per row:
//   e = pred - target (the gradient seed in units of 1/alpha);  h = gelu(u);  du = (e . W3) * gelu'(u)   -> fp16
//   dW3 += alpha * sum_rows e (x) h;   db3 += alpha * sum_rows e
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                       const float* __restrict__ u, long ldu, const float* __restrict__ w3, long rows,
                                                       int C, __half* du16, long lddu, float* dw3, float* db3, float alpha) {
  __shared__ float red[8][128];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nvec = C / 128;
  float4 pw[3][kMaxVec];
  float pbias[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) pw[o][i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long row = static_cast<long>(blockIdx.x) * 8 + wib; row < rows; row += static_cast<long>(gridDim.x) * 8) {
    float e[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      e[o] = pred[row * 3 + o] - target[row * 3 + o];
      pbias[o] += e[o];
    }
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        const int c = (lane + 32 * i) * 4;
        const float4 u4 = *reinterpret_cast<const float4*>(u + row * ldu + c);
        const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
        float dh[4] = {0.f, 0.f, 0.f, 0.f}, hh[4];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(w3 + o * C + c));
          dh[0] = fmaf(e[o], w4.x, dh[0]); dh[1] = fmaf(e[o], w4.y, dh[1]);
          dh[2] = fmaf(e[o], w4.z, dh[2]); dh[3] = fmaf(e[o], w4.w, dh[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          hh[k] = gelu_erf(uu[k]);
          dh[k] *= gelu_grad(uu[k]);
        }
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          pw[o][i].x = fmaf(e[o], hh[0], pw[o][i].x); pw[o][i].y = fmaf(e[o], hh[1], pw[o][i].y);
          pw[o][i].z = fmaf(e[o], hh[2], pw[o][i].z); pw[o][i].w = fmaf(e[o], hh[3], pw[o][i].w);
        }
        store4_half(du16 + row * lddu + c, make_float4(dh[0], dh[1], dh[2], dh[3]));
      }
  }
#pragma unroll
  for (int o = 0; o < 3; ++o)
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        __syncthreads();
        red[wib][lane * 4 + 0] = pw[o][i].x; red[wib][lane * 4 + 1] = pw[o][i].y;
        red[wib][lane * 4 + 2] = pw[o][i].z; red[wib][lane * 4 + 3] = pw[o][i].w;
        __syncthreads();
        if (threadIdx.x < 128) {
          float t = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
          atomicAdd(dw3 + o * C + i * 128 + threadIdx.x, alpha * t);
        }
      }
  // bias: every lane of a warp saw the same e -> lane 0 of each warp contributes
  __syncthreads();
  if (lane == 0) { red[wib][0] = pbias[0]; red[wib][1] = pbias[1]; red[wib][2] = pbias[2]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(db3 + threadIdx.x, alpha * t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Bias gradients: db[c] += alpha * sum_rows dy16[row, c].  Block = 64 rows x 128 columns tile walk.
__global__ void __launch_bounds__(256) colsum_kernel(const __half* __restrict__ dy, long ld, long rows, int cols, float* db, float alpha) {
  __shared__ float red[8][64];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + 2 * lane;
  float2 acc = make_float2(0.f, 0.f);
  if (c < cols)
    for (long row = static_cast<long>(blockIdx.y) * 8 + wib; row < rows; row += static_cast<long>(gridDim.y) * 8) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(dy + row * ld + c));
      acc.x += v.x;
      acc.y += v.y;
    }
  red[wib][2 * lane] = acc.x;
  red[wib][2 * lane + 1] = acc.y;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(db + blockIdx.x * 64 + threadIdx.x, alpha * t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// out[g_out(r), c] (+)= scale * sum_{k < ngroups} in[k * group_stride + r_in(r), c]: sums a gradient over the frames that
// shared one operand in the forward (the decoder's per-point feature and query, Pcd_motion.py:539-560; the mesh tokens
// and special tokens broadcast to every frame, :495-507).  Row mapping: r_in = (r / rpg) * in_gstride + in_goff + r % rpg
// when rpg > 0, else r.  fp32 in, fp32 and / or fp16 out.
__global__ void __launch_bounds__(256) sum_groups_kernel(const float* __restrict__ in, long ld_in, int ngroups, long group_stride,
                                                         int rpg, long in_gstride, long in_goff, long rows, int cols, float scale,
                                                         int accumulate, float* out32, long ldo32, __half* out16, long ldo16) {
  pdl_trigger();
  pdl_wait();
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c4 = cols / 4;
  if (idx >= rows * c4) return;
  const long r = idx / c4;
  const int c = static_cast<int>(idx % c4) * 4;
  const long rin = rpg > 0 ? (r / rpg) * in_gstride + in_goff + r % rpg : r;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < ngroups; ++k) {
    const float4 v = *reinterpret_cast<const float4*>(in + (k * group_stride + rin) * ld_in + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
  if (out32) {
    float4* dst = reinterpret_cast<float4*>(out32 + r * ldo32 + c);
    if (accumulate) { const float4 p = *dst; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
    *dst = acc;
  }
  if (out16) store4_half(out16 + r * ldo16 + c, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 [N, K] -> fp16 [K, ldo] transposed (the dgrad operand W^T of a weight), zero padding of columns [N, npad).
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ src, long lds, int N, int K, __half* dst, long ldo,
                                                             int npad) {
  __shared__ float tile[32][33];
  pdl_trigger();
  pdl_wait();
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ty + 8 * j, k = k0 + tx;
    tile[ty + 8 * j][tx] = (n < N && k < K) ? src[static_cast<long>(n) * lds + k] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + ty + 8 * j, n = n0 + tx;
    if (k < K && n < npad) dst[static_cast<long>(k) * ldo + n] = __float2half_rn(tile[tx][ty + 8 * j]);
  }
}

// D[row, h] = sum_d dO[row, 64 h + d] * O[row, 64 h + d]  (the softmax-backward row term), fp16 in, fp32 out.
__global__ void __launch_bounds__(256) attn_dot_kernel(const __half* __restrict__ dO, long lddo, const __half* __restrict__ O, long ldo,
                                                       long rows, int H, float* D, long ldd) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  for (int h = 0; h < H; ++h) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(dO + row * lddo + h * 64 + 2 * lane));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(O + row * ldo + h * 64 + 2 * lane));
    const float s = warp_sum(a.x * b.x + a.y * b.y);
    if (lane == 0) D[row * ldd + h] = s;
  }
}

// out[r, c] (+)= scale * in[r, c] for c < cols, arbitrary row strides and column counts (scalar accesses): moves a weight
// gradient computed at the K-padded operand width (832 / 64 columns) into the parameter's own [768, 774] / [768, 51] layout,
// and rescales a gradient buffer in place (in == out, accumulate = 0).
__global__ void __launch_bounds__(256) add_block_kernel(const float* in, long ld_in, long rows, int cols, float scale, int accumulate,
                                                        float* out, long ldo) {
  pdl_trigger();
  pdl_wait();
  const long total = rows * cols;
  for (long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = idx / cols;
    const int c = static_cast<int>(idx % cols);
    const float v = scale * in[r * ld_in + c];
    float* dst = out + r * ldo + c;
    *dst = accumulate ? *dst + v : v;
  }
}

// buf[i] *= (*sa + cb * *sb): the upstream gradient of the autograd seam (train.py:159-166: loss / grad_accum_steps, GradScaler)
// read from DEVICE memory, so that loss.backward() needs no device -> host synchronisation.
__global__ void __launch_bounds__(256) scale_dev_kernel(float* buf, long n, const float* sa, const float* sb, float cb) {
  pdl_trigger();
  pdl_wait();
  const float s = (sa != nullptr ? *sa : 0.f) + (sb != nullptr ? cb * *sb : 0.f);
  float4* b4 = reinterpret_cast<float4*>(buf);
  const long n4 = n >> 2;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 v = b4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    b4[i] = v;
  }
  for (long i = (n4 << 2) + static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) buf[i] *= s;
}

int grid_for_rows(long rows) {
  long g = (rows + 7) / 8;
  const long cap = static_cast<long>(sm_count() > 0 ? sm_count() : 148) * 8;
  return static_cast<int>(g < cap ? g : cap);
}

}  // namespace

int layernorm_bwd(const float* dy, long lddy, const float* x, long ldx, const float* w, float eps, long rows, int cols, int src_rpg,
                  long src_gstride, long src_goff, const float* dres, long lddres, float* dx32, long lddx32, __half* dx16,
                  long lddx16, float* dgamma, float* dbeta, float alpha, cudaStream_t stream) {
  M324_REQUIRE(dy && x && w && dgamma && rows > 0, "layernorm_bwd: null pointer / empty");
  M324_REQUIRE(cols % 128 == 0 && cols <= 128 * kMaxVec, "layernorm_bwd: cols=%d must be a multiple of 128, <= %d", cols, 128 * kMaxVec);
  M324_REQUIRE(lddy % 4 == 0 && ldx % 4 == 0 && (!dres || lddres % 4 == 0) && (!dx32 || lddx32 % 4 == 0) && (!dx16 || lddx16 % 4 == 0),
               "layernorm_bwd: row strides must be multiples of 4");
  M324_CUDA(launch_pdl(layernorm_bwd_kernel, dim3(grid_for_rows(rows)), dim3(256), 0, stream, dy, lddy, x, ldx, w, eps, rows, cols,
                       src_rpg, src_gstride, src_goff, dres, lddres, dx32, lddx32, dx16, lddx16, dgamma, dbeta, alpha));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int qknorm_bwd(const float* d_in, long ld_in, const __half* y16, long ldy, const float* rstd, long ld_rstd, const float* wq,
               const float* wk, int q_cols, int norm_cols, int cols, long rows, __half* out16, long ldo, float* dwq, float* dwk,
               float alpha, cudaStream_t stream) {
  M324_REQUIRE(d_in && out16 && rows > 0 && cols > 0 && cols % 64 == 0 && norm_cols % 64 == 0 && q_cols % 64 == 0 && norm_cols <= cols,
               "qknorm_bwd: bad shape (cols=%d norm_cols=%d q_cols=%d)", cols, norm_cols, q_cols);
  M324_REQUIRE(norm_cols == 0 || (y16 && rstd), "qknorm_bwd: normalised columns need y16 and rstd");
  M324_REQUIRE(ld_in % 2 == 0 && ldo % 2 == 0 && ldy % 2 == 0, "qknorm_bwd: row strides must be even");
  M324_CUDA(launch_pdl(qknorm_bwd_kernel, dim3(grid_for_rows(rows)), dim3(256), 0, stream, d_in, ld_in, y16, ldy, rstd, ld_rstd, wq, wk,
                       q_cols, norm_cols, cols, rows, out16, ldo, dwq, dwk, alpha));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int head_bwd(const float* pred, const float* target, const float* u, long ldu, const float* w3, long rows, int C, __half* du16,
             long lddu, float* dw3, float* db3, float alpha, cudaStream_t stream) {
  M324_REQUIRE(pred && target && u && w3 && du16 && dw3 && db3 && rows > 0, "head_bwd: null pointer / empty");
  M324_REQUIRE(C % 128 == 0 && C <= 128 * kMaxVec && ldu % 4 == 0 && lddu % 4 == 0, "head_bwd: C=%d must be a multiple of 128", C);
  M324_CUDA(launch_pdl(head_bwd_kernel, dim3(grid_for_rows(rows)), dim3(256), 0, stream, pred, target, u, ldu, w3, rows, C, du16, lddu,
                       dw3, db3, alpha));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int colsum(const __half* dy, long ld, long rows, int cols, float* db, float alpha, cudaStream_t stream) {
  M324_REQUIRE(dy && db && rows > 0 && cols > 0 && cols % 2 == 0 && ld % 2 == 0, "colsum: bad arguments");
  long gy = (rows + 63) / 64;
  if (gy > 256) gy = 256;
  M324_CUDA(launch_pdl(colsum_kernel, dim3((cols + 63) / 64, static_cast<unsigned>(gy)), dim3(256), 0, stream, dy, ld, rows, cols, db, alpha));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int sum_groups(const float* in, long ld_in, int ngroups, long group_stride, int rpg, long in_gstride, long in_goff, long rows,
               int cols, float scale, int accumulate, float* out32, long ldo32, __half* out16, long ldo16, cudaStream_t stream) {
  M324_REQUIRE(in && (out32 || out16) && ngroups > 0 && rows > 0 && cols > 0 && cols % 4 == 0 && ld_in % 4 == 0, "sum_groups: bad arguments");
  const long total = rows * (cols / 4);
  M324_CUDA(launch_pdl(sum_groups_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, in, ld_in, ngroups,
                       group_stride, rpg, in_gstride, in_goff, rows, cols, scale, accumulate, out32, ldo32, out16, ldo16));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int cast_transpose_f16(const float* src, long lds, int N, int K, __half* dst, long ldo, int npad, cudaStream_t stream) {
  M324_REQUIRE(src && dst && N > 0 && K > 0 && npad >= N && ldo >= npad, "cast_transpose_f16: bad arguments");
  M324_CUDA(launch_pdl(cast_transpose_kernel, dim3((npad + 31) / 32, (K + 31) / 32), dim3(256), 0, stream, src, lds, N, K, dst, ldo, npad));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int add_block(const float* in, long ld_in, long rows, int cols, float scale, int accumulate, float* out, long ldo, cudaStream_t stream) {
  M324_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ldo >= cols, "add_block: bad arguments");
  long g = (rows * cols + 255) / 256;
  const long cap = static_cast<long>(sm_count() > 0 ? sm_count() : 148) * 16;
  if (g > cap) g = cap;
  M324_CUDA(launch_pdl(add_block_kernel, dim3(static_cast<unsigned>(g)), dim3(256), 0, stream, in, ld_in, rows, cols, scale, accumulate, out, ldo));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int scale_by_device_scalars(float* buf, long n, const float* sa, const float* sb, float cb, cudaStream_t stream) {
  M324_REQUIRE(buf && n > 0 && (sa || sb) && (reinterpret_cast<uintptr_t>(buf) & 15) == 0, "scale_by_device_scalars: bad arguments");
  long g = (n / 4 + 255) / 256;
  const long cap = static_cast<long>(sm_count() > 0 ? sm_count() : 148) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  M324_CUDA(launch_pdl(scale_dev_kernel, dim3(static_cast<unsigned>(g)), dim3(256), 0, stream, buf, n, sa, sb, cb));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int attn_dot(const __half* dO, long lddo, const __half* O, long ldo, long rows, int H, float* D, long ldd, cudaStream_t stream) {
  M324_REQUIRE(dO && O && D && rows > 0 && H > 0 && lddo % 2 == 0 && ldo % 2 == 0, "attn_dot: bad arguments");
  M324_CUDA(launch_pdl(attn_dot_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, stream, dO, lddo, O, ldo, rows, H, D, ldd));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

}  // namespace m324
