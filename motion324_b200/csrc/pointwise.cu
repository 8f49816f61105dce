// HBM-bound kernels of the Motion_Latent_Model forward (no tensor cores): coalesced, vectorised, warp-shuffle /
// block reductions.  Each kernel cites the reference lines it replaces.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace m324 {

namespace {

constexpr int kMaxVec = 8;  // rows up to 8 * 128 = 1024 columns per warp-row kernel

__device__ __forceinline__ void store_split_half(__half* dst, int lo_off, float x) {
  const __half h = __float2half_rn(x);
  dst[0] = h;
  if (lo_off > 0) dst[lo_off] = __float2half_rn(x - __half2float(h));
}

__device__ __forceinline__ void store4_split_half(__half* dst, int lo_off, float4 x) {
  const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&h01);
  u.y = *reinterpret_cast<const uint32_t*>(&h23);
  *reinterpret_cast<uint2*>(dst) = u;
  if (lo_off > 0) {
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(x.x - f01.x, x.y - f01.y);
    const __half2 l23 = __floats2half2_rn(x.z - f23.x, x.w - f23.y);
    u.x = *reinterpret_cast<const uint32_t*>(&l01);
    u.y = *reinterpret_cast<const uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(dst + lo_off) = u;
  }
}

// Row LayerNorm held in registers: v[i] = float4 #(lane + 32 i) of the row.  Two-pass (mean, then centred variance).
template <bool kBias>
__device__ __forceinline__ void warp_layernorm(float4 (&v)[kMaxVec], int nvec, int lane, int C, const float* w,
                                               const float* b, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / static_cast<float>(C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(C) + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) {
      const int c = (lane + 32 * i) * 4;
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
      v[i].x = v[i].x * rstd * w4.x; v[i].y = v[i].y * rstd * w4.y;
      v[i].z = v[i].z * rstd * w4.z; v[i].w = v[i].w * rstd * w4.w;
      if (kBias) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + c));
        v[i].x += b4.x; v[i].y += b4.y; v[i].z += b4.z; v[i].w += b4.w;
      }
    }
}

// nn.LayerNorm (transformer.py:345-357,400,411; Pcd_motion.py:326,337; DINOv2 norm1/norm2/norm) -> fp16 GEMM operand.
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long ldx, const float* __restrict__ w,
                                                        const float* __restrict__ b, float eps, long rows, int C,
                                                        int src_rpg, long src_gstride, long src_goff, __half* out16,
                                                        long ldo16, int lo_off, float* out32, long ldo32) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long srow = src_rpg > 0 ? (row / src_rpg) * src_gstride + src_goff + row % src_rpg : row;
  const int nvec = C / 128;
  float4 v[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) v[i] = *reinterpret_cast<const float4*>(x + srow * ldx + (lane + 32 * i) * 4);
  if (b) warp_layernorm<true>(v, nvec, lane, C, w, b, eps);
  else warp_layernorm<false>(v, nvec, lane, C, w, b, eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) {
      const int c = (lane + 32 * i) * 4;
      if (out16) store4_split_half(out16 + row * ldo16 + c, lo_off, v[i]);
      if (out32) *reinterpret_cast<float4*>(out32 + row * ldo32 + c) = v[i];
    }
}

// PointEmbed.embed (Pcd_motion.py:177-187): proj[a*8+k] = x_a * (2^k * pi); row = [sin proj | cos proj | x | 0...] (64 wide).
__global__ void __launch_bounds__(256) point_embed_kernel(const float* __restrict__ xyz, int n, __half* out, long ldo,
                                                          int lo_off) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int pt = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pt >= n) return;
  const float x0 = xyz[pt * 3 + 0], x1 = xyz[pt * 3 + 1], x2 = xyz[pt * 3 + 2];
  const float kPi = 3.14159274101257324f;  // float32(np.pi); basis = 2^k * float32(pi) exactly
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const int c = lane + 32 * rep;
    float val = 0.f;
    if (c < 48) {
      const int j = c < 24 ? c : c - 24;
      const int a = j >> 3, k = j & 7;
      const float xa = a == 0 ? x0 : (a == 1 ? x1 : x2);
      const float proj = xa * (kPi * static_cast<float>(1 << k));
      val = c < 24 ? sinf(proj) : cosf(proj);
    } else if (c < 51) {
      val = c == 48 ? x0 : (c == 49 ? x1 : x2);
    }
    store_split_half(out + static_cast<long>(pt) * ldo + c, lo_off, val);
  }
}

// Pcd_motion.py:459 / :551-553: cat[emb, normal, rgb] -> columns [col0, col0+6) of the K-padded operand.
__global__ void __launch_bounds__(256) point_extra_kernel(const float* __restrict__ normal, const float* __restrict__ rgb,
                                                          int n, __half* out, long ldo, int col0, int kpad, int lo_off) {
  pdl_trigger();
  pdl_wait();
  // one thread per (point, column): consecutive threads write consecutive columns of a row (coalesced)
  const int w = kpad - col0;
  const long total = static_cast<long>(n) * w;
  for (long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long pt = idx / w;
    const int k = static_cast<int>(idx - pt * w);
    float val = 0.f;
    if (k < 3) val = normal[pt * 3 + k];
    else if (k < 6) val = rgb[pt * 3 + (k - 3)];
    store_split_half(out + pt * ldo + col0 + k, lo_off, val);
  }
}

// Pcd_motion.py:470-472 (permute + bilinear resize, align_corners=False) + dinov2.py:78-80 (ImageNet normalise) +
// the patch-embed im2col (Conv2d k=14 s=14 as a GEMM operand): out[f*hp*hp + py*hp + px, c*196 + iy*14 + ix].
__global__ void __launch_bounds__(256) preprocess_kernel(const float* __restrict__ video, int F, int Hin, int Win, int S,
                                                         __half* patches, long ldp, int kpad) {
  pdl_trigger();
  pdl_wait();
  const int hp = S / 14;
  const long total = static_cast<long>(F) * hp * hp * kpad;
  const float sy = static_cast<float>(Hin) / static_cast<float>(S), sx = static_cast<float>(Win) / static_cast<float>(S);
  for (long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % kpad);
    const long prow = idx / kpad;
    float val = 0.f;
    if (k < 588) {
      const int c = k / 196, rem = k - c * 196, iy = rem / 14, ix = rem - iy * 14;
      const int f = static_cast<int>(prow / (hp * hp)), pp = static_cast<int>(prow % (hp * hp));
      const int y = (pp / hp) * 14 + iy, x = (pp % hp) * 14 + ix;
      float fy = fmaxf((static_cast<float>(y) + 0.5f) * sy - 0.5f, 0.f);
      float fx = fmaxf((static_cast<float>(x) + 0.5f) * sx - 0.5f, 0.f);
      const int y0 = min(static_cast<int>(fy), Hin - 1), x0 = min(static_cast<int>(fx), Win - 1);
      const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
      const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
      const float* img = video + static_cast<long>(f) * Hin * Win * 3;
      const float v00 = img[(static_cast<long>(y0) * Win + x0) * 3 + c], v01 = img[(static_cast<long>(y0) * Win + x1) * 3 + c];
      const float v10 = img[(static_cast<long>(y1) * Win + x0) * 3 + c], v11 = img[(static_cast<long>(y1) * Win + x1) * 3 + c];
      const float hy = 1.f - ly, hx = 1.f - lx;
      const float pix = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
      const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
      const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
      val = (pix - mean) / stdv;
    }
    patches[prow * ldp + k] = __float2half_rn(val);
  }
}

// DINOv2 prepare_tokens: x[f,0] = cls + pos[0]; x[f,1+i] = patch[f*np+i] + pos[1+i]  (pos already interpolated).
__global__ void __launch_bounds__(256) dino_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                                            const float* __restrict__ pos, int F, int np, int C, float* x) {
  pdl_trigger();
  pdl_wait();
  const int c4n = C / 4;
  const long total = static_cast<long>(F) * (np + 1) * c4n;
  for (long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(idx % c4n);
    const long row = idx / c4n;
    const int tok = static_cast<int>(row % (np + 1));
    const long f = row / (np + 1);
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long>(tok) * C) + c4);
    float4 v = tok == 0 ? __ldg(reinterpret_cast<const float4*>(cls) + c4)
                        : *(reinterpret_cast<const float4*>(patch + (f * np + tok - 1) * C) + c4);
    v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
    *(reinterpret_cast<float4*>(x + row * C) + c4) = v;
  }
}

// uniform [0, 1) from (seed, element index): splitmix64 finaliser, top 24 bits
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(static_cast<unsigned>(z >> 40)) * (1.0f / 16777216.0f);
}

// DINOv2 final norm + x_norm_patchtokens (dinov2.py:99-103) + pos_embed add (Pcd_motion.py:489) + special / mesh /
// video token concat (Pcd_motion.py:495-507) + transformer_input_layernorm (Pcd_motion.py:509), one warp per token.
__global__ void __launch_bounds__(256) assemble_tokens_kernel(
    const float* __restrict__ dino_x, const float* __restrict__ dino_nw, const float* __restrict__ dino_nb, float dino_eps,
    const float* __restrict__ pos_embed, const float* __restrict__ sp0, const float* __restrict__ sprest,
    const float* __restrict__ mesh_feat, const float* __restrict__ ln_w, float ln_eps, int B, int T, int ntok, int npatch,
    int C, float* out, float drop_p, unsigned long long seed, float* pre_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int L = 4 + ntok + npatch;
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long>(B) * T * L) return;
  const int l = static_cast<int>(row % L);
  const long f = row / L;  // b*T + t
  const int t = static_cast<int>(f % T);
  const long bb = f / T;
  const int nvec = C / 128;
  float4 v[kMaxVec];
  const float* src;
  if (l < 4) src = (t == 0 ? sp0 : sprest) + static_cast<long>(l) * C;
  else if (l < 4 + ntok) src = mesh_feat + (bb * ntok + (l - 4)) * C;
  else src = dino_x + (f * (npatch + 1) + 1 + (l - 4 - ntok)) * C;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) v[i] = *reinterpret_cast<const float4*>(src + (lane + 32 * i) * 4);
  if (l >= 4 + ntok) {
    warp_layernorm<true>(v, nvec, lane, C, dino_nw, dino_nb, dino_eps);
    const float* pe = pos_embed + (static_cast<long>(t) * npatch + (l - 4 - ntok)) * C;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) {
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(pe + (lane + 32 * i) * 4));
        v[i].x += p4.x; v[i].y += p4.y; v[i].z += p4.z; v[i].w += p4.w;
      }
    if (drop_p > 0.f) {   // pos_drop (Pcd_motion.py:369-370, 490), train() only: Bernoulli keep mask from a counter hash
      const float keep = 1.0f / (1.0f - drop_p);
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i)
        if (i < nvec) {
          const unsigned long long e0 = static_cast<unsigned long long>(row) * C + (lane + 32 * i) * 4;
          v[i].x = hash_uniform(seed, e0) < drop_p ? 0.f : v[i].x * keep;
          v[i].y = hash_uniform(seed, e0 + 1) < drop_p ? 0.f : v[i].y * keep;
          v[i].z = hash_uniform(seed, e0 + 2) < drop_p ? 0.f : v[i].z * keep;
          v[i].w = hash_uniform(seed, e0 + 3) < drop_p ? 0.f : v[i].w * keep;
        }
    }
  }
  if (pre_out != nullptr) {   // training: the LayerNorm input is kept for the backward pass
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nvec) *reinterpret_cast<float4*>(pre_out + row * C + (lane + 32 * i) * 4) = v[i];
  }
  warp_layernorm<false>(v, nvec, lane, C, ln_w, nullptr, ln_eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nvec) *reinterpret_cast<float4*>(out + row * C + (lane + 32 * i) * 4) = v[i];
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  return t;  // valid in warp 0
}

// shared_mlp_output.3 (Pcd_motion.py:340,561): out[r,:] = h[r,:] . W3^T + b3 in fp32, fused with the squared-error
// partial sums of MSELossComputer (model/loss.py:59-61).  One warp per row, fixed grid -> deterministic partials.
__global__ void __launch_bounds__(256) head3_mse_kernel(const float* __restrict__ h, long ldh, const float* __restrict__ w3,
                                                        const float* __restrict__ b3, long rows, int C, float* out,
                                                        const float* __restrict__ target, float* partials, int pre_gelu) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[8];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float se = 0.f;
  for (long row = static_cast<long>(blockIdx.x) * 8 + wib; row < rows; row += static_cast<long>(gridDim.x) * 8) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      float4 x = *reinterpret_cast<const float4*>(h + row * ldh + c);
      if (pre_gelu) { x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w); }   // training: h is the pre-activation
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w3 + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w3 + C + c));
      const float4 w2 = __ldg(reinterpret_cast<const float4*>(w3 + 2 * C + c));
      a0 += x.x * w0.x + x.y * w0.y + x.z * w0.z + x.w * w0.w;
      a1 += x.x * w1.x + x.y * w1.y + x.z * w1.z + x.w * w1.w;
      a2 += x.x * w2.x + x.y * w2.y + x.z * w2.z + x.w * w2.w;
    }
    a0 = warp_sum(a0) + b3[0];
    a1 = warp_sum(a1) + b3[1];
    a2 = warp_sum(a2) + b3[2];
    if (lane == 0) {
      out[row * 3 + 0] = a0; out[row * 3 + 1] = a1; out[row * 3 + 2] = a2;
      if (target) {
        const float d0 = a0 - target[row * 3 + 0], d1 = a1 - target[row * 3 + 1], d2 = a2 - target[row * 3 + 2];
        se += d0 * d0 + d1 * d1 + d2 * d2;
      }
    }
  }
  if (partials) {
    const float t = block_sum_256(se, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
  }
}

// Second half of the fused head: out[r, c] = sum over the 64-column groups (in index order) of the GEMM epilogue's partial dot
// products + b3[c]; squared-error partial sums like head3_mse_kernel.  One thread per row: its groups are 16 * groups contiguous bytes.
__global__ void __launch_bounds__(256) head3_from_partials_kernel(const float* __restrict__ part, int groups, const float* __restrict__ b3,
                                                                  long rows, float* out, const float* __restrict__ target, float* partials) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[8];
  float se = 0.f;
  for (long row = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows; row += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4* pr = reinterpret_cast<const float4*>(part) + row * groups;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int g = 0; g < groups; ++g) {
      const float4 v = pr[g];
      a0 += v.x; a1 += v.y; a2 += v.z;
    }
    a0 += b3[0]; a1 += b3[1]; a2 += b3[2];
    out[row * 3 + 0] = a0; out[row * 3 + 1] = a1; out[row * 3 + 2] = a2;
    if (target) {
      const float d0 = a0 - target[row * 3 + 0], d1 = a1 - target[row * 3 + 1], d2 = a2 - target[row * 3 + 2];
      se += d0 * d0 + d1 * d1 + d2 * d2;
    }
  }
  if (partials) {
    const float t = block_sum_256(se, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long n,
                                                          float* partials) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[8];
  float se = 0.f;
  const long n4 = n / 4;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    se += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) {
    const float d = a[n4 * 4 + threadIdx.x] - b[n4 * 4 + threadIdx.x];
    se += d * d;
  }
  const float t = block_sum_256(se, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256) mse_finalize_kernel(const float* __restrict__ partials, int n, double count,
                                                           float weight, float* loss) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += static_cast<double>(partials[i]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mse = weight > 0.f ? static_cast<float>(sh[0] / count) : 0.f;
    loss[0] = mse;            // coord_mse_loss
    loss[1] = weight * mse;   // loss
  }
}

__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ src, long lds, int rows, int cols,
                                                       __half* dst, long ldo, int kpad, int lo_off) {
  pdl_trigger();
  pdl_wait();
  const long total = static_cast<long>(rows) * kpad;
  for (long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % kpad);
    const long r = idx / kpad;
    const float v = c < cols ? src[r * lds + c] : 0.f;
    store_split_half(dst + r * ldo + c, lo_off, v);
  }
}


// smooth_trajectories(method='combined' | 'threshold' | 'gaussian') of utils/inference_utils.py:99-145, one thread per
// vertex, single pass over time:
//   threshold : out[t] = |raw[t] - raw[t-1]| < thr ? out[t-1] : raw[t]        (raw displacement, smoothed carry)
//   gaussian  : scipy.ndimage.gaussian_filter1d(sigma, mode='nearest', truncate=4): radius R = int(4 sigma + 0.5),
//               weights exp(-k^2 / (2 sigma^2)) normalised, accumulated in fp64 like scipy, over a delay line in registers.
// trajs / out: [B, T, N, 3] fp32.  HBM-bound: 12 B read + 12 B written per vertex-frame, coalesced over vertices.
constexpr int kMaxRadius = 8;
struct SmoothWeights { double w[2 * kMaxRadius + 1]; };
__global__ void __launch_bounds__(256) smooth_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int N,
                                                     float thr, int do_thr, int do_gauss, int radius, const SmoothWeights sw) {
  pdl_trigger();
  pdl_wait();
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long>(B) * N) return;
  const int b = static_cast<int>(gid / N), n = static_cast<int>(gid % N);
  const float* src = in + (static_cast<long>(b) * T * N + n) * 3;
  float* dst = out + (static_cast<long>(b) * T * N + n) * 3;
  const long st = static_cast<long>(N) * 3;   // stride between frames
  float win[2 * kMaxRadius + 1][3];           // delay line of thresholded values, win[radius] = centre
  float prev_raw[3], prev_out[3];
  const int R = do_gauss ? radius : 0;
  // t2 runs over produced (thresholded) samples; output t = t2 - R is emitted once its right neighbours exist
  for (int t2 = 0; t2 < T + R; ++t2) {
    float cur[3];
    if (t2 < T) {
      const float x = src[t2 * st], y = src[t2 * st + 1], z = src[t2 * st + 2];
      cur[0] = x; cur[1] = y; cur[2] = z;
      if (t2 > 0 && do_thr) {
        const float dx = x - prev_raw[0], dy = y - prev_raw[1], dz = z - prev_raw[2];
        if (sqrtf(dx * dx + dy * dy + dz * dz) < thr) { cur[0] = prev_out[0]; cur[1] = prev_out[1]; cur[2] = prev_out[2]; }
      }
      prev_raw[0] = x; prev_raw[1] = y; prev_raw[2] = z;
      prev_out[0] = cur[0]; prev_out[1] = cur[1]; prev_out[2] = cur[2];
    } else {  // 'nearest' padding on the right
      cur[0] = prev_out[0]; cur[1] = prev_out[1]; cur[2] = prev_out[2];
    }
    if (!do_gauss) {
      dst[t2 * st] = cur[0]; dst[t2 * st + 1] = cur[1]; dst[t2 * st + 2] = cur[2];
      continue;
    }
    if (t2 == 0) {   // 'nearest' padding on the left: fill the whole line with the first sample
#pragma unroll
      for (int k = 0; k < 2 * kMaxRadius + 1; ++k) { win[k][0] = cur[0]; win[k][1] = cur[1]; win[k][2] = cur[2]; }
    } else {
#pragma unroll
      for (int k = 0; k < 2 * kMaxRadius; ++k) {
        if (k < 2 * R) { win[k][0] = win[k + 1][0]; win[k][1] = win[k + 1][1]; win[k][2] = win[k + 1][2]; }
      }
      win[2 * R][0] = cur[0]; win[2 * R][1] = cur[1]; win[2 * R][2] = cur[2];
    }
    const int t = t2 - R;
    if (t >= 0) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
      for (int k = 0; k < 2 * kMaxRadius + 1; ++k) {
        if (k <= 2 * R) {
          const double wk = sw.w[k];
          a0 += wk * static_cast<double>(win[k][0]); a1 += wk * static_cast<double>(win[k][1]); a2 += wk * static_cast<double>(win[k][2]);
        }
      }
      dst[t * st] = static_cast<float>(a0); dst[t * st + 1] = static_cast<float>(a1); dst[t * st + 2] = static_cast<float>(a2);
    }
  }
}

// The two other methods of utils/inference_utils.py:99-195, one thread per (vertex, channel) series along time:
//   mode 1 'savgol'  (:148-163): scipy.signal.savgol_filter(window, polyorder, mode='nearest') = a symmetric FIR whose taps
//                    the host computes (least-squares fit, motion324_b200/inference.py), fp64 accumulation like scipy.ndimage,
//                    the series padded with its end values;
//   mode 2 'oneeuro' (:58-96, 165-175): the One-Euro recurrence, in fp32 without FMA contraction, operation for operation as
//                    NumPy evaluates it on float32 scalars (Python-float constants are rounded to fp32 first: NEP 50).
struct FilterParams {
  double taps[2 * kMaxRadius + 1];
  int radius;
  float alpha_d, one_minus_alpha_d, mincutoff, beta, two_pi;
};
__global__ void __launch_bounds__(256) filter_series_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int N, int mode,
                                                            const FilterParams fp) {
  pdl_trigger();
  pdl_wait();
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long>(B) * N * 3) return;
  const int b = static_cast<int>(gid / (static_cast<long>(N) * 3));
  const long off = static_cast<long>(b) * T * N * 3 + gid % (static_cast<long>(N) * 3);
  const long st = static_cast<long>(N) * 3;
  const float* src = in + off;
  float* dst = out + off;
  if (mode == 1) {
    const int R = fp.radius;
    for (int t = 0; t < T; ++t) {
      double acc = 0.0;
      for (int k = -R; k <= R; ++k) {
        int tt = t + k;
        tt = tt < 0 ? 0 : (tt > T - 1 ? T - 1 : tt);
        acc += fp.taps[k + R] * static_cast<double>(src[tt * st]);
      }
      dst[t * st] = static_cast<float>(acc);
    }
  } else {
    float x_prev = src[0], dx_prev = 0.f;
    dst[0] = x_prev;
    for (int t = 1; t < T; ++t) {
      const float x = src[t * st];
      const float dx = __fsub_rn(x, x_prev);
      const float dx_hat = __fadd_rn(__fmul_rn(fp.alpha_d, dx), __fmul_rn(fp.one_minus_alpha_d, dx_prev));
      const float cutoff = __fadd_rn(fp.mincutoff, __fmul_rn(fp.beta, fabsf(dx_hat)));
      const float r = __fmul_rn(__fmul_rn(fp.two_pi, cutoff), 1.0f);
      const float alpha = __fdiv_rn(r, __fadd_rn(r, 1.0f));
      const float x_hat = __fadd_rn(__fmul_rn(alpha, x), __fmul_rn(__fsub_rn(1.0f, alpha), x_prev));
      dst[t * st] = x_hat;
      x_prev = x_hat;
      dx_prev = dx_hat;
    }
  }
}

inline int grid_for(long total, int block, int cap = 148 * 16) {
  long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

int layernorm(const float* x, long ldx, const float* w, const float* b, float eps, long rows, int cols, int src_rpg,
              long src_gstride, long src_goff, __half* out16, long ldo16, int lo_off, float* out32, long ldo32,
              cudaStream_t stream) {
  M324_REQUIRE(x && w && (out16 || out32), "layernorm: null pointer");
  M324_REQUIRE(cols % 128 == 0 && cols <= 128 * kMaxVec, "layernorm: cols=%d must be a multiple of 128, <= 1024", cols);
  M324_REQUIRE(ldx % 4 == 0 && (!out16 || ldo16 % 4 == 0) && (!out32 || ldo32 % 4 == 0), "layernorm: strides must be multiples of 4");
  if (rows <= 0) return M324_OK;
  M324_CUDA(launch_pdl(layernorm_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, stream, x, ldx, w, b, eps, rows, cols, src_rpg, src_gstride,
                                                                             src_goff, out16, ldo16, lo_off, out32, ldo32));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int point_embed_features(const float* xyz, int n, __half* out, long ldo, int lo_off, cudaStream_t stream) {
  M324_REQUIRE(xyz && out && ldo >= 64, "point_embed_features: bad arguments");
  if (n <= 0) return M324_OK;
  M324_CUDA(launch_pdl(point_embed_kernel, dim3((n + 7) / 8), dim3(256), 0, stream, xyz, n, out, ldo, lo_off));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int point_extra_features(const float* normal, const float* rgb, int n, __half* out, long ldo, int col0, int kpad, int lo_off,
                         cudaStream_t stream) {
  M324_REQUIRE(normal && rgb && out && kpad >= col0 + 6 && ldo >= kpad, "point_extra_features: bad arguments");
  if (n <= 0) return M324_OK;
  M324_CUDA(launch_pdl(point_extra_kernel, dim3(grid_for(static_cast<long>(n) * (kpad - col0), 256)), dim3(256), 0, stream, normal, rgb, n, out,
                       ldo, col0, kpad, lo_off));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int preprocess_frames(const float* video, int F, int Hin, int Win, int S, __half* patches, long ldp, int kpad,
                      cudaStream_t stream) {
  M324_REQUIRE(video && patches, "preprocess_frames: null pointer");
  M324_REQUIRE(S % 14 == 0 && kpad >= 588 && ldp >= kpad && Hin > 0 && Win > 0, "preprocess_frames: bad geometry");
  if (F <= 0) return M324_OK;
  const long total = static_cast<long>(F) * (S / 14) * (S / 14) * kpad;
  M324_CUDA(launch_pdl(preprocess_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, video, F, Hin, Win, S, patches, ldp, kpad));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int dino_assemble(const float* patch, const float* cls, const float* pos, int F, int np, int C, float* x, cudaStream_t stream) {
  M324_REQUIRE(patch && cls && pos && x && C % 4 == 0, "dino_assemble: bad arguments");
  if (F <= 0) return M324_OK;
  const long total = static_cast<long>(F) * (np + 1) * (C / 4);
  M324_CUDA(launch_pdl(dino_assemble_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, patch, cls, pos, F, np, C, x));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int assemble_tokens(const float* dino_x, const float* dino_nw, const float* dino_nb, float dino_eps, const float* pos_embed,
                    const float* sp0, const float* sprest, const float* mesh_feat, const float* ln_w, float ln_eps, int B,
                    int T, int ntok, int npatch, int C, float* out, float drop_p, unsigned long long seed, float* pre_out,
                    cudaStream_t stream) {
  M324_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "assemble_tokens: drop_p=%f outside [0, 1)", drop_p);
  M324_REQUIRE(dino_x && dino_nw && dino_nb && pos_embed && sp0 && sprest && mesh_feat && ln_w && out, "assemble_tokens: null pointer");
  M324_REQUIRE(C % 128 == 0 && C <= 128 * kMaxVec, "assemble_tokens: C=%d unsupported", C);
  const long rows = static_cast<long>(B) * T * (4 + ntok + npatch);
  if (rows <= 0) return M324_OK;
  M324_CUDA(launch_pdl(assemble_tokens_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, stream, 
      dino_x, dino_nw, dino_nb, dino_eps, pos_embed, sp0, sprest, mesh_feat, ln_w, ln_eps, B, T, ntok, npatch, C, out, drop_p, seed, pre_out));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int head3_mse(const float* h, long ldh, const float* w3, const float* b3, long rows, int C, float* out, const float* target,
              float* partials, int* n_partials, int pre_gelu, cudaStream_t stream) {
  M324_REQUIRE(h && w3 && b3 && out && C % 128 == 0 && ldh % 4 == 0, "head3_mse: bad arguments");
  M324_REQUIRE(!target || partials, "head3_mse: target given without a partials buffer");
  int grid = grid_for(rows, 8, 148 * 4);
  if (n_partials) *n_partials = grid;
  if (rows <= 0) return M324_OK;
  M324_CUDA(launch_pdl(head3_mse_kernel, dim3(grid), dim3(256), 0, stream, h, ldh, w3, b3, rows, C, out, target, target ? partials : nullptr, pre_gelu));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int head3_from_partials(const float* part, int groups, const float* b3, long rows, float* out, const float* target, float* partials,
                        int* n_partials, cudaStream_t stream) {
  M324_REQUIRE(part && b3 && out && groups > 0 && (reinterpret_cast<uintptr_t>(part) & 15) == 0, "head3_from_partials: bad arguments");
  M324_REQUIRE(!target || partials, "head3_from_partials: target given without a partials buffer");
  const int grid = grid_for(rows, 256, 148 * 4);
  if (n_partials) *n_partials = grid;
  if (rows <= 0) return M324_OK;
  M324_CUDA(launch_pdl(head3_from_partials_kernel, dim3(grid), dim3(256), 0, stream, part, groups, b3, rows, out, target, target ? partials : nullptr));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int mse_finalize(const float* partials, int n, double count, float weight, float* loss, cudaStream_t stream) {
  M324_REQUIRE(partials && loss && n > 0 && count > 0, "mse_finalize: bad arguments");
  M324_CUDA(launch_pdl(mse_finalize_kernel, dim3(1), dim3(256), 0, stream, partials, n, count, weight, loss));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int mse_loss(const float* pred, const float* target, long n, float weight, float* partials, float* loss, cudaStream_t stream) {
  M324_REQUIRE(pred && target && partials && loss && n > 0, "mse_loss: bad arguments");
  M324_REQUIRE((reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0, "mse_loss: pointers must be 16-byte aligned");
  const int grid = grid_for(n / 4 + 1, 256, 148 * 4);
  M324_CUDA(launch_pdl(mse_partial_kernel, dim3(grid), dim3(256), 0, stream, pred, target, n, partials));
  M324_CUDA(cudaGetLastError());
  return mse_finalize(partials, grid, static_cast<double>(n), weight, loss, stream);
}

int cast_pad_f16(const float* src, long lds, int rows, int cols, __half* dst, long ldo, int kpad, int lo_off,
                 cudaStream_t stream) {
  M324_REQUIRE(src && dst && kpad >= cols && ldo >= kpad, "cast_pad_f16: bad arguments");
  if (rows <= 0) return M324_OK;
  M324_CUDA(launch_pdl(cast_pad_kernel, dim3(grid_for(static_cast<long>(rows) * kpad, 256)), dim3(256), 0, stream, src, lds, rows, cols, dst, ldo, kpad, lo_off));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}


int smooth_trajectories(const float* trajs, float* out, int B, int T, int N, float motion_threshold, float sigma, int do_threshold,
                        int do_gaussian, cudaStream_t stream) {
  M324_REQUIRE(trajs && out && trajs != out && B > 0 && T > 0 && N > 0, "smooth_trajectories: bad arguments (in-place is not supported)");
  int radius = 0;
  SmoothWeights sw;
  for (int k = 0; k < 2 * kMaxRadius + 1; ++k) sw.w[k] = 0.0;
  if (do_gaussian) {
    M324_REQUIRE(sigma > 0.f, "smooth_trajectories: gaussian needs sigma > 0");
    radius = static_cast<int>(4.0 * static_cast<double>(sigma) + 0.5);   // scipy: int(truncate * sd + 0.5), truncate = 4
    M324_REQUIRE(radius >= 0 && radius <= kMaxRadius, "smooth_trajectories: sigma=%f gives radius %d > %d", sigma, radius, kMaxRadius);
    double sum = 0.0;
    const double s2 = static_cast<double>(sigma) * static_cast<double>(sigma);
    for (int k = -radius; k <= radius; ++k) { sw.w[k + radius] = exp(-0.5 / s2 * k * k); sum += sw.w[k + radius]; }
    for (int k = 0; k <= 2 * radius; ++k) sw.w[k] /= sum;   // passed by value: capture-safe, no device workspace
  }
  const long total = static_cast<long>(B) * N;
  M324_CUDA(launch_pdl(smooth_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, trajs, out, B, T, N, motion_threshold, do_threshold,
                                                                             do_gaussian, radius, sw));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

int filter_trajectories(const float* trajs, float* out, int B, int T, int N, int mode, const double* taps_host, int ntaps, float mincutoff,
                        float beta, cudaStream_t stream) {
  M324_REQUIRE(trajs && out && trajs != out && B > 0 && T > 0 && N > 0, "filter_trajectories: bad arguments (in-place is not supported)");
  M324_REQUIRE(mode == 1 || mode == 2, "filter_trajectories: mode must be 1 (savgol) or 2 (oneeuro)");
  FilterParams fp = {};
  if (mode == 1) {
    M324_REQUIRE(taps_host && ntaps >= 1 && (ntaps & 1) && ntaps <= 2 * kMaxRadius + 1, "filter_trajectories: savgol needs an odd tap count <= %d",
                 2 * kMaxRadius + 1);
    fp.radius = ntaps / 2;
    for (int k = 0; k < ntaps; ++k) fp.taps[k] = taps_host[k];
  } else {
    // OneEuroFilter(mincutoff, beta, dcutoff = 1.0): alpha_d = r / (r + 1), r = 2 pi * dcutoff * te in Python floats (fp64), then used
    // against float32 scalars (rounded to fp32, NEP 50); 2 * np.pi meets the float32 cutoff the same way
    const double rd = 2.0 * 3.141592653589793 * 1.0 * 1.0, ad = rd / (rd + 1.0);
    fp.alpha_d = static_cast<float>(ad);
    fp.one_minus_alpha_d = static_cast<float>(1.0 - ad);
    fp.mincutoff = mincutoff;
    fp.beta = beta;
    fp.two_pi = static_cast<float>(2.0 * 3.141592653589793);
  }
  const long total = static_cast<long>(B) * N * 3;
  M324_CUDA(launch_pdl(filter_series_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, trajs, out, B, T, N, mode, fp));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

}  // namespace m324
