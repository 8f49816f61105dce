// tcgen05 GEMM with fused epilogues:  C[M,N] = epilogue( A[M,K] . W[N,K]^T )
//
// Replaces every nn.Linear on the Motion_Latent_Model forward path (reference call sites:
// model/transformer.py:124-126,200,217,73-78; model/Pcd_motion.py:186,459,551-553,561; the DINOv2 ViT linears
// behind model/image_encoder/dinov2.py:99).  Operands are fp16 (or bf16), K-major, staged by TMA into 128B-swizzled
// shared memory; accumulation is fp32 in TMEM (two accumulator stages, so the epilogue of tile i overlaps the
// mainloop of tile i+1); persistent grid, one CTA per SM, warp-specialised:
//   warp 0   : TMA producer (one elected thread)       warp 1 : tcgen05.mma issuer (one elected thread)
//   warp 2   : TMEM allocator                          warp 3 : idle
//   warps 4+ : epilogue (8 warps; warp w reads TMEM lane quarter w%4 and column half (w-4)/4)
//
// Epilogue (all optional, fused, fp32 math):  per-head RMS q/k-norm (transformer.py:30-42,206-207) -> +bias ->
// GELU(erf) -> *gamma (DINOv2 LayerScale) -> +fp32 residual (row index optionally modulo, for the decoder where the
// residual is the per-point feature shared by all frames) -> fp32 and/or fp16 stores (fp16 optionally as a hi|lo split
// pair).  Stores/residual loads are transposed through shared memory so every warp instruction touches 4 full 128 B
// lines (the TMEM register layout is one row per thread, which would otherwise scatter 32 lines per instruction).
//
// Split precision: passes == 3 runs A_hi.W_hi + A_lo.W_hi + A_hi.W_lo into the same accumulator (A and W stored as
// [rows, hi | lo]); ~21-bit operand mantissa at 3x the tensor work, used only on the three small GEMMs that dominate the
// output error (DESIGN.md "Precision").
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace m324 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = (4 + EPI_WARPS) * 32;

// Warp-level timeline of one CTA pair of gemm2_kernel, profiling builds only (-DM324_TIMELINE=1, scripts/gemm_timeline.py):
// when the MMA warp waits for a free accumulator stage, when the epilogue warps wait for a full one and how long a tile's
// epilogue takes, against the main loop.  Compiled out by default.
#if defined(M324_TIMELINE) && M324_TIMELINE
__device__ unsigned long long* g_tlg_buf = nullptr;
__device__ int g_tlg_cluster = -1, g_tlg_cap = 0;
enum : int { TLG_LOAD_TILE = 1, TLG_ACC_WAIT = 8, TLG_ACC_OK, TLG_MMA_ISSUED, TLG_EPI_WAIT = 16, TLG_EPI_START, TLG_EPI_END };
__device__ __forceinline__ void tlg_record(int ev, int cluster, unsigned rank) {
  if (cluster != g_tlg_cluster || (threadIdx.x & 31) != 0 || g_tlg_buf == nullptr) return;
  const unsigned long long i = atomicAdd(g_tlg_buf, 1ull);
  if (i + 1 < static_cast<unsigned long long>(g_tlg_cap))
    g_tlg_buf[i + 1] = (static_cast<unsigned long long>(clock64()) << 16) | (static_cast<unsigned long long>((threadIdx.x >> 5) + 16 * rank) << 8) | ev;
}
#define TLG(ev) tlg_record(ev, cluster_id, rank)
#else
#define TLG(ev) ((void)0)
#endif

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * 32 * 4;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES + STG_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;  // + barriers + alignment slack
  static constexpr int TMEM_COLS = 2 * BN;                 // 512 or 256: power of two
};

// Epilogue of one warp's share of a 128-row accumulator tile: rows [row0, row0+32) (its TMEM lane quarter), columns
// [n_col0 + chalf*BN/2, +BN/2) in 64-column groups.  tmem_acc = TMEM address of (lane quarter, column 0 of the tile).
//
// All math happens in the TMEM register layout (one output row per thread): q/k RMS-norm, +bias, GELU, *gamma.  Results
// leave through a 4 KB 128B-swizzled staging tile per warp and the TMA unit:
//   fp32  : 32x32 boxes, plain store, or -- when the output IS the residual (in-place residual stream update) -- a
//           cp.reduce.async.bulk .add, so the residual is never loaded by the SM at all;
//   fp16  : 32x64 boxes (optionally a second box with the fp16 remainder for hi|lo split operands).
// A residual that is not the output (the decoder's per-point feature, indexed modulo) takes the transposed LDG path.
struct EpiFlags {
  bool qk, inplace, generic_resid;
};

// Residual prefetch of the 2-CTA kernel's full epilogue: a residual that is NOT the output (the decoder's per-point feature, indexed
// modulo the point count, Pcd_motion.py:556-560) arrives as 32 x 32 fp32 boxes by TMA, one box ahead of its use, in a second 128B-
// swizzled staging tile per warp and column half; the thread adds its own row from there.  (The fallback for row mappings a box cannot
// follow is the transposed LDG path below.)
struct ResidPF {
  float* rstg;              // this warp's two 4 KB boxes (column halves 0 / 1 of a 64-column group)
  uint64_t* rbar;           // one mbarrier per box
  const CUtensorMap* tmR;
  uint32_t phase;           // bit h: parity of the next arrival on rbar[h]
  bool on;
};
__device__ __forceinline__ long resid_row(const GemmArgs& p, long r) {
  return p.resid_mod > 0 ? (p.resid_div > 0 ? (r / p.resid_div) * p.resid_mod : 0) + r % p.resid_mod : r;
}
__device__ __forceinline__ void resid_prefetch(const GemmArgs& p, ResidPF& pf, int half, int col, long row) {
  float* dst = pf.rstg + half * 1024;
  fence_proxy_async_smem();      // the generic-proxy reads of the previous box (ordered by the __syncwarp before this call) precede the TMA write
  mbar_expect_tx(&pf.rbar[half], 4096);
  tma_load_2d(dst, pf.tmR, &pf.rbar[half], col, static_cast<int>(resid_row(p, row)));
}

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& p, const CUtensorMap* tmO32, const CUtensorMap* tmO16,
                                              float* stg, uint32_t tmem_acc, long row0, int n_col0, int chalf, int lane,
                                              const EpiFlags f, ResidPF* pf = nullptr, long next_row0 = 0, int next_n_col0 = 0,
                                              bool has_next = false) {
  if (p.force_bn128 & 16) return;  // profiling aid: mainloop only (outputs are NOT written)
  uint8_t* stg8 = reinterpret_cast<uint8_t*>(stg);
  for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 64) {
    const int gcol0 = n_col0 + c0;
    if (gcol0 >= p.N) break;
    float v[64];
    {
      uint32_t* vu = reinterpret_cast<uint32_t*>(v);
      tmem_ld_32x32b_x32(tmem_acc + c0, vu);
      tmem_ld_32x32b_x32(tmem_acc + c0 + 32, vu + 32);
      tmem_ld_wait();
    }
    if (EPI >= 1 && f.qk && gcol0 < 2 * p.qk_cols) {
      // per-head RMSNorm over this 64-column group (one head), row = this thread
      const float* w = gcol0 < p.qk_cols ? p.qn_w : p.kn_w;
      if (w != nullptr) {
        float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          ss0 = fmaf(v[j], v[j], ss0); ss1 = fmaf(v[j + 1], v[j + 1], ss1);
          ss2 = fmaf(v[j + 2], v[j + 2], ss2); ss3 = fmaf(v[j + 3], v[j + 3], ss3);
        }
        const float r = rsqrtf(((ss0 + ss1) + (ss2 + ss3)) * (1.0f / 64.0f) + p.qk_eps);
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = v[j] * r * __ldg(w + j);
        if (p.qk_rstd != nullptr && row0 + lane < p.M) p.qk_rstd[(row0 + lane) * p.ld_rstd + (gcol0 >> 6)] = r;
      }
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] += __ldg(p.bias + gcol0 + j);
    }
    if (EPI == 2 && p.aux_mode == 1) {
      // training forward: keep the pre-activation (fp16) for the backward pass; one full 128 B line per thread
      if (row0 + lane < p.M) {
        uint4* dst = reinterpret_cast<uint4*>(p.aux16 + (row0 + lane) * p.ldaux + gcol0);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          dst[c] = make_uint4(pack_half2(v[8 * c], v[8 * c + 1]), pack_half2(v[8 * c + 2], v[8 * c + 3]),
                              pack_half2(v[8 * c + 4], v[8 * c + 5]), pack_half2(v[8 * c + 6], v[8 * c + 7]));
      }
    } else if (EPI == 2 && p.aux_mode == 2) {
      // backward through GELU: dU = dG * gelu'(U), U = the saved pre-activation
      if (row0 + lane < p.M) {
        const uint4* src = reinterpret_cast<const uint4*>(p.aux16 + (row0 + lane) * p.ldaux + gcol0);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = __ldg(src + c);
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 uf = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
            v[8 * c + 2 * e] *= gelu_grad(uf.x);
            v[8 * c + 2 * e + 1] *= gelu_grad(uf.y);
          }
        }
      }
    }
    if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < 64; j += 2) gelu_poly2(v[j], v[j + 1]);
    }
    if (p.gamma) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] *= __ldg(p.gamma + gcol0 + j);
    }
    if (p.out_scale != 0.f) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] *= p.out_scale;
    }

    if (EPI == 2 && p.head_w != nullptr) {
      // shared_mlp_output.3 folded into the epilogue of shared_mlp_output.1 (Pcd_motion.py:340, 561): this thread's row, this 64-column
      // group: three fp32 partial dot products -> head_part[row, group]; head3_from_partials adds the groups in order (+ bias, MSE)
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      const float4* w0 = reinterpret_cast<const float4*>(p.head_w + gcol0);
      const float4* w1 = reinterpret_cast<const float4*>(p.head_w + p.N + gcol0);
      const float4* w2 = reinterpret_cast<const float4*>(p.head_w + 2 * p.N + gcol0);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 a = __ldg(w0 + j), b = __ldg(w1 + j), c = __ldg(w2 + j);
        d0 = fmaf(v[4 * j], a.x, fmaf(v[4 * j + 1], a.y, fmaf(v[4 * j + 2], a.z, fmaf(v[4 * j + 3], a.w, d0))));
        d1 = fmaf(v[4 * j], b.x, fmaf(v[4 * j + 1], b.y, fmaf(v[4 * j + 2], b.z, fmaf(v[4 * j + 3], b.w, d1))));
        d2 = fmaf(v[4 * j], c.x, fmaf(v[4 * j + 1], c.y, fmaf(v[4 * j + 2], c.z, fmaf(v[4 * j + 3], c.w, d2))));
      }
      if (row0 + lane < p.M)
        *reinterpret_cast<float4*>(p.head_part + ((row0 + lane) * (p.N >> 6) + (gcol0 >> 6)) * 4) = make_float4(d0, d1, d2, 0.f);
    }
    if (p.out32) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int gcol = gcol0 + half * 32;
        if (gcol >= p.N) break;
        bool resid_added = false;
        if (EPI == 2 && pf != nullptr && pf->on) {
          // the residual box of this (64-column group, half) was requested one box ago: add this thread's row, then request the next one
          mbar_wait(&pf->rbar[half], (pf->phase >> half) & 1u);
          pf->phase ^= 1u << half;
          const float* rs = pf->rstg + half * 1024;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = *reinterpret_cast<const float4*>(rs + lane * 32 + ((j ^ (lane & 7)) << 2));
            v[half * 32 + 4 * j] += r.x; v[half * 32 + 4 * j + 1] += r.y; v[half * 32 + 4 * j + 2] += r.z; v[half * 32 + 4 * j + 3] += r.w;
          }
          __syncwarp();
          if (lane == 0) {
            if (c0 + 64 < (chalf + 1) * (BN / 2)) resid_prefetch(p, *pf, half, gcol0 + 64 + half * 32, row0);
            else if (has_next) resid_prefetch(p, *pf, half, next_n_col0 + chalf * (BN / 2) + half * 32, next_row0);
          }
          resid_added = true;
        }
        if (lane == 0) tma_store_wait_read();   // previous box has been read out of the staging tile
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = make_float4(v[half * 32 + 4 * j], v[half * 32 + 4 * j + 1], v[half * 32 + 4 * j + 2],
                                       v[half * 32 + 4 * j + 3]);
          *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = x;
        }
        if (EPI < 2 || !f.generic_resid || resid_added) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (f.inplace) tma_reduce_add_2d(tmO32, stg, gcol, static_cast<int>(row0));
            else tma_store_2d(tmO32, stg, gcol, static_cast<int>(row0));
            tma_store_commit();
          }
        } else {
          // out = acc + resid[(row / div) * mod + row % mod]: transposed, coalesced; all 8 residual loads in flight first
          __syncwarp();
          const int c4 = lane & 7;
          float4 rr[8];
          long grow[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            grow[it] = row0 + it * 4 + (lane >> 3);
            long r2 = grow[it];
            if (p.resid_mod > 0) r2 = (p.resid_div > 0 ? (r2 / p.resid_div) * p.resid_mod : 0) + r2 % p.resid_mod;
            rr[it] = grow[it] < p.M ? *reinterpret_cast<const float4*>(p.resid + r2 * p.ldr + gcol + c4 * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rl = it * 4 + (lane >> 3);
            float4 x = *reinterpret_cast<const float4*>(stg + rl * 32 + ((c4 ^ (rl & 7)) << 2));
            x.x += rr[it].x; x.y += rr[it].y; x.z += rr[it].z; x.w += rr[it].w;
            if (grow[it] < p.M) *reinterpret_cast<float4*>(p.out32 + grow[it] * p.ldo32 + gcol + c4 * 4) = x;
          }
          __syncwarp();
        }
      }
    }
    if (p.out16) {
      uint32_t h[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = p.out16_bf16 ? pack_bf16x2(v[2 * j], v[2 * j + 1]) : pack_half2(v[2 * j], v[2 * j + 1]);
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(stg8 + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmO16, stg, gcol0, static_cast<int>(row0));
        tma_store_commit();
      }
      if (EPI == 2 && p.out16_lo_off > 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
          h[j] = pack_half2(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(stg8 + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmO16, stg, p.out16_lo_off + gcol0, static_cast<int>(row0));
          tma_store_commit();
        }
      }
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16, const GemmArgs p) {
  using C = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* staging = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int num_tiles = num_m * num_n * ksplit;          // tile = (m_blk, n_blk, k split); k split fastest
  const int kpb = (p.K + BK - 1) / BK;                   // K is a multiple of BK except in tn mode (TMA zero-fills the tail)
  const int kchunk = (kpb + ksplit - 1) / ksplit;        // k blocks per split (host guarantees every split is non-empty)

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ks = tile % ksplit, tmn = tile / ksplit;
        const int m_blk = tmn / num_n, n_blk = tmn % num_n;
        const int kb0 = ks * kchunk, kb1 = min(kpb, kb0 + kchunk);
        const int nkb = (kb1 - kb0) * p.passes;
        for (int kb = 0; kb < nkb; ++kb) {
          const int pass = kb / (kb1 - kb0), kk = kb0 + kb - pass * (kb1 - kb0);
          const int a_col = kk * BK + (pass == 1 ? p.a_lo_off : 0);
          const int w_col = kk * BK + (pass == 2 ? p.w_lo_off : 0);
          mbar_wait(&empty[stage], phase ^ 1);
          if ((p.force_bn128 & 32) && (tile != static_cast<int>(blockIdx.x) || kb >= C::STAGES)) {
            mbar_arrive(&full[stage]);   // profiling aid: no loads after the first ring fill (results are garbage)
          } else {
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          uint8_t* sa = tiles + stage * C::STAGE_BYTES;
          if (p.tn) {
            // MN-major operands: boxes of 64 (M or N, contiguous) x 64 (k rows); MN atom i of a tile at + i * 8 KB
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * 8192, &tmA, &full[stage], m_blk * BM + i * 64, kk * BK);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(sa + C::A_BYTES + i * 8192, &tmW, &full[stage], n_blk * BN + i * 64, kk * BK);
          } else {
          tma_load_2d(sa, &tmA, &full[stage], a_col, m_blk * BM);
          tma_load_2d(sa + C::A_BYTES, &tmW, &full[stage], w_col, n_blk * BN);
          }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the pipeline (keeps addresses / descriptors warp-uniform); one elected lane issues.
    const bool tn = p.tn != 0;
    const uint32_t idesc = umma_idesc_f16_ex(BM, BN, p.bf16 != 0, p.bf16 != 0, tn, tn);
    // constant descriptor fields; the start address is added per stage.  K-major: LBO unused; MN-major: LBO = 8 KB atom stride
    const uint64_t desc_hi = tn ? umma_desc_sw128(0, 8192, 1024) : umma_desc_sw128(0, 16, 1024);
    const int kstep = tn ? (2048 >> 4) : 2;   // 16 k-rows of 128 B, or 32 B along a K-major row
    const uint32_t tiles_addr = smem_u32(tiles);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ks = tile % ksplit;
      const int nkb = (min(kpb, ks * kchunk + kchunk) - ks * kchunk) * p.passes;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = desc_hi | static_cast<uint64_t>(((tiles_addr + stage * C::STAGE_BYTES) & 0x3FFFF) >> 4);
          const uint64_t db = da + (C::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16_ss(d_tmem, da + kstep * k, db + kstep * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (kb == nkb - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int chalf = ew >> 2;     // column half of the tile
    float* stg = staging + ew * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    EpiFlags ef;
    ef.qk = p.qn_w != nullptr;
    ef.inplace = p.accumulate != 0 || (p.resid != nullptr && p.resid == p.out32 && p.ldr == p.ldo32 && p.resid_mod == 0);
    ef.generic_resid = p.resid != nullptr && !ef.inplace;
    if (lane == 0 && warp == 4) { tma_prefetch_desc(&tmO32); tma_prefetch_desc(&tmO16); }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tmn = tile / ksplit;
      const int m_blk = tmn / num_n, n_blk = tmn % num_n;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN, EPI>(p, &tmO32, &tmO16, stg, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN,
                        static_cast<long>(m_blk) * BM + q * 32, n_blk * BN, chalf, lane, ef);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2-CTA variant: a cluster of two CTAs (one SM pair) computes a 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA
// stages its own 128 rows of A and its own BN/2 rows of W (half the L2 -> SMEM traffic and half the SMEM operand reads per
// SM of the 1-CTA kernel, which is L2-bandwidth bound at 128x256: (M+N)/(M*N) bytes per flop), holds the 128 x BN fp32
// accumulator of its rows in its own TMEM and runs the same epilogue.  Only the leader (even) CTA issues MMAs; its
// commits are multicast to both CTAs' barriers; both CTAs' TMA bytes are credited to the leader's full barriers.
template <int BN, int EPI>
struct Gemm2Cfg {
  // the full epilogue (EPI 2) trades two ring stages for the residual-prefetch boxes (2 x 4 KB per epilogue warp)
  static constexpr int STAGES = BN == 256 ? (EPI == 2 ? 4 : 6) : (EPI == 2 ? 5 : 8);
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * 32 * 4;
  static constexpr int RSTG_BYTES = EPI == 2 ? EPI_WARPS * 2 * 32 * 32 * 4 : 0;
  static constexpr int RSTG_OFF = STAGES * STAGE_BYTES + STG_BYTES;
  static constexpr int BAR_OFF = RSTG_OFF + RSTG_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 512 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
             const __grid_constant__ CUtensorMap tmR, const GemmArgs p) {
  using C = Gemm2Cfg<BN, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* staging = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* full = bars;                      // used in the leader CTA only
  uint64_t* empty = bars + C::STAGES;         // per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * C::STAGES;     // per CTA (multicast commit)
  uint64_t* tempty = bars + 2 * C::STAGES + 2;  // leader only: 2 * EPI_WARPS arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
  uint64_t* rbars = bars + 2 * C::STAGES + 6;   // EPI 2: 2 per epilogue warp (residual prefetch boxes)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int kpb = p.K / BK;
  const int nkb = kpb * p.passes;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * EPI_WARPS);
    }
    if (EPI == 2) {
      for (int s = 0; s < 2 * EPI_WARPS; ++s) mbar_init(&rbars[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation of BOTH CTAs visible before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        TLG(TLG_LOAD_TILE);
        for (int kb = 0; kb < nkb; ++kb) {
          const int pass = kb / kpb, kk = kb - pass * kpb;
          const int a_col = kk * BK + (pass == 1 ? p.a_lo_off : 0);
          const int w_col = kk * BK + (pass == 2 ? p.w_lo_off : 0);
          mbar_wait(&empty[stage], phase ^ 1);
          if ((p.force_bn128 & 32) && (tile != cluster_id || kb >= C::STAGES)) {
            if (rank == 0) mbar_arrive(&full[stage]);   // profiling aid: no loads after the first ring fill
          } else {
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
          uint8_t* sa = tiles + stage * C::STAGE_BYTES;
          tma_load_2d_2sm(sa, &tmA, &full[stage], a_col, m_blk * 2 * BM + static_cast<int>(rank) * BM);
          tma_load_2d_2sm(sa + C::A_BYTES, &tmW, &full[stage], w_col, n_blk * BN + static_cast<int>(rank) * (BN / 2));
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16_ex(2 * BM, BN, p.bf16 != 0, p.bf16 != 0, false, false);
      const uint64_t desc_hi = umma_desc_sw128(0, 16, 1024);
      const uint32_t tiles_addr = smem_u32(tiles);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        TLG(TLG_ACC_WAIT);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        TLG(TLG_ACC_OK);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = desc_hi | static_cast<uint64_t>(((tiles_addr + stage * C::STAGE_BYTES) & 0x3FFFF) >> 4);
            const uint64_t db = da + (C::A_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty[stage], 3);
            if (kb == nkb - 1) umma_commit_2sm(&tfull[acc], 3);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        TLG(TLG_MMA_ISSUED);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int q = warp & 3;
    const int chalf = ew >> 2;
    float* stg = staging + ew * (32 * 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    EpiFlags ef;
    ef.qk = p.qn_w != nullptr;
    ef.inplace = p.accumulate != 0 || (p.resid != nullptr && p.resid == p.out32 && p.ldr == p.ldo32 && p.resid_mod == 0);
    ef.generic_resid = p.resid != nullptr && !ef.inplace;
    if (lane == 0 && warp == 4) { tma_prefetch_desc(&tmO32); tma_prefetch_desc(&tmO16); }
    ResidPF pf;
    pf.on = false;
    if (EPI == 2) {
      pf.rstg = reinterpret_cast<float*>(smem + C::RSTG_OFF) + ew * 2048;
      pf.rbar = rbars + 2 * ew;
      pf.tmR = &tmR;
      pf.phase = 0;
      pf.on = p.fast_resid != 0 && ef.generic_resid;
    }
    auto tile_row0 = [&](int t) { return static_cast<long>(t / num_n) * 2 * BM + static_cast<long>(rank) * BM + q * 32; };
    if (EPI == 2 && pf.on && cluster_id < num_tiles) {
      if (lane == 0) {
        tma_prefetch_desc(&tmR);
        const int col = (cluster_id % num_n) * BN + chalf * (BN / 2);
        resid_prefetch(p, pf, 0, col, tile_row0(cluster_id));
        resid_prefetch(p, pf, 1, col + 32, tile_row0(cluster_id));
      }
      __syncwarp();
    }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int nxt = tile + num_clusters;
      TLG(TLG_EPI_WAIT);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      TLG(TLG_EPI_START);
      epilogue_tile<BN, EPI>(p, &tmO32, &tmO16, stg, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN,
                        static_cast<long>(m_blk) * 2 * BM + static_cast<long>(rank) * BM + q * 32, n_blk * BN, chalf, lane, ef,
                        EPI == 2 ? &pf : nullptr, nxt < num_tiles ? tile_row0(nxt) : 0, nxt < num_tiles ? (nxt % num_n) * BN : 0,
                        nxt < num_tiles);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[acc]);
      TLG(TLG_EPI_END);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();   // the peer's smem / barriers / TMEM must stay alive until both CTAs are done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int EPI>
int launch2(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO32,
            const CUtensorMap& tmO16, const CUtensorMap& tmR, cudaStream_t stream) {
  using C = Gemm2Cfg<BN, EPI>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    M324_CUDA(cudaFuncSetAttribute(gemm2_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured.mark();
  }
  const int num_tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * ((a.N + BN - 1) / BN);
  int clusters = sm_count() / 2;
  if (clusters <= 0) clusters = 74;
  if (num_tiles < clusters) clusters = num_tiles;
  M324_CUDA(launch_pdl(gemm2_kernel<BN, EPI>, dim3(2 * clusters), dim3(GEMM_THREADS), C::SMEM_BYTES, stream, tmA, tmW, tmO32, tmO16, tmR, a));
  return M324_OK;
}

template <int BN, int EPI>
int launch(const GemmArgs& a, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO32,
           const CUtensorMap& tmO16, const CUtensorMap&, cudaStream_t stream) {
  using C = GemmCfg<BN>;
  static PerDeviceOnce configured;
  if (configured.need()) {
    M324_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured.mark();
  }
  const int num_tiles = ((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN) * (a.ksplit > 1 ? a.ksplit : 1);
  int grid = sm_count();
  if (grid <= 0) grid = 148;
  if (num_tiles < grid) grid = num_tiles;
  M324_CUDA(launch_pdl(gemm_kernel<BN, EPI>, dim3(grid), dim3(GEMM_THREADS), C::SMEM_BYTES, stream, tmA, tmW, tmO32, tmO16, a));
  return M324_OK;
}

}  // namespace

int gemm(const GemmArgs& a_in, cudaStream_t stream) {
  GemmArgs a = a_in;
  M324_REQUIRE(a.A && a.W, "gemm: null operand");
  M324_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  M324_REQUIRE(a.tn || a.K % BK == 0, "gemm: K=%d must be a multiple of %d (pad the operand)", a.K, BK);
  M324_REQUIRE(!a.tn || (a.passes == 1 && a.M % 8 == 0), "gemm: tn mode needs passes == 1 and M %% 8 == 0");
  M324_REQUIRE(a.ksplit <= 1 || (a.accumulate && a.passes == 1 && a.out32 && !a.out16 && !a.resid && !a.bias && a.act == 0 && !a.qn_w &&
                                 !a.gamma && a.aux_mode == 0),
               "gemm: split-K needs accumulate = 1 into an fp32 output and a linear epilogue");
  M324_REQUIRE(!a.accumulate || (a.out32 && !a.resid), "gemm: accumulate needs an fp32 output and no residual");
  M324_REQUIRE(a.aux_mode == 0 || (a.aux16 && a.ldaux % 8 == 0 && a.N % 64 == 0 && (reinterpret_cast<uintptr_t>(a.aux16) & 15) == 0),
               "gemm: aux tensor must be 16-byte aligned with ldaux %% 8 == 0 and N %% 64 == 0");
  if (a.ksplit > 1) {   // make every split non-empty
    const int kpb = (a.K + BK - 1) / BK;
    const int kchunk = (kpb + a.ksplit - 1) / a.ksplit;
    a.ksplit = (kpb + kchunk - 1) / kchunk;
  }
  M324_REQUIRE(a.N % 32 == 0 && (!a.out16 || a.N % 64 == 0), "gemm: N=%d must be a multiple of 32 (64 with an fp16 output)", a.N);
  M324_REQUIRE(a.passes == 1 || a.passes == 3, "gemm: passes must be 1 or 3");
  M324_REQUIRE(a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 elements");
  M324_REQUIRE(a.out32 || a.out16 || a.head_part, "gemm: no output");
  M324_REQUIRE((a.head_w == nullptr) == (a.head_part == nullptr), "gemm: head_w and head_part go together");
  M324_REQUIRE(!a.head_w || (a.N % 64 == 0 && !a.tn && a.ksplit <= 1 && (reinterpret_cast<uintptr_t>(a.head_w) & 15) == 0 &&
                             (reinterpret_cast<uintptr_t>(a.head_part) & 15) == 0),
               "gemm: the fused head needs N %% 64 == 0 and 16-byte aligned head_w / head_part");
  M324_REQUIRE(!a.out32 || (a.ldo32 % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out32) & 15) == 0), "gemm: out32 must be 16-byte aligned with ldo32 %% 4 == 0");
  M324_REQUIRE(!a.out16 || (a.ldo16 % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out16) & 15) == 0), "gemm: out16 must be 16-byte aligned with ldo16 %% 8 == 0");
  M324_REQUIRE(!a.out16 || a.out16_lo_off % 8 == 0, "gemm: out16_lo_off must be a multiple of 8");
  M324_REQUIRE(!a.resid || a.ldr % 4 == 0, "gemm: ldr must be a multiple of 4");
  if (a.qn_w) {
    M324_REQUIRE(a.qk_cols % 64 == 0 && a.N % 64 == 0, "gemm: q/k-norm needs 64-col heads");
  }
  const int ka = a.passes == 3 ? a.a_lo_off + a.K : a.K;
  const int kw = a.passes == 3 ? a.w_lo_off + a.K : a.K;
  M324_REQUIRE(a.tn ? (a.M <= a.lda && a.N <= a.ldw) : (ka <= a.lda && kw <= a.ldw), "gemm: operand row shorter than its extent (lda=%ld ldw=%ld)", a.lda, a.ldw);
  const int mode = a.force_bn128 & 15;
  // mode: 0 auto, 1 = 1-CTA 128x128, 2 = 1-CTA 128x256, 3 = 2-CTA 256x128, 4 = 2-CTA 256x256 (BN 256 needs N % 256 == 0)
  bool two_cta, bn256;
  if (mode != 0) {
    two_cta = mode >= 3;
    bn256 = (a.N % 256 == 0) && mode != 1 && mode != 3;
  } else {
    // auto: CTA pair on 256 x 256 (256 x 128 if N is not a multiple of 256); single CTA for M <= 128.  The four shapes were
    // measured at the model's sizes (scripts/gemm_modes.py): they are within a few percent of each other, wave
    // quantisation included, so there is no shape heuristic.
    two_cta = a.M > 128;
    bn256 = a.N % 256 == 0;
    if (two_cta && bn256 && get_tuning_knob(4) == 1) {   // opt-in: measured SLOWER in the step (11.8 vs 11.45 ms, scripts/step_ab.py, profiles/r2f_step_ab.txt)
      // wave quantisation: a persistent grid of C CTA pairs runs ceil(tiles / C) rounds; 256 x 128 tiles (a few percent less
      // efficient per tile: twice the A traffic per flop) win when they fill the last round much better.  DINOv2's 8224 rows:
      // N = 768 -> 99 tiles = 2 rounds at 67 % with BN = 256, 198 tiles = 3 rounds at 89 % with BN = 128.
      int clusters = sm_count() / 2;
      if (clusters <= 0) clusters = 74;
      const long mt = (a.M + 2 * BM - 1) / (2 * BM);
      const long t256 = mt * (a.N / 256), t128 = mt * (a.N / 128);
      auto eff = [&](long t) { return static_cast<double>(t) / static_cast<double>(((t + clusters - 1) / clusters) * clusters); };
      if (t256 >= clusters && eff(t128) * 0.94 > eff(t256)) bn256 = false;
    }
  }
  if (a.tn || a.ksplit > 1) two_cta = false;   // MN-major operands and split-K live in the 1-CTA kernel
  CUtensorMap tmA, tmW;
  if (a.tn) {
    uint32_t box[2] = {64, BK};
    uint64_t dimsA[2] = {static_cast<uint64_t>(a.M), static_cast<uint64_t>(a.K)};
    uint64_t strA[1] = {static_cast<uint64_t>(a.lda) * 2};
    int e = make_tmap_16b(&tmA, a.A, 2, dimsA, strA, box);
    if (e) return e;
    uint64_t dimsW[2] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.K)};
    uint64_t strW[1] = {static_cast<uint64_t>(a.ldw) * 2};
    e = make_tmap_16b(&tmW, a.W, 2, dimsW, strW, box);
    if (e) return e;
  } else {
  {
    uint64_t dims[2] = {static_cast<uint64_t>(ka), static_cast<uint64_t>(a.M)};
    uint64_t str[1] = {static_cast<uint64_t>(a.lda) * 2};
    uint32_t box[2] = {BK, BM};
    int e = make_tmap_16b(&tmA, a.A, 2, dims, str, box);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(kw), static_cast<uint64_t>(a.N)};
    uint64_t str[1] = {static_cast<uint64_t>(a.ldw) * 2};
    const uint32_t bn = bn256 ? 256 : 128;
    uint32_t box[2] = {BK, two_cta ? bn / 2 : bn};
    int e = make_tmap_16b(&tmW, a.W, 2, dims, str, box);
    if (e) return e;
  }
  }
  CUtensorMap tmO32, tmO16;
  memset(&tmO32, 0, sizeof(tmO32));
  memset(&tmO16, 0, sizeof(tmO16));
  if (a.out32) {
    uint64_t dims[2] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M)};
    uint64_t str[1] = {static_cast<uint64_t>(a.ldo32) * 4};
    uint32_t box[2] = {32, 32};
    int e = make_tmap_f32(&tmO32, a.out32, 2, dims, str, box);
    if (e) return e;
  }
  if (a.out16) {
    uint64_t dims[2] = {static_cast<uint64_t>(a.out16_lo_off > 0 ? a.out16_lo_off + a.N : a.N), static_cast<uint64_t>(a.M)};
    uint64_t str[1] = {static_cast<uint64_t>(a.ldo16) * 2};
    uint32_t box[2] = {64, 32};
    int e = make_tmap_16b(&tmO16, a.out16, 2, dims, str, box);
    if (e) return e;
  }
  // Three epilogue instantiations (register budget: the 8 epilogue warps share the 168-register cap with the TMA / MMA warps):
  //   0 lean : bias / GELU / LayerScale / scale -> fp16 and / or fp32 store or in-place reduce-add (trunk, DINOv2, decoder MLP GEMMs)
  //   1 qk   : lean + per-head RMS q/k-norm (+ reciprocal RMS for the backward): every to_qkv / to_q / to_kv projection
  //   2 full : + modulo residual, hi|lo split output, training aux tensor
  const bool inplace = a.accumulate != 0 || (a.resid != nullptr && a.resid == a.out32 && a.ldr == a.ldo32 && a.resid_mod == 0);
  const bool plain = a.aux_mode == 0 && (a.resid == nullptr || inplace) && a.out16_lo_off == 0 && a.head_w == nullptr;
  const int epi = !plain ? 2 : (a.qn_w ? 1 : 0);
  // residual prefetch by TMA (2-CTA kernel, full epilogue): needs whole 32-row boxes to map to 32 consecutive residual rows
  CUtensorMap tmR;
  memset(&tmR, 0, sizeof(tmR));
  const int bn_sel = bn256 ? 256 : 128;
  a.fast_resid = 0;
  if (two_cta && epi == 2 && a.resid != nullptr && !inplace && a.out32 != nullptr && a.out16 == nullptr && a.N % bn_sel == 0 && a.N % 64 == 0 &&
      (a.resid_mod == 0 || a.resid_mod % 32 == 0) && (a.resid_div == 0 || a.resid_div % 32 == 0) && a.ldr % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(a.resid) & 15) == 0 && get_tuning_knob(5) != 1) {
    const long rrows = a.resid_mod > 0 ? (a.resid_div > 0 ? ((a.M - 1) / a.resid_div + 1) * static_cast<long>(a.resid_mod) : a.resid_mod) : a.M;
    uint64_t dims[2] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(rrows)};
    uint64_t str[1] = {static_cast<uint64_t>(a.ldr) * 4};
    uint32_t box[2] = {32, 32};
    const int e = make_tmap_f32(&tmR, a.resid, 2, dims, str, box);
    if (e) return e;
    a.fast_resid = 1;
  }
#define M324_GEMM_DISPATCH(FN)                                                                                                    \
  switch (epi) {                                                                                                                    \
    case 0: return bn256 ? FN<256, 0>(a, tmA, tmW, tmO32, tmO16, tmR, stream) : FN<128, 0>(a, tmA, tmW, tmO32, tmO16, tmR, stream);   \
    case 1: return bn256 ? FN<256, 1>(a, tmA, tmW, tmO32, tmO16, tmR, stream) : FN<128, 1>(a, tmA, tmW, tmO32, tmO16, tmR, stream);   \
    default: return bn256 ? FN<256, 2>(a, tmA, tmW, tmO32, tmO16, tmR, stream) : FN<128, 2>(a, tmA, tmW, tmO32, tmO16, tmR, stream);  \
  }
  if (two_cta) { M324_GEMM_DISPATCH(launch2) }
  M324_GEMM_DISPATCH(launch)
#undef M324_GEMM_DISPATCH
}

#if defined(M324_TIMELINE) && M324_TIMELINE
// Profiling builds only: timeline buffer (word 0 = record counter) and the CTA pair (cluster index) of gemm2_kernel to trace.
extern "C" int m324_timeline_set_gemm(void* buf, int cap, int cluster) {
  unsigned long long* b = static_cast<unsigned long long*>(buf);
  if (cudaMemcpyToSymbol(g_tlg_buf, &b, sizeof(b)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_tlg_cap, &cap, sizeof(cap)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_tlg_cluster, &cluster, sizeof(cluster)) != cudaSuccess) return -1;
  return 0;
}
#endif

}  // namespace m324
