// tcgen05 flash-attention forward, head dim 64, non-causal, no bias:  O = softmax(Q K^T * scale) V
//
// Replaces xformers.ops.memory_efficient_attention as called by the reference (model/transformer.py:134-139,
// 209-214; layout [B, L, H, Dh], attn_bias=None, p=0) and the attention inside the DINOv2 ViT blocks.
//
// One work item = one (batch, head) x 256 query rows (two 128-row Q tiles), walking 128-row K/V tiles:
//   warp 0      : TMA producer  (Q once; K_j / V_j through a 3-stage ring)
//   warps 1, 2  : tcgen05.mma issuers, one per Q tile.   S^q = Q^q K_j^T  (128x128x64, fp32 in TMEM, operands in smem)
//                                                         O^q += P^q V_j   (128x64x128, fp32 in TMEM, A = P^q FROM TMEM)
//   warp 3      : idle
//   warps 4-7   : softmax for Q tile 0 (one thread per query row)      warps 8-11 : softmax for Q tile 1
// The two Q tiles ping-pong on the tensor pipe (while one tile is in softmax the other's MMAs run).  Softmax is online
// with a LAZY rescale: the running max only moves (and O in TMEM is only rescaled) when it grows by more than 2^8,
// so P <= 256 fits fp16 and the O read-modify-write is rare.  P is written as packed fp16 pairs into TENSOR memory
// (tcgen05.st) and is the A operand of the second MMA straight from there (kPTmem; the shared-memory variant is kept as a
// compile option): P never crosses the shared-memory port, which at 128 B/clk was as binding as the MUFU unit.  V is consumed
// MN-major straight from its TMA tile (no transpose).  TMEM: S 2 x 128 | O 2 x 64 | P 2 x 64 columns = 512.
// The softmax denominators are fp32 adds in the softmax threads.  The exponential phases of the two softmax warps of an SM
// sub-partition take turns on the MUFU unit (named barriers, handed over two chunks early), see softmax_tile.
// Normalisation by the row sum happens once, in the epilogue.  Q/K/V/P are fp16, all statistics fp32.
//
// Work decomposition (attention_plan, host side, CPU-tested through m324_attention_plan): the grid is linear over work items;
//   * tail split  : the items of a partly filled last wave are cut into 2-8 K/V ranges (one CTA each) whose (O, m, l) go to a
//                   caller-lent workspace and are combined by attn_merge_kernel;
//   * frame loop  : launches with shared queries and one K/V tile per batch (the decoder) walk 8 batches per CTA through the
//                   K/V ring with Q resident and an epilogue per tile;
//   * partial     : a caller may run one attention as several launches over disjoint K/V ranges (+ m324_attention_merge).
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace m324 {

namespace {

constexpr int ATT_THREADS = 384;  // warpgroup 0: TMA / MMA / 2 idle warps; warpgroups 1,2: softmax
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 fp16
constexpr int ONES_BYTES = 2048;          // 16 x 64 fp16 of 1.0: the (constant) B operand of the row-sum MMA
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_P = OFF_V + KV_STAGES * TILE_BYTES;
constexpr int OFF_ONES = OFF_P + 4 * TILE_BYTES;
constexpr int OFF_BAR = OFF_ONES + ONES_BYTES;
constexpr int ATT_SMEM = OFF_BAR + 256 + 1024;
constexpr uint32_t TM_S = 0;     // S^0 at cols [0,128), S^1 at [128,256)
constexpr uint32_t TM_O = 256;   // O^0 at cols [256,320), O^1 at [320,384)
constexpr uint32_t TM_L = 384;   // row sums l^0 at cols [384,400), l^1 at [400,416): P . ones, same MMA stream as P . V
constexpr float LOG2E = 1.4426950408889634f;

// ---------------------------------------------------------------------------------------------------------------
// Warp-level timeline (profiling builds only: scripts/build_variant.sh out.so -DM324_TIMELINE=1, read by
// scripts/attn_timeline.py).  ncu's sampler cannot say WHEN a warp waits for which barrier; with this flag lane 0 of every
// warp of ONE chosen CTA appends (clock, warp, event) records to a global buffer.  Compiled out by default (TL is empty).
#if defined(M324_TIMELINE) && M324_TIMELINE
__device__ unsigned long long* g_tl_buf = nullptr;   // [0] = record counter, [1..] = records
__device__ int g_tl_cta = -1, g_tl_cap = 0;
enum : int { TL_S_WAIT = 1, TL_S_READY, TL_S_LOADED, TL_MAX_DONE, TL_ODONE_OK, TL_TURN_OK, TL_EXP_DONE, TL_TURN_PASSED, TL_P_ARRIVED,
             TL_QK_WAIT = 16, TL_QK_ISSUED, TL_PV_WAIT, TL_PV_ISSUED, TL_KV_LOADED = 24 };
__device__ __forceinline__ void tl_record(int ev) {
  if (static_cast<int>(blockIdx.x) != g_tl_cta || (threadIdx.x & 31) != 0 || g_tl_buf == nullptr) return;
  const unsigned long long i = atomicAdd(g_tl_buf, 1ull);
  if (i + 1 < static_cast<unsigned long long>(g_tl_cap))
    g_tl_buf[i + 1] = (static_cast<unsigned long long>(clock64()) << 16) | (static_cast<unsigned long long>(threadIdx.x >> 5) << 8) | ev;
}
#define TL(ev) tl_record(ev)
#else
#define TL(ev) ((void)0)
#endif

#ifndef M324_POLY_MASK
#define M324_POLY_MASK 0x00
#endif
#ifndef M324_POLY_PAIRS
#define M324_POLY_PAIRS 0x00
#endif
// Which of the 8 element PAIRS of every 16-column chunk take the packed FMA-pipe exponential (exp2_poly2) instead of MUFU.EX2
// (full tiles, P in TMEM).  The scale-and-shift x = s * c - m * c and the row sums are packed (FFMA2 / FADD2) for every pair.
constexpr int kPolyPairs = M324_POLY_PAIRS;
#ifndef M324_PACKED_SOFTMAX
#define M324_PACKED_SOFTMAX 0
#endif
// 1: the scale-and-shift and the row sums of the MUFU path as packed fp32x2 ops (FFMA2 / FADD2: 2.5 instead of 3.5 issue slots per element)
constexpr bool kPackedSoftmax = M324_PACKED_SOFTMAX != 0 || kPolyPairs != 0;
constexpr int kPolyMask = M324_POLY_MASK;   // which of every 8 consecutive exponentials go to the FMA pipes (0 = none: measured fastest)
#ifndef M324_ROWSUM_MMA
#define M324_ROWSUM_MMA 0
#endif
// Row sums l = sum_k P: 0 = fp32 adds in the softmax threads; 1 = a third MMA (P . ones, N = 16).  The MMA version frees
// 128 FADDs per row and tile, but every tcgen05.mma re-reads its 4 KB A tile (P) from shared memory, so the eight N = 16
// row-sum MMAs cost ~32 clk each (A-read bound, not the 8 clk their math needs) on the path the softmax warps wait on.
constexpr bool kRowSumMMA = M324_ROWSUM_MMA != 0;
#ifndef M324_P_TMEM
#define M324_P_TMEM 1
#endif
// Where P (fp16, the A operand of O += P V) lives: 1 = tensor memory (tcgen05.st by the softmax threads, tcgen05.mma with
// A from TMEM), 0 = 128B-swizzled shared memory.  Per 128 x 128 tile the shared-memory port then carries 64 KB (Q/K operand
// reads of S = Q K^T, V operand reads, the K/V TMA fill) instead of 128 KB (+ 32 KB of P stores + 32 KB of P operand reads):
// at 128 B/clk that is 512 instead of 1024 clk -- with P in shared memory the port was as binding as the MUFU unit.
constexpr bool kPTmem = M324_P_TMEM != 0;
#ifndef M324_TURN_EARLY
#define M324_TURN_EARLY 2
#endif
constexpr int kTurnEarly = M324_TURN_EARLY;   // 0 = pass the MUFU turn after the last exponential of the tile; measured 0/1/2/3/4: 767/794/819/812/812 TFLOP/s (T=32)
constexpr uint32_t TM_P = 384;   // P^0 at cols [384,448), P^1 at [448,512): 128 keys x fp16 = 64 columns per Q tile
static_assert(!(kPTmem && kRowSumMMA), "the row-sum MMA reads P from shared memory and its accumulator overlaps the TMEM P tiles");

// ---------------------------------------------------------------------------------------------------------------
// Softmax of one K/V tile for one query row (thread) of one softmax group.  Shared by both kernels below.
struct SoftmaxCtx {
  uint32_t t_s, t_o, t_l;  // TMEM addresses of this thread's lane quarter: S (128 cols), O (64 cols), row sum (col 0 of 16)
  uint32_t t_p;            // ... and P (64 cols of packed fp16 pairs), kPTmem
  uint32_t p_row;      // shared-space address of this row in the group's P buffer (two 16 KB K-major sub-blocks) + ((r & 7) << 4)
  int r;               // row within the 128-row Q tile
  float c;             // scale * log2(e)
  float m_run;         // running max (raw score units)
  float l_run;         // running row sum (kRowSumMMA: lives in TMEM at t_l instead)
};

__device__ __forceinline__ void rescale_o(const SoftmaxCtx& cx, float alpha) {
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t o[32];
    tmem_ld_32x32b_x32(cx.t_o + ch * 32, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
    tmem_st_32x32b_x32(cx.t_o + ch * 32, o);
  }
  if constexpr (kRowSumMMA) {
    uint32_t l;
    tmem_ld_32x32b_x1(cx.t_l, l);
    tmem_ld_wait();
    l = __float_as_uint(__uint_as_float(l) * alpha);
    tmem_st_32x32b_x1(cx.t_l, l);
  }
  tmem_st_wait();
}

// exp2 on the FMA / ALU pipes (Cody-Waite split + degree-4 polynomial, rel. error 3.6e-6 << fp16 rounding of P): the MUFU
// unit retires only 16 ex2 / clk / SM and is the binding unit of head-dim-64 attention, so a fixed fraction of the
// exponentials of every row is computed here instead, in the issue slots the MUFU-bound loop leaves idle.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;            // 1.5 * 2^23: the low mantissa bits of t now hold round(x)
  const float f = x - (t - 12582912.0f);      // fractional part in [-0.5, 0.5]
  float p = fmaf(0.009666368515383477f, f, 0.05592197584225006f);
  p = fmaf(p, f, 0.24022349038020416f);
  p = fmaf(p, f, 0.6931210452034274f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // * 2^round(x)
}

// The same for TWO values with packed fp32x2 arithmetic (FFMA2 / FADD2): 3 FADD2 + 4 FFMA2 + 2 FMNMX + 2 LEA = 5.5 issue slots
// per element instead of 11, which is what lets a share of the exponentials leave the MUFU unit without making the softmax
// warps issue-bound (kPolyPairs below).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  const uint64_t x = f2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t magic = f2_pack(12582912.0f, 12582912.0f);
  const uint64_t t = f2_add(x, magic);
  const uint64_t f = f2_sub(x, f2_sub(t, magic));
  uint64_t p = f2_fma(f2_pack(0.009666368515383477f, 0.009666368515383477f), f, f2_pack(0.05592197584225006f, 0.05592197584225006f));
  p = f2_fma(p, f, f2_pack(0.24022349038020416f, 0.24022349038020416f));
  p = f2_fma(p, f, f2_pack(0.6931210452034274f, 0.6931210452034274f));
  p = f2_fma(p, f, f2_pack(1.0f, 1.0f));
  float q0, q1, t0, t1;
  f2_unpack(p, q0, q1);
  f2_unpack(t, t0, t1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// P row r -> 128B-swizzled K-major tile pair: 16-byte chunk c8 (8 halves) of sub-block sb lives at
// sb*16KB + r*128 + ((c8 ^ (r&7)) << 4) = sb*16KB + ((r*128 + ((r&7) << 4)) ^ (c8 << 4)): one XOR of a per-thread constant
// (cx.p_row, a 32-bit shared-space address) and a st.shared.v4 with an immediate offset per store.
__device__ __forceinline__ void store_p8(const SoftmaxCtx& cx, int col0, const float (&pv)[8]) {
  const int sb = col0 >> 6, c8 = (col0 & 63) >> 3;
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((cx.p_row ^ static_cast<uint32_t>(c8 << 4)) + sb * TILE_BYTES),
               "r"(pack_half2(pv[0], pv[1])), "r"(pack_half2(pv[2], pv[3])), "r"(pack_half2(pv[4], pv[5])),
               "r"(pack_half2(pv[6], pv[7]))
               : "memory");
}

// Online-softmax state update: returns alpha (rescale of the running sum / O), sets mc = m * c.
__device__ __forceinline__ float advance_max(SoftmaxCtx& cx, float mx, float& mc, bool& warp_need) {
  float m_new = fmaxf(cx.m_run, mx);
  const bool need = (m_new - cx.m_run) * cx.c > 8.0f;   // lazy: only move the max when it grows by more than 2^8
  warp_need = __any_sync(0xffffffffu, need);
  if (!warp_need) m_new = cx.m_run;
  const float alpha = ex2_approx((cx.m_run - m_new) * cx.c);
  mc = m_new * cx.c;
  cx.m_run = m_new;
  return alpha;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One K/V tile: wait S, read it, release it (s_free), exponentiate into P (smem), update l / O scale, signal p_full.
// turn_wait / turn_pass (named-barrier ids, 0 = none) order the exponential phases of the two softmax warps that share an
// SM sub-partition (one per Q tile): the MUFU unit is the binding pipe, and without an order the two warps drift into
// running their exponentials at the same time (each at half rate) and then sit in their tensor-pipe waits at the same
// time (ncu: 24 % of softmax-warp samples in the o_done wait, XU pipe 65 % while active).  With strict alternation one
// warp owns the MUFU while the other reads S, takes its row max and waits for its P V.
__device__ __forceinline__ void softmax_tile(SoftmaxCtx& cx, int nvalid, bool first, uint64_t* s_full, uint32_t s_par,
                                             uint64_t* s_free, uint64_t* o_done, uint32_t o_par, uint64_t* p_full,
                                             int turn_wait = 0, int turn_pass = 0) {
  TL(TL_S_WAIT);
  mbar_wait(s_full, s_par);
  tc_fence_after();
  TL(TL_S_READY);
  float alpha, mc;
  bool warp_need;
  float ls[4] = {0.f, 0.f, 0.f, 0.f};
  uint64_t ls2[2] = {0ull, 0ull};      // packed row-sum accumulators of the kPolyPairs path (two fp32 zeros each)
  if (nvalid == 128) {
    // ---- full tile: S read once into 128 registers; the next Q K^T may overwrite S as soon as it is loaded
    uint32_t s[128];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(cx.t_s + cc * 32, &s[cc * 32]);
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(s_free);
    TL(TL_S_LOADED);
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 128; i += 8) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1])));
    }
    alpha = advance_max(cx, fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), mc, warp_need);
    TL(TL_MAX_DONE);
    if (!first) {  // the previous P V of this group must be complete before P (smem) is overwritten / O rescaled
      mbar_wait(o_done, o_par);
      tc_fence_after();
    }
    TL(TL_ODONE_OK);
    if (turn_wait) named_bar_sync(turn_wait, 64);
    TL(TL_TURN_OK);
    if constexpr (kPTmem) {
#pragma unroll
      for (int i0 = 0; i0 < 128; i0 += 16) {
        // the turn is passed kTurnEarly 16-column chunks before the end: the other warp's wake-up and its first exponentials
        // then overlap the tail of this warp's MUFU stream instead of leaving the unit idle during the hand-over
        if (kTurnEarly > 0 && i0 == 128 - 16 * kTurnEarly && turn_pass) {
          named_bar_arrive(turn_pass, 64);
          TL(TL_TURN_PASSED);
        }
        uint32_t pk[8];
        if constexpr (kPackedSoftmax) {
          const uint64_t c2 = f2_pack(cx.c, cx.c), nmc2 = f2_pack(-mc, -mc);
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(s[i0 + e]), __uint_as_float(s[i0 + e + 1])), c2, nmc2);
            float p0, p1;
            if ((kPolyPairs >> (e >> 1)) & 1) {
              exp2_poly2(x2, p0, p1);
            } else {
              float x0, x1;
              f2_unpack(x2, x0, x1);
              p0 = ex2_approx(x0);
              p1 = ex2_approx(x1);
            }
            ls2[(e >> 1) & 1] = f2_add(ls2[(e >> 1) & 1], f2_pack(p0, p1));
            pk[e >> 1] = pack_half2(p0, p1);
          }
        } else {
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float x0 = fmaf(__uint_as_float(s[i0 + e]), cx.c, -mc), x1 = fmaf(__uint_as_float(s[i0 + e + 1]), cx.c, -mc);
          const float p0 = ((kPolyMask >> (e & 7)) & 1) ? exp2_poly(x0) : ex2_approx(x0);
          const float p1 = ((kPolyMask >> ((e + 1) & 7)) & 1) ? exp2_poly(x1) : ex2_approx(x1);
          ls[e & 3] += p0;
          ls[(e + 1) & 3] += p1;
          pk[e >> 1] = pack_half2(p0, p1);
        }
        }
        tmem_st_32x32b_x8(cx.t_p + (i0 >> 1), pk);
      }
    } else {
#pragma unroll
      for (int i0 = 0; i0 < 128; i0 += 8) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = fmaf(__uint_as_float(s[i0 + e]), cx.c, -mc);
          pv[e] = ((kPolyMask >> e) & 1) ? exp2_poly(x) : ex2_approx(x);
          if constexpr (!kRowSumMMA) ls[e & 3] += pv[e];
        }
        store_p8(cx, i0, pv);
      }
    }
    if (turn_pass && !(kPTmem && kTurnEarly > 0)) named_bar_arrive(turn_pass, 64);
    TL(TL_EXP_DONE);
  } else {
    // ---- last, partial tile (key padding): two passes over TMEM through a 32-column buffer (register-light)
    float mx = -INFINITY;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      if (cc * 32 < nvalid) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(cx.t_s + cc * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, cc * 32 + i < nvalid ? __uint_as_float(t[i]) : -INFINITY);
      }
    }
    alpha = advance_max(cx, mx, mc, warp_need);
    if (!first) {
      mbar_wait(o_done, o_par);
      tc_fence_after();
    }
    if (turn_wait) named_bar_sync(turn_wait, 64);
    const int ncols_w = (nvalid + 15) & ~15;   // the P.V MMA reads whole 16-column groups
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      if (cc * 32 < ncols_w) {
        uint32_t t[32];
        tmem_ld_32x32b_x32(cx.t_s + cc * 32, t);
        tmem_ld_wait();
        if constexpr (kPTmem) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = cc * 32 + e < nvalid ? ex2_approx(fmaf(__uint_as_float(t[e]), cx.c, -mc)) : 0.f;
            const float p1 = cc * 32 + e + 1 < nvalid ? ex2_approx(fmaf(__uint_as_float(t[e + 1]), cx.c, -mc)) : 0.f;
            ls[e & 3] += p0;
            ls[(e + 1) & 3] += p1;
            pk[e >> 1] = pack_half2(p0, p1);
          }
          tmem_st_32x32b_x16(cx.t_p + cc * 16, pk);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (cc * 32 + g * 8 < ncols_w) {
              float pv[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float pe = ex2_approx(fmaf(__uint_as_float(t[g * 8 + e]), cx.c, -mc));
                pv[e] = cc * 32 + g * 8 + e < nvalid ? pe : 0.f;
                if constexpr (!kRowSumMMA) ls[e & 3] += pv[e];
              }
              store_p8(cx, cc * 32 + g * 8, pv);
            }
          }
        }
      }
    }
    tc_fence_before();
    mbar_arrive(s_free);
    if (turn_pass) named_bar_arrive(turn_pass, 64);
  }
  if constexpr (kPackedSoftmax) {
    float a0, a1, b0, b1;
    f2_unpack(ls2[0], a0, a1);
    f2_unpack(ls2[1], b0, b1);
    ls[0] += a0; ls[1] += a1; ls[2] += b0; ls[3] += b1;
  }
  if constexpr (!kRowSumMMA) cx.l_run = fmaf(cx.l_run, alpha, (ls[0] + ls[1]) + (ls[2] + ls[3]));   // alpha == 1 unless the max moved
  if (!first && warp_need) rescale_o(cx, alpha);   // rare (lazy rescale)
  if constexpr (kPTmem) tmem_st_wait();
  else fence_proxy_async_smem();
  tc_fence_before();
  mbar_arrive(p_full);
  TL(TL_P_ARRIVED);
}

// O row (fp32, TMEM) * scale -> fp16 -> global (64 contiguous halves of one head)
__device__ __forceinline__ void store_o_row(uint32_t t_o, float scale, __half* dst, bool valid) {
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t o[32];
    tmem_ld_32x32b_x32(t_o + ch * 32, o);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 val;
        val.x = pack_half2(__uint_as_float(o[8 * i + 0]) * scale, __uint_as_float(o[8 * i + 1]) * scale);
        val.y = pack_half2(__uint_as_float(o[8 * i + 2]) * scale, __uint_as_float(o[8 * i + 3]) * scale);
        val.z = pack_half2(__uint_as_float(o[8 * i + 4]) * scale, __uint_as_float(o[8 * i + 5]) * scale);
        val.w = pack_half2(__uint_as_float(o[8 * i + 6]) * scale, __uint_as_float(o[8 * i + 7]) * scale);
        *reinterpret_cast<uint4*>(dst + ch * 32 + 8 * i) = val;
      }
    }
  }
}

// MMA issue helpers.  Descriptors are (constant high word | start address >> 4), so stepping an operand by `bytes` is a
// 64-bit add of bytes >> 4; everything here is warp-uniform and fully unrolled so that the issuing warp's instruction
// stream (which has to feed ~40 small MMAs per K/V tile) stays a handful of uniform-datapath adds per MMA.
__device__ __forceinline__ uint64_t desc_of(uint32_t smem_addr) { return umma_desc_sw128(smem_addr, 16, 1024); }

__device__ __forceinline__ void issue_qk(uint32_t tmem_s, uint64_t dQ, uint64_t dK, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_s, dQ + 2 * k, dK + 2 * k, idesc, k > 0 ? 1u : 0u);
}
__device__ __forceinline__ void issue_pv(uint32_t tmem_o, uint32_t tmem_l, uint32_t tmem_p, uint64_t dP, uint64_t dV, uint64_t dOnes,
                                         uint32_t idesc_pv, uint32_t idesc_l, int nk16, bool acc) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    if (kk < nk16) {
      const uint64_t da = dP + (kk >> 2) * (TILE_BYTES >> 4) + (kk & 3) * 2;
      if constexpr (kPTmem) umma_f16_ts(tmem_o, tmem_p + kk * 8, dV + kk * (2048 >> 4), idesc_pv, (acc || kk > 0) ? 1u : 0u);
      else umma_f16_ss(tmem_o, da, dV + kk * (2048 >> 4), idesc_pv, (acc || kk > 0) ? 1u : 0u);
      if constexpr (kRowSumMMA) umma_f16_ss(tmem_l, da, dOnes, idesc_l, (acc || kk > 0) ? 1u : 0u);   // l += P . 1
    }
  }
}

// All-ones fp16 tile [16 x 64] (2 KB; any swizzle of a constant tile is itself), written once per CTA.
__device__ __forceinline__ void init_ones_tile(uint8_t* ones) {
  if (threadIdx.x < ONES_BYTES / 16) reinterpret_cast<uint4*>(ones)[threadIdx.x] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  fence_proxy_async_smem();
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 1 ("pair"): CTA = 256 query rows (two Q tiles), both groups walk the SAME K/V tiles.  Used for short K/V.
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // KV_STAGES
  uint64_t* v_full = k_full + KV_STAGES;   // KV_STAGES
  uint64_t* kv_empty = v_full + KV_STAGES; // KV_STAGES
  uint64_t* s_full = kv_empty + KV_STAGES; // 2
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* o_done = p_full + 2;           // 2
  uint64_t* s_free = o_done + 2;           // 2: S^q has been read into registers -> the next Q K^T may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  // Work item = (Q tile pair qt, head h, batch b), linear in blockIdx.x (qt fastest).  The first p.items_whole CTAs walk all
  // K/V tiles of their item.  The items of the last, partly filled wave are cut into p.split_parts K/V ranges, one CTA each
  // (launched last, so that they fill the SMs the last whole wave frees): those CTAs leave (O unnormalised, m, l) in the
  // workspace and attn_merge_kernel combines them.  With 492 items on 148 SMs (global layer, 32 frames) the launch takes
  // 3 1/3 instead of 4 item times.
  int item = blockIdx.x, j0 = 0, j1 = (p.Lk + 127) / 128, slot = -1;
  if (item >= p.items_whole) {
    slot = item - p.items_whole;
    const int part = slot % p.split_parts;
    item = p.items_whole + slot / p.split_parts;
    const int n_all = j1;
    j0 = static_cast<int>(static_cast<long>(part) * n_all / p.split_parts);
    j1 = static_cast<int>(static_cast<long>(part + 1) * n_all / p.split_parts);
  } else if (p.partial_parts > 0) {
    slot = item * p.partial_parts + p.partial_index;     // this launch is one K/V range of a multi-launch attention
  }
  // Frame loop (p.frame_loop = F > 1; the decoder: every frame attends with the SAME queries to its own <= 128 keys): the CTA
  // keeps its Q tiles and walks F consecutive batches through the K/V ring, one complete attention (fresh max / sum, own
  // epilogue) per "tile".  6144 one-tile work items become 768 eight-tile ones: the per-CTA latency (TMEM allocation, barrier
  // set-up, first TMA round trip) is paid once per 8 frames.
  const bool fl = p.frame_loop > 1;
  const int qt = item % p.n_qt, h = (item / p.n_qt) % p.H, bg = item / (p.n_qt * p.H);
  const int b = fl ? bg * p.frame_loop : bg;      // (first) batch of this CTA
  const int n_kv = fl ? min(p.frame_loop, p.B - b) : j1 - j0;   // K/V tiles of this CTA: global tile index j0 + i, or frames
  const long q_row0 = static_cast<long>(b / p.q_batch_div) * p.q_batch_rows + static_cast<long>(qt) * 256;
  const long kv_row0 = static_cast<long>(b) * p.kv_batch_rows + static_cast<long>(j0) * 128;
  const long kv_step = fl ? p.kv_batch_rows : 128;           // row distance between consecutive tiles
  const int Lk_loc = min(p.Lk - j0 * 128, n_kv * 128);   // keys of this CTA's range
  auto tile_keys = [&](int j) { return fl ? p.Lk : min(128, Lk_loc - j * 128); };
  const int nq = qt * 256 + 128 < p.Lq ? 2 : 1;   // the second Q tile of a ragged last CTA may be entirely out of range

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], nq);   // one commit per active Q tile (each has its own MMA-issuing warp)
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&s_full[q], 1);
      mbar_init(&p_full[q], 128);
      mbar_init(&o_done[q], 1);
      mbar_init(&s_free[q], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  init_ones_tile(smem + OFF_ONES);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, nq * TILE_BYTES);
      tma_load_2d(smem + OFF_Q, &tmQ, q_full, h * 64, static_cast<int>(q_row0));
      if (nq == 2) tma_load_2d(smem + OFF_Q + TILE_BYTES, &tmQ, q_full, h * 64, static_cast<int>(q_row0 + 128));
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_2d(smem + OFF_K + st * TILE_BYTES, &tmK, &k_full[st], h * 64, static_cast<int>(kv_row0 + j * kv_step));
        mbar_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_2d(smem + OFF_V + st * TILE_BYTES, &tmV, &v_full[st], h * 64, static_cast<int>(kv_row0 + j * kv_step));
      }
    }
  } else if (warp == 1 || warp == 2) {
    // One MMA-issuing warp PER Q TILE (warps 1 and 2): each walks only its own group's sequence
    //   Q K_0^T | for j: [S read out -> Q K_{j+1}^T] , [P_j written -> P_j V_j and l += P_j 1]
    // with blocking waits, so a group that runs ahead is never held back by the other group's barriers (with a single
    // in-order issuer, 12 % of the softmax warps' time was spent waiting for a P V that had not even been issued).
    // The whole warp walks the loop (warp-uniform addresses); one elected lane issues the MMAs / commits.
    const int q = warp - 1;
    if (q < nq) {
      const uint32_t idesc_qk = umma_idesc_f16(128, 128, false, false);
      const uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);  // B = V, MN-major
      const uint32_t idesc_l = umma_idesc_f16(128, 16, false, false);  // B = ones, K-major
      constexpr uint64_t kTile = TILE_BYTES >> 4;
      const uint64_t dQ = desc_of(smem_u32(smem + OFF_Q)) + q * kTile, dK = desc_of(smem_u32(smem + OFF_K)),
                     dV = desc_of(smem_u32(smem + OFF_V)), dP = desc_of(smem_u32(smem + OFF_P)) + q * 2 * kTile,
                     dOnes = desc_of(smem_u32(smem + OFF_ONES));
      const uint32_t t_s = tmem_base + TM_S + q * 128, t_o = tmem_base + TM_O + q * 64, t_l = tmem_base + TM_L + q * 16;
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        issue_qk(t_s, dQ, dK, idesc_qk);
        umma_commit(&s_full[q]);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        const int nk16 = (tile_keys(j) + 15) >> 4;
        if (j + 1 < n_kv) {
          const int st1 = (j + 1) % KV_STAGES;
          TL(TL_QK_WAIT);
          mbar_wait(&k_full[st1], ((j + 1) / KV_STAGES) & 1);
          mbar_wait(&s_free[q], j & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_qk(t_s, dQ, dK + st1 * kTile, idesc_qk);
            umma_commit(&s_full[q]);
          }
          __syncwarp();
          TL(TL_QK_ISSUED);
        }
        TL(TL_PV_WAIT);
        mbar_wait(&v_full[st], ph);
        mbar_wait(&p_full[q], j & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(t_o, t_l, tmem_base + TM_P + q * 64, dP, dV + st * kTile, dOnes, idesc_pv, idesc_l, nk16, !fl && j > 0);
          umma_commit(&o_done[q]);
          umma_commit(&kv_empty[st]);
        }
        __syncwarp();
        TL(TL_PV_ISSUED);
      }
    }
  } else if (warp >= 4 && ((warp - 4) >> 2) < nq) {
    // ---------------- softmax / correction / epilogue: one thread per query row ----------------
    const int q = (warp - 4) >> 2;            // Q tile of this warpgroup
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    SoftmaxCtx cx;
    cx.r = quarter * 32 + lane;
    cx.t_s = tmem_base + t_lane + TM_S + q * 128;
    cx.t_o = tmem_base + t_lane + TM_O + q * 64;
    cx.t_l = tmem_base + t_lane + TM_L + q * 16;
    cx.t_p = tmem_base + t_lane + TM_P + q * 64;
    cx.p_row = smem_u32(smem + OFF_P + q * 2 * TILE_BYTES) + cx.r * 128 + ((cx.r & 7) << 4);
    cx.c = p.scale * LOG2E;
    cx.m_run = -INFINITY;
    cx.l_run = 0.f;
    // MUFU turn-taking between this warp and the other Q tile's warp on the same SM sub-partition (named barriers
    // 1 + 2*quarter + q, 64 threads: 32 wait + 32 arrive).  Group 0 goes first: group 1 pre-arrives on its barrier.
    const bool turns = nq == 2 && n_kv >= 3 && p.tune_skew != 1;   // short K/V (decoder, Lk = 64): nothing to order, the barrier only costs
    const int bar_mine = turns ? 1 + 2 * quarter + q : 0, bar_other = turns ? 1 + 2 * quarter + (q ^ 1) : 0;
    if (turns && q == 1) named_bar_arrive(bar_other, 64);
    const long lq = static_cast<long>(qt) * 256 + q * 128 + cx.r;
    // normalised O row + log-sum-exp of batch bb -> global
    auto write_out = [&](int bb) {
      float lsum = cx.l_run;
      if constexpr (kRowSumMMA) {
        uint32_t lt;
        tmem_ld_32x32b_x1(cx.t_l, lt);
        tmem_ld_wait();
        lsum = __uint_as_float(lt);
      }
      store_o_row(cx.t_o, 1.0f / lsum, p.out + (static_cast<long>(bb) * p.Lq + lq) * p.o_ld + h * 64, lq < p.Lq);
      if (p.lse != nullptr && lq < p.Lq) p.lse[(static_cast<long>(bb) * p.Lq + lq) * p.lse_ld + h] = fmaf(cx.m_run, cx.c, log2f(lsum));
    };
    for (int j = 0; j < n_kv; ++j) {
      softmax_tile(cx, tile_keys(j), fl || j == 0, &s_full[q], j & 1, &s_free[q], &o_done[q], (j - 1) & 1, &p_full[q],
                   bar_mine, (q == 1 && j == n_kv - 1) ? 0 : bar_other);
      if (fl) {   // frame loop: this tile was a whole attention of batch b + j
        mbar_wait(&o_done[q], j & 1);
        tc_fence_after();
        write_out(b + j);
        tc_fence_before();         // O has been read: the next frame's P V (issued after this thread's next p_full) may overwrite it
        cx.m_run = -INFINITY;
        cx.l_run = 0.f;
      }
    }
    if (!fl) {
      mbar_wait(&o_done[q], (n_kv - 1) & 1);
      tc_fence_after();
      if (slot < 0) {
        write_out(b);
      } else {
        // partial result of a K/V range: O unnormalised (fp32), running max (log2 units) and row sum -> workspace
        float lsum = cx.l_run;
        if constexpr (kRowSumMMA) {
          uint32_t lt;
          tmem_ld_32x32b_x1(cx.t_l, lt);
          tmem_ld_wait();
          lsum = __uint_as_float(lt);
        }
        const long wrow = static_cast<long>(slot) * 256 + q * 128 + cx.r;
        float* wo = p.ws + wrow * 64;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(cx.t_o + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(wo + ch * 32 + 4 * i) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
        float2* wml = reinterpret_cast<float2*>(p.ws + static_cast<long>(p.split_slots) * 256 * 64);
        wml[wrow] = make_float2(cx.m_run * cx.c, lsum);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 2 ("split"): CTA = ONE 128-row Q tile; the two softmax groups take the first and second half of the K/V tiles
// and are merged in the CTA at the end (log-sum-exp combine through shared memory).  Twice as many, half as long work
// items as kernel 1: fills the SMs better when B*H*Lq/256 is only a few waves (the global layers: 492 -> 972 items).
constexpr int SPLIT_STAGES = 4;                      // ring entries of one (K, V) tile pair; entry e belongs to group e & 1
constexpr int SOFF_Q = 0;
constexpr int SOFF_KV = SOFF_Q + TILE_BYTES;         // entry: K at +0, V at +TILE_BYTES
constexpr int SOFF_P = SOFF_KV + SPLIT_STAGES * 2 * TILE_BYTES;
constexpr int SOFF_ML = SOFF_P + 4 * TILE_BYTES;     // m, l of group 1 (128 floats each)
constexpr int SOFF_ONES = SOFF_ML + 1024;
constexpr int SOFF_BAR = SOFF_ONES + ONES_BYTES;
constexpr int SPLIT_SMEM = SOFF_BAR + 256 + 1024;

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SOFF_BAR);
  uint64_t* q_full = bars;                       // 1
  uint64_t* k_full = bars + 1;                   // SPLIT_STAGES
  uint64_t* v_full = k_full + SPLIT_STAGES;      // SPLIT_STAGES
  uint64_t* kv_empty = v_full + SPLIT_STAGES;    // SPLIT_STAGES
  uint64_t* s_full = kv_empty + SPLIT_STAGES;    // 2
  uint64_t* p_full = s_full + 2;                 // 2
  uint64_t* o_done = p_full + 2;                 // 2
  uint64_t* s_free = o_done + 2;                 // 2
  uint64_t* merge_bar = s_free + 2;              // 1: group 1 has published (O, m, l)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(merge_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;
  const int n0 = (n_kv + 1) >> 1, n1 = n_kv - n0;        // tiles of group 0 / group 1 (n1 >= 1: dispatched for n_kv >= 2)
  const long q_row0 = static_cast<long>(b / p.q_batch_div) * p.q_batch_rows + static_cast<long>(qt) * 128;
  const long kv_row0 = static_cast<long>(b) * p.kv_batch_rows;
  // ring entry of (group g, local tile i): interleaved g0,g1,g0,g1,...; the odd tile out (n0 > n1) comes last
  auto entry = [&](int g, int i) { return i < n1 ? 2 * i + g : 2 * n1; };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < SPLIT_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&s_full[q], 1);
      mbar_init(&p_full[q], 128);
      mbar_init(&o_done[q], 1);
      mbar_init(&s_free[q], 128);
    }
    mbar_init(merge_bar, 128);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  init_ones_tile(smem + SOFF_ONES);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(smem + SOFF_Q, &tmQ, q_full, h * 64, static_cast<int>(q_row0));
      for (int e = 0; e < n_kv; ++e) {
        const int g = e < 2 * n1 ? (e & 1) : 0, i = e < 2 * n1 ? (e >> 1) : n1;
        const int jt = g == 0 ? i : n0 + i;                     // global K/V tile index
        const int st = e % SPLIT_STAGES;
        const uint32_t ph = (e / SPLIT_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        uint8_t* dst = smem + SOFF_KV + st * 2 * TILE_BYTES;
        mbar_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_2d(dst, &tmK, &k_full[st], h * 64, static_cast<int>(kv_row0 + jt * 128));
        mbar_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_2d(dst + TILE_BYTES, &tmV, &v_full[st], h * 64, static_cast<int>(kv_row0 + jt * 128));
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_qk = umma_idesc_f16(128, 128, false, false);
    const uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);
    const uint32_t idesc_l = umma_idesc_f16(128, 16, false, false);
    const uint64_t dQ = desc_of(smem_u32(smem + SOFF_Q)), dKV = desc_of(smem_u32(smem + SOFF_KV)),
                   dP = desc_of(smem_u32(smem + SOFF_P)), dOnes = desc_of(smem_u32(smem + SOFF_ONES));
    constexpr uint64_t kTile = TILE_BYTES >> 4;
    mbar_wait(q_full, 0);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int e = entry(g, 0);
      mbar_wait(&k_full[e % SPLIT_STAGES], (e / SPLIT_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_qk(tmem_base + TM_S + g * 128, dQ, dKV + (e % SPLIT_STAGES) * 2 * kTile, idesc_qk);
        umma_commit(&s_full[g]);
      }
      __syncwarp();
    }
    for (int i = 0; i < n0; ++i) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int ng = g == 0 ? n0 : n1;
        if (i + 1 < ng) {
          const int e = entry(g, i + 1);
          mbar_wait(&k_full[e % SPLIT_STAGES], (e / SPLIT_STAGES) & 1);
          mbar_wait(&s_free[g], i & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_qk(tmem_base + TM_S + g * 128, dQ, dKV + (e % SPLIT_STAGES) * 2 * kTile, idesc_qk);
            umma_commit(&s_full[g]);
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int ng = g == 0 ? n0 : n1;
        if (i < ng) {
          const int e = entry(g, i), st = e % SPLIT_STAGES;
          const int jt = g == 0 ? i : n0 + i;
          mbar_wait(&v_full[st], (e / SPLIT_STAGES) & 1);
          mbar_wait(&p_full[g], i & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(tmem_base + TM_O + g * 64, tmem_base + TM_L + g * 16, tmem_base + TM_P + g * 64, dP + g * 2 * kTile, dKV + (st * 2 + 1) * kTile, dOnes,
                     idesc_pv, idesc_l, (min(128, p.Lk - jt * 128) + 15) >> 4, i > 0);
            umma_commit(&o_done[g]);
            umma_commit(&kv_empty[st]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    SoftmaxCtx cx;
    cx.r = quarter * 32 + lane;
    cx.t_s = tmem_base + t_lane + TM_S + g * 128;
    cx.t_o = tmem_base + t_lane + TM_O + g * 64;
    cx.t_l = tmem_base + t_lane + TM_L + g * 16;
    cx.t_p = tmem_base + t_lane + TM_P + g * 64;
    cx.p_row = smem_u32(smem + SOFF_P + g * 2 * TILE_BYTES) + cx.r * 128 + ((cx.r & 7) << 4);
    cx.c = p.scale * LOG2E;
    cx.m_run = -INFINITY;
    cx.l_run = 0.f;
    const int ng = g == 0 ? n0 : n1;
    for (int i = 0; i < ng; ++i) {
      const int jt = g == 0 ? i : n0 + i;
      softmax_tile(cx, min(128, p.Lk - jt * 128), i == 0, &s_full[g], i & 1, &s_free[g], &o_done[g], (i - 1) & 1, &p_full[g]);
    }
    mbar_wait(&o_done[g], (ng - 1) & 1);
    tc_fence_after();
    float l_own = cx.l_run;
    if constexpr (kRowSumMMA) {
      uint32_t lsum;
      tmem_ld_32x32b_x1(cx.t_l, lsum);
      tmem_ld_wait();
      l_own = __uint_as_float(lsum);
    }
    float* ml = reinterpret_cast<float*>(smem + SOFF_ML);
    float* o1 = reinterpret_cast<float*>(smem + SOFF_P + 2 * TILE_BYTES);   // group 1's P buffer, free now: [128][64] fp32
    if (g == 1) {
      // publish (O1 unnormalised, m1, l1); row r, 16-byte chunk c at r*256 + ((c ^ (r & 15)) << 4)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(cx.t_o + ch * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const int c = ch * 8 + c4;
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(o1) + cx.r * 256 + ((c ^ (cx.r & 15)) << 4)) =
              make_uint4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
        }
      }
      ml[cx.r] = cx.m_run;
      ml[128 + cx.r] = l_own;
      mbar_arrive(merge_bar);     // release: the smem writes above are visible to the waiters
    } else {
      mbar_wait(merge_bar, 0);
      const float m1 = ml[cx.r], l1 = ml[128 + cx.r];
      const float m = fmaxf(cx.m_run, m1);
      const float a0 = ex2_approx((cx.m_run - m) * cx.c), a1 = ex2_approx((m1 - m) * cx.c);
      const float inv = 1.0f / (l_own * a0 + l1 * a1);
      const float w0 = a0 * inv, w1 = a1 * inv;
      const long lq = static_cast<long>(qt) * 128 + cx.r;
      __half* dst = p.out + (static_cast<long>(b) * p.Lq + lq) * p.o_ld + h * 64;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(cx.t_o + ch * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = ch * 8 + 2 * i;
          const float4 x0 = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(o1) + cx.r * 256 + ((c ^ (cx.r & 15)) << 4));
          const float4 x1 = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(o1) + cx.r * 256 + (((c + 1) ^ (cx.r & 15)) << 4));
          uint4 val;
          val.x = pack_half2(fmaf(__uint_as_float(o[8 * i + 0]), w0, x0.x * w1), fmaf(__uint_as_float(o[8 * i + 1]), w0, x0.y * w1));
          val.y = pack_half2(fmaf(__uint_as_float(o[8 * i + 2]), w0, x0.z * w1), fmaf(__uint_as_float(o[8 * i + 3]), w0, x0.w * w1));
          val.z = pack_half2(fmaf(__uint_as_float(o[8 * i + 4]), w0, x1.x * w1), fmaf(__uint_as_float(o[8 * i + 5]), w0, x1.y * w1));
          val.w = pack_half2(fmaf(__uint_as_float(o[8 * i + 6]), w0, x1.z * w1), fmaf(__uint_as_float(o[8 * i + 7]), w0, x1.w * w1));
          if (lq < p.Lq) *reinterpret_cast<uint4*>(dst + ch * 32 + 8 * i) = val;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 3 ("items"): PERSISTENT CTAs for launches made of many SHORT work items -- the local blocks (324 tokens per frame), the
// DINOv2 blocks (257) and the latent-token blocks: <= 4 K/V tiles per item, hundreds to thousands of items.  One CTA per item
// paid the CTA set-up (TMEM allocation, barrier init, descriptor fetch, first TMA round trip) and a drained pipeline per ~3
// tile steps: 768 items ran as 5.2 waves of ~9 us (ncu: 47-52 us for 10 GFLOP).  Here every CTA owns a CONTIGUOUS chunk of items
// (the ragged second Q-tile pair of a (frame, head) follows its full first pair, so chunks mix both and share K/V in L2) and
// streams them through the same pipeline: Q is double-buffered (the second slot lives in the P region, unused with P in TMEM),
// the K/V ring runs across item boundaries, so the loads of item n+1 and its first Q K^T are in flight while item n finishes; each
// item ends with its own epilogue (fresh max / sum), like the decoder's frame loop.  Per-group tile counters carry the barrier
// parities across items; in an item whose second Q tile is out of range group 1 issues nothing but still walks the item's loads and
// arrives on the barriers that count both groups (so no group can run more than one barrier phase ahead of the ring).  No MUFU
// turn-taking (the groups are not in lock step across items).
static_assert(kPTmem, "attn_items_kernel keeps its second Q slot in the shared-memory P region");
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_items_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                 // 2
  uint64_t* q_empty = bars + 2;            // 2
  uint64_t* k_full = bars + 4;             // KV_STAGES
  uint64_t* v_full = k_full + KV_STAGES;   // KV_STAGES
  uint64_t* kv_empty = v_full + KV_STAGES; // KV_STAGES
  uint64_t* s_full = kv_empty + KV_STAGES; // 2
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* o_done = p_full + 2;           // 2
  uint64_t* s_free = o_done + 2;           // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_kv = (p.Lk + 127) / 128;
  const int total = p.items_whole;
  const int it0 = static_cast<int>(static_cast<long>(blockIdx.x) * total / gridDim.x);
  const int it1 = static_cast<int>(static_cast<long>(blockIdx.x + 1) * total / gridDim.x);
  auto nq_of = [&](int qt) { return qt * 256 + 128 < p.Lq ? 2 : 1; };
  auto keys_of = [&](int j) { return min(128, p.Lk - j * 128); };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 2);
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&s_full[q], 1);
      mbar_init(&p_full[q], 128);
      mbar_init(&o_done[q], 1);
      mbar_init(&s_free[q], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      int jj = 0;
      for (int item = it0, n = 0; item < it1; ++item, ++n) {
        const int qt = item % p.n_qt, h = (item / p.n_qt) % p.H, b = item / (p.n_qt * p.H);
        const int nq = nq_of(qt), slot = n & 1;
        const long q_row0 = static_cast<long>(b / p.q_batch_div) * p.q_batch_rows + static_cast<long>(qt) * 256;
        const long kv_row0 = static_cast<long>(b) * p.kv_batch_rows;
        uint8_t* qdst = smem + (slot ? OFF_P : OFF_Q);
        mbar_wait(&q_empty[slot], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[slot], nq * TILE_BYTES);
        tma_load_2d(qdst, &tmQ, &q_full[slot], h * 64, static_cast<int>(q_row0));
        if (nq == 2) tma_load_2d(qdst + TILE_BYTES, &tmQ, &q_full[slot], h * 64, static_cast<int>(q_row0 + 128));
        for (int j = 0; j < n_kv; ++j, ++jj) {
          const int st = jj % KV_STAGES;
          mbar_wait(&kv_empty[st], ((jj / KV_STAGES) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], TILE_BYTES);
          tma_load_2d(smem + OFF_K + st * TILE_BYTES, &tmK, &k_full[st], h * 64, static_cast<int>(kv_row0 + j * 128));
          mbar_expect_tx(&v_full[st], TILE_BYTES);
          tma_load_2d(smem + OFF_V + st * TILE_BYTES, &tmV, &v_full[st], h * 64, static_cast<int>(kv_row0 + j * 128));
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    const int q = warp - 1;
    const uint32_t idesc_qk = umma_idesc_f16(128, 128, false, false);
    const uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);
    constexpr uint64_t kTile = TILE_BYTES >> 4;
    const uint64_t dQ0 = desc_of(smem_u32(smem + OFF_Q)) + q * kTile, dQ1 = desc_of(smem_u32(smem + OFF_P)) + q * kTile,
                   dK = desc_of(smem_u32(smem + OFF_K)), dV = desc_of(smem_u32(smem + OFF_V));
    const uint32_t t_s = tmem_base + TM_S + q * 128, t_o = tmem_base + TM_O + q * 64, t_p = tmem_base + TM_P + q * 64;
    int jj = 0, cnt = 0;      // jj: tiles the CTA has walked (K/V ring position); cnt: tiles THIS group has processed (barrier parities)
    for (int item = it0, n = 0; item < it1; ++item, ++n, jj += n_kv) {
      const int qt = item % p.n_qt;
      const int nq = nq_of(qt), slot = n & 1;
      if (q >= nq) {
        // This group has no Q tile in the item, but it still WALKS it: it waits for every load and arrives on the barriers that
        // count both groups.  A group that skipped the item could run two phases ahead of the ring, where an mbarrier parity
        // wait aliases (parity only tells adjacent phases apart) and would let it issue Q K^T on a stage that is still being
        // filled -- seen as run-to-run differences of back-to-back forwards before this walk existed (scripts/determinism_check.py).
        mbar_wait(&q_full[slot], (n >> 1) & 1);
        if (elect_one()) mbar_arrive(&q_empty[slot]);
        __syncwarp();
        for (int j = 0; j < n_kv; ++j) {
          const int t = jj + j, st = t % KV_STAGES;
          mbar_wait(&k_full[st], (t / KV_STAGES) & 1);
          mbar_wait(&v_full[st], (t / KV_STAGES) & 1);
          if (elect_one()) mbar_arrive(&kv_empty[st]);
          __syncwarp();
        }
        continue;
      }
      const uint64_t dQ = slot ? dQ1 : dQ0;
      mbar_wait(&q_full[slot], (n >> 1) & 1);
      mbar_wait(&k_full[jj % KV_STAGES], (jj / KV_STAGES) & 1);
      if (cnt > 0) mbar_wait(&s_free[q], (cnt - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_qk(t_s, dQ, dK + (jj % KV_STAGES) * kTile, idesc_qk);
        umma_commit(&s_full[q]);
        if (n_kv == 1) umma_commit(&q_empty[slot]);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j, ++cnt) {
        const int t = jj + j, st = t % KV_STAGES;
        if (j + 1 < n_kv) {
          const int st1 = (t + 1) % KV_STAGES;
          mbar_wait(&k_full[st1], ((t + 1) / KV_STAGES) & 1);
          mbar_wait(&s_free[q], cnt & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_qk(t_s, dQ, dK + st1 * kTile, idesc_qk);
            umma_commit(&s_full[q]);
            if (j + 2 == n_kv) umma_commit(&q_empty[slot]);   // the item's last Q K^T: its Q slot may be refilled once these MMAs have completed
          }
          __syncwarp();
        }
        mbar_wait(&v_full[st], (t / KV_STAGES) & 1);
        mbar_wait(&p_full[q], cnt & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(t_o, 0, t_p, 0, dV + st * kTile, 0, idesc_pv, 0, (keys_of(j) + 15) >> 4, j > 0);
          umma_commit(&o_done[q]);
          umma_commit(&kv_empty[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int q = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    SoftmaxCtx cx;
    cx.r = quarter * 32 + lane;
    cx.t_s = tmem_base + t_lane + TM_S + q * 128;
    cx.t_o = tmem_base + t_lane + TM_O + q * 64;
    cx.t_l = tmem_base + t_lane + TM_L + q * 16;
    cx.t_p = tmem_base + t_lane + TM_P + q * 64;
    cx.p_row = 0;
    cx.c = p.scale * LOG2E;
    int cnt = 0;
    for (int item = it0; item < it1; ++item) {
      const int qt = item % p.n_qt, h = (item / p.n_qt) % p.H, b = item / (p.n_qt * p.H);
      if (q >= nq_of(qt)) continue;
      cx.m_run = -INFINITY;
      cx.l_run = 0.f;
      for (int j = 0; j < n_kv; ++j, ++cnt)
        softmax_tile(cx, keys_of(j), j == 0, &s_full[q], cnt & 1, &s_free[q], &o_done[q], (cnt - 1) & 1, &p_full[q]);
      mbar_wait(&o_done[q], (cnt - 1) & 1);
      tc_fence_after();
      const long lq = static_cast<long>(qt) * 256 + q * 128 + cx.r;
      const long orow = static_cast<long>(b) * p.Lq + lq;
      store_o_row(cx.t_o, 1.0f / cx.l_run, p.out + orow * p.o_ld + h * 64, lq < p.Lq);
      if (p.lse != nullptr && lq < p.Lq) p.lse[orow * p.lse_ld + h] = fmaf(cx.m_run, cx.c, log2f(cx.l_run));
      tc_fence_before();       // O has been read: the next item's first P V (issued after this thread's next p_full) may overwrite it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Combine the split_parts partial results of every split work item (log-sum-exp merge), one warp per query row.
__global__ void __launch_bounds__(256) attn_merge_kernel(const AttnArgs p) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);      // (split item, row of its 256 queries)
  const int n_split_items = p.split_slots / p.split_parts;
  if (row >= static_cast<long>(n_split_items) * 256) return;
  const int it = static_cast<int>(row >> 8), r = static_cast<int>(row & 255);
  const int item = p.items_whole + it;
  const int qt = item % p.n_qt, h = (item / p.n_qt) % p.H, b = item / (p.n_qt * p.H);
  const long lq = static_cast<long>(qt) * 256 + r;
  if (lq >= p.Lq) return;
  const float2* wml = reinterpret_cast<const float2*>(p.ws + static_cast<long>(p.split_slots) * 256 * 64);
  float m = -INFINITY;
  for (int s = 0; s < p.split_parts; ++s) m = fmaxf(m, wml[(static_cast<long>(it) * p.split_parts + s) * 256 + r].x);
  float l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int s = 0; s < p.split_parts; ++s) {
    const long wrow = (static_cast<long>(it) * p.split_parts + s) * 256 + r;
    const float2 ml = wml[wrow];
    const float w = ex2_approx(ml.x - m);
    const float2 o = *reinterpret_cast<const float2*>(p.ws + wrow * 64 + 2 * lane);
    l = fmaf(ml.y, w, l);
    o0 = fmaf(o.x, w, o0);
    o1 = fmaf(o.y, w, o1);
  }
  const float inv = 1.0f / l;
  const long orow = static_cast<long>(b) * p.Lq + lq;
  *reinterpret_cast<uint32_t*>(p.out + orow * p.o_ld + h * 64 + 2 * lane) = pack_half2(o0 * inv, o1 * inv);
  if (p.lse != nullptr && lane == 0) p.lse[orow * p.lse_ld + h] = m + log2f(l);
}

}  // namespace

long attention_partial_bytes(int B, int H, int Lq, int parts) {
  return static_cast<long>((Lq + 255) / 256) * H * B * parts * 256 * (64 + 2) * 4;
}

int attention_merge(const AttnArgs& a_in, cudaStream_t stream) {
  AttnArgs a = a_in;
  M324_REQUIRE(a.out && a.ws && a.partial_parts >= 1 && a.partial_parts <= 16, "attention_merge: bad arguments");
  M324_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0, "attention_merge: empty problem");
  M324_REQUIRE(a.ws_bytes >= attention_partial_bytes(a.B, a.H, a.Lq, a.partial_parts), "attention_merge: workspace too small");
  a.n_qt = (a.Lq + 255) / 256;
  const long items = static_cast<long>(a.n_qt) * a.H * a.B;
  a.items_whole = 0;
  a.split_parts = a.partial_parts;
  a.split_slots = static_cast<int>(items * a.partial_parts);
  const long rows = items * 256;
  M324_CUDA(launch_pdl(attn_merge_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, stream, a));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

long attention_workspace_bytes() {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  return static_cast<long>(sms) * 256 * (64 + 2) * 4;
}

// Work decomposition of one attn_kernel launch (host side, no device needed: m324_attention_plan exposes it to the CPU tests).
// Fills a.n_qt / frame_loop / items_whole / split_parts / split_slots (see attn_kernel for their meaning), the grid size and
// the number of query rows attn_merge_kernel has to combine (0 = no merge launch).
int attention_plan(AttnArgs& a, int sms, int* grid_x, long* merge_rows) {
  M324_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0 && sms > 0, "attention: empty problem B=%d H=%d Lq=%d Lk=%d", a.B, a.H, a.Lq, a.Lk);
  const int n_kv = (a.Lk + 127) / 128;
  a.n_qt = (a.Lq + 255) / 256;
  // frame loop: shared queries, one K/V tile per batch, enough batches to amortise the CTA set-up over
  a.frame_loop = (n_kv == 1 && a.q_batch_rows == 0 && a.B >= 8 && a.partial_parts == 0 && a.tune_event != 1) ? 8 : 1;
  const long items = static_cast<long>(a.n_qt) * a.H * ((a.B + a.frame_loop - 1) / a.frame_loop);
  M324_REQUIRE(items < (1l << 31), "attention: too many work items");
  a.items_whole = static_cast<int>(items);
  a.split_parts = 1;
  a.split_slots = 0;
  // item loop (attn_items_kernel): many short items -> one persistent CTA per SM walking a contiguous chunk of them
  a.item_loop = (kPTmem && !kRowSumMMA && a.frame_loop == 1 && a.partial_parts == 0 && n_kv <= 4 && items >= 2l * sms && a.tune_event != 1 &&
                 a.tune_skew != 2) ? 1 : 0;
  if (a.item_loop) {
    *grid_x = sms;
    *merge_rows = 0;
    return M324_OK;
  }
  const int rem = static_cast<int>(items % sms);
  const long waves = items / sms;
  if (a.partial_parts > 0) {
    M324_REQUIRE(a.partial_index >= 0 && a.partial_index < a.partial_parts && a.partial_parts <= 16, "attention: bad partial index %d of %d",
                 a.partial_index, a.partial_parts);
    M324_REQUIRE(a.ws != nullptr && a.ws_bytes >= attention_partial_bytes(a.B, a.H, a.Lq, a.partial_parts),
                 "attention: partial launch needs a workspace of attention_partial_bytes()");
    M324_REQUIRE(items * a.partial_parts < (1l << 31), "attention: too many partial slots");
    a.split_parts = a.partial_parts;
    a.split_slots = static_cast<int>(items * a.partial_parts);
  } else if (a.frame_loop == 1 && a.ws != nullptr && a.tune_event != 1 && waves <= 24 && rem > 0 && 2 * rem <= sms) {
    // Tail split: only when the launch is a few waves long, the last wave is at most half full and the caller lent a workspace
    int parts = sms / rem;
    const int cap = waves >= 1 ? 4 : 8;           // a launch smaller than one wave (encoder cross-attention: 12 items) splits further
    if (parts > cap) parts = cap;
    if (parts > n_kv / 4) parts = n_kv / 4;       // at least 4 K/V tiles per part
    if (parts >= 2 && static_cast<long>(rem) * parts * 256 * (64 + 2) * 4 <= a.ws_bytes) {
      a.items_whole = static_cast<int>(items - rem);
      a.split_parts = parts;
      a.split_slots = rem * parts;
    }
  }
  // partial launches: one CTA per work item (split_slots only sizes the workspace layout there)
  *grid_x = a.items_whole + (a.partial_parts > 0 ? 0 : a.split_slots);
  *merge_rows = (a.split_slots > 0 && a.partial_parts == 0) ? static_cast<long>(a.split_slots / a.split_parts) * 256 : 0;
  return M324_OK;
}

int attention(const AttnArgs& a_in, cudaStream_t stream) {
  AttnArgs a = a_in;
  M324_REQUIRE(a.q && a.k && a.v && a.out, "attention: null pointer");
  M324_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "attention: empty problem B=%d H=%d Lq=%d Lk=%d", a.B, a.H, a.Lq, a.Lk);
  M324_REQUIRE(a.q_ld % 8 == 0 && a.k_ld % 8 == 0 && a.v_ld % 8 == 0 && a.o_ld % 8 == 0, "attention: row strides must be multiples of 8");
  M324_REQUIRE(a.q_ld >= a.H * 64 && a.k_ld >= a.H * 64 && a.v_ld >= a.H * 64 && a.o_ld >= a.H * 64, "attention: row stride < H*64");
  M324_REQUIRE(a.q_batch_div >= 1, "attention: q_batch_div must be >= 1");
  M324_REQUIRE(a.q_rows >= (long)((a.B - 1) / a.q_batch_div) * a.q_batch_rows + a.Lq && a.kv_rows >= (long)(a.B - 1) * a.kv_batch_rows + a.Lk,
               "attention: q_rows / kv_rows smaller than the addressed range");
  static PerDeviceOnce configured;
  if (configured.need()) {
    M324_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    configured.mark();
  }
  CUtensorMap tq, tk, tv;
  uint32_t box[2] = {64, 128};
  {
    uint64_t dims[2] = {static_cast<uint64_t>(a.H) * 64, static_cast<uint64_t>(a.q_rows)};
    uint64_t str[1] = {static_cast<uint64_t>(a.q_ld) * 2};
    int e = make_tmap_16b(&tq, a.q, 2, dims, str, box);
    if (e) return e;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(a.H) * 64, static_cast<uint64_t>(a.kv_rows)};
    uint64_t str[1] = {static_cast<uint64_t>(a.k_ld) * 2};
    int e = make_tmap_16b(&tk, a.k, 2, dims, str, box);
    if (e) return e;
    str[0] = static_cast<uint64_t>(a.v_ld) * 2;
    e = make_tmap_16b(&tv, a.v, 2, dims, str, box);
    if (e) return e;
  }
  // Work-item shape: "pair" by default; "split" (128-row items, K/V halves merged in the CTA) on request (knob 0 = 2).
  const int n_kv = (a.Lk + 127) / 128;
  const bool split = a.tune_event == 2 && n_kv >= 2 && a.lse == nullptr && a.partial_parts == 0;   // measured on B200: the pair kernel is faster at every model shape
  if (split) {
    static PerDeviceOnce configured2;
    if (configured2.need()) {
      M324_CUDA(cudaFuncSetAttribute(attn_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SPLIT_SMEM));
      configured2.mark();
    }
    dim3 grid((a.Lq + 127) / 128, a.H, a.B);
    M324_CUDA(launch_pdl(attn_split_kernel, grid, dim3(ATT_THREADS), SPLIT_SMEM, stream, tq, tk, tv, a));
  } else {
    int grid_x = 0;
    long merge_rows = 0;
    const int e = attention_plan(a, sm_count() > 0 ? sm_count() : 148, &grid_x, &merge_rows);
    if (e) return e;
    dim3 grid(static_cast<unsigned>(grid_x), 1, 1);
    if (a.item_loop) {
      static PerDeviceOnce configured3;
      if (configured3.need()) {
        M324_CUDA(cudaFuncSetAttribute(attn_items_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured3.mark();
      }
      M324_CUDA(launch_pdl(attn_items_kernel, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tq, tk, tv, a));
    } else
    M324_CUDA(launch_pdl(attn_kernel, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tq, tk, tv, a));
    if (merge_rows > 0)
      M324_CUDA(launch_pdl(attn_merge_kernel, dim3(static_cast<unsigned>((merge_rows + 7) / 8)), dim3(256), 0, stream, a));
  }
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

#if defined(M324_TIMELINE) && M324_TIMELINE
// Profiling builds only (not part of include/m324.h): point the timeline at a device buffer of `cap` u64 words (word 0 = the
// record counter, zero it first) and choose the CTA to trace.
extern "C" int m324_timeline_set(void* buf, int cap, int cta) {
  unsigned long long* b = static_cast<unsigned long long*>(buf);
  if (cudaMemcpyToSymbol(g_tl_buf, &b, sizeof(b)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_tl_cap, &cap, sizeof(cap)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_tl_cta, &cta, sizeof(cta)) != cudaSuccess) return -1;
  return 0;
}
#endif

}  // namespace m324
