// tcgen05 flash-attention forward, head dim 64, non-causal, no bias:  O = softmax(Q K^T * scale) V
//
// Replaces xformers.ops.memory_efficient_attention as called by the reference (model/transformer.py:134-139,
// 209-214; layout [B, L, H, Dh], attn_bias=None, p=0) and the attention inside the DINOv2 ViT blocks.
//
// One CTA = one (batch, head) x 256 query rows (two 128-row Q tiles), looping over 128-row K/V tiles:
//   warp 0      : TMA producer  (Q once; K_j / V_j through a 3-stage ring)
//   warp 1      : tcgen05.mma issuer + TMEM owner.   S^q = Q^q K_j^T  (128x128x64, fp32 in TMEM)
//                                                    O^q += P^q V_j   (128x64x128, fp32 in TMEM)
//   warps 2-3   : idle
//   warps 4-7   : softmax for Q tile 0 (one thread per query row)      warps 8-11 : softmax for Q tile 1  
// The two Q tiles ping-pong on the tensor pipe (while one tile is in softmax the other's MMAs run).  Softmax is online
// with a LAZY rescale: the running max only moves (and O in TMEM is only rescaled) when it grows by more than 2^8,
// so P <= 256 fits fp16 and the O read-modify-write is rare.  P is written to 128B-swizzled shared memory as the
// K-major A operand of the second MMA; V is consumed MN-major straight from its TMA tile (no transpose).
// Normalisation by the fp32 row sum happens once, in the epilogue.  Q/K/V/P are fp16, all statistics fp32.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace m324 {

namespace {

constexpr int ATT_THREADS = 384;  // warpgroup 0: TMA / MMA / 2 idle warps; warpgroups 1,2: softmax
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB: 128 rows x 64 fp16
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + 2 * TILE_BYTES;
constexpr int OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
constexpr int OFF_P = OFF_V + KV_STAGES * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + 4 * TILE_BYTES;
constexpr int ATT_SMEM = OFF_BAR + 256 + 1024;
constexpr uint32_t TM_S = 0;     // S^0 at cols [0,128), S^1 at [128,256)
constexpr uint32_t TM_O = 256;   // O^0 at cols [256,320), O^1 at [320,384)
constexpr float LOG2E = 1.4426950408889634f;

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
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Lk + 127) / 128;
  const long q_row0 = static_cast<long>(b / p.q_batch_div) * p.q_batch_rows + static_cast<long>(qt) * 256;
  const long kv_row0 = static_cast<long>(b) * p.kv_batch_rows;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
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
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      tma_load_2d(smem + OFF_Q, &tmQ, q_full, h * 64, static_cast<int>(q_row0));
      tma_load_2d(smem + OFF_Q + TILE_BYTES, &tmQ, q_full, h * 64, static_cast<int>(q_row0 + 128));
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_2d(smem + OFF_K + st * TILE_BYTES, &tmK, &k_full[st], h * 64, static_cast<int>(kv_row0 + j * 128));
        mbar_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_2d(smem + OFF_V + st * TILE_BYTES, &tmV, &v_full[st], h * 64, static_cast<int>(kv_row0 + j * 128));
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_qk = umma_idesc_f16(128, 128, false, false);
      const uint32_t idesc_pv = umma_idesc_f16(128, 64, false, true);  // B = V, MN-major
      const uint32_t sQ = smem_u32(smem + OFF_Q), sK = smem_u32(smem + OFF_K), sV = smem_u32(smem + OFF_V),
                     sP = smem_u32(smem + OFF_P);
      auto issue_qk = [&](int q, int st) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = umma_desc_sw128(sQ + q * TILE_BYTES + k * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(sK + st * TILE_BYTES + k * 32, 16, 1024);
          umma_f16_ss(tmem_base + TM_S + q * 128, da, db, idesc_qk, k > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int q, int st, int nk16, bool acc) {
        for (int kk = 0; kk < nk16; ++kk) {
          const uint64_t da = umma_desc_sw128(sP + q * 2 * TILE_BYTES + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(sV + st * TILE_BYTES + kk * 2048, 16, 1024);
          umma_f16_ss(tmem_base + TM_O + q * 64, da, db, idesc_pv, (acc || kk > 0) ? 1u : 0u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int q = 0; q < 2; ++q) {
        issue_qk(q, 0);
        umma_commit(&s_full[q]);
      }
      if (!p.tune_event) {
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        const int nvalid = min(128, p.Lk - j * 128);
        const int nk16 = (nvalid + 15) >> 4;
        const int st1 = (j + 1) % KV_STAGES;
        const uint32_t ph1 = ((j + 1) / KV_STAGES) & 1;
        // S^q_{j+1} = Q^q K_{j+1}^T is issued as soon as the softmax warps have pulled S^q_j into registers, i.e. it runs
        // on the tensor pipe while they exponentiate; P^q_j V_j follows when P^q_j has been written.
        if (j + 1 < n_kv) {
          mbar_wait(&k_full[st1], ph1);
          for (int q = 0; q < 2; ++q) {
            mbar_wait(&s_free[q], j & 1);
            tc_fence_after();
            issue_qk(q, st1);
            umma_commit(&s_full[q]);
          }
        }
        mbar_wait(&v_full[st], ph);
        for (int q = 0; q < 2; ++q) {
          mbar_wait(&p_full[q], j & 1);
          tc_fence_after();
          issue_pv(q, st, nk16, j > 0);
          umma_commit(&o_done[q]);
        }
        umma_commit(&kv_empty[st]);
      }
      } else {
      // Event-driven issue: per Q tile the order is Q K_1^T, P_0 V_0, Q K_2^T, P_1 V_1, ...; each step is issued as soon
      // as ITS barrier completes (S^q read out -> next Q K^T; P^q written -> P V), whichever tile is ready first.  The
      // two softmax groups therefore run out of phase (group 1 starts half a period late) and share the MUFU unit
      // instead of idling and saturating it together.
      int nqk[2] = {1, 1};      // next K/V tile whose Q K^T is to be issued, per Q tile
      int npv[2] = {0, 0};      // next K/V tile whose P V is to be issued, per Q tile
      uint32_t polls = 0;
      uint64_t t_start = 0;
      while (npv[0] < n_kv || npv[1] < n_kv) {
        if ((++polls & 0x3FFF) == 0) {   // bounded: a protocol bug must trap, not hang the GPU
          const uint64_t t = global_ns();
          if (t_start == 0) t_start = t;
          else if (t - t_start > 20000000000ull) mbar_timeout(0xA77E, polls);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (nqk[q] < n_kv) {
            const int jn = nqk[q], st1 = jn % KV_STAGES;
            if (mbar_test_wait(&s_free[q], (jn - 1) & 1) && mbar_test_wait(&k_full[st1], (jn / KV_STAGES) & 1)) {
              tc_fence_after();
              issue_qk(q, st1);
              umma_commit(&s_full[q]);
              nqk[q] = jn + 1;
            }
          }
          if (npv[q] < n_kv) {
            const int jp = npv[q], st = jp % KV_STAGES;
            if (mbar_test_wait(&p_full[q], jp & 1) && mbar_test_wait(&v_full[st], (jp / KV_STAGES) & 1)) {
              tc_fence_after();
              const int nvalid = min(128, p.Lk - jp * 128);
              issue_pv(q, st, (nvalid + 15) >> 4, jp > 0);
              umma_commit(&o_done[q]);
              npv[q] = jp + 1;
              // the K/V stage is free once BOTH tiles' P V (and, earlier in program order, both Q K^T) were issued
              if (npv[q ^ 1] > jp) umma_commit(&kv_empty[st]);
            }
          }
        }
      }
      }
    }
  }
  } else {
    // ---------------- softmax / correction / epilogue: one thread per query row ----------------
    const int q = (warp - 4) >> 2;            // Q tile of this warpgroup
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;        // row within the Q tile
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + t_lane + TM_S + q * 128;
    const uint32_t t_o = tmem_base + t_lane + TM_O + q * 64;
    uint8_t* sPq = smem + OFF_P + q * 2 * TILE_BYTES;
    const float c = p.scale * LOG2E;
    float m_run = -INFINITY, l_run = 0.f;
    // Online-softmax state update shared by both tile paths: returns alpha (rescale of the running sum / O), sets mc.
    auto advance_max = [&](float mx, float& mc, bool& warp_need) -> float {
      float m_new = fmaxf(m_run, mx);
      const bool need = (m_new - m_run) * c > 8.0f;   // lazy: only move the max when it grows by more than 2^8
      warp_need = __any_sync(0xffffffffu, need);
      if (!warp_need) m_new = m_run;
      const float alpha = ex2_approx((m_run - m_new) * c);
      mc = m_new * c;
      m_run = m_new;
      return alpha;
    };
    auto rescale_o = [&](float alpha) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_o + ch * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_32x32b_x32(t_o + ch * 32, o);
      }
      tmem_st_wait();
    };
    // P^q row r -> 128B-swizzled K-major tile pair: 16-byte chunk c8 (8 halves) of sub-block sb lives at
    // sb*16KB + r*128 + ((c8 ^ (r&7)) << 4)
    auto store_p8 = [&](int col0, const float (&pv)[8]) {
      const int sb = col0 >> 6, c8 = (col0 & 63) >> 3;
      const uint4 val = make_uint4(pack_half2(pv[0], pv[1]), pack_half2(pv[2], pv[3]), pack_half2(pv[4], pv[5]),
                                   pack_half2(pv[6], pv[7]));
      *reinterpret_cast<uint4*>(sPq + sb * TILE_BYTES + r * 128 + ((c8 ^ (r & 7)) << 4)) = val;
    };

    if (q == 1 && n_kv > 2 && p.tune_skew > 0) {  // optionally start group 1 late (phase skew experiment)
      const long long t0 = clock64();
      while (clock64() - t0 < p.tune_skew) {}
    }
    for (int j = 0; j < n_kv; ++j) {
      const int nvalid = min(128, p.Lk - j * 128);
      mbar_wait(&s_full[q], j & 1);
      tc_fence_after();
      float alpha, mc;
      bool warp_need;
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      if (nvalid == 128) {
        // ---- full tile: S^q_j read once into 128 registers; Q K_{j+1}^T may overwrite S as soon as it is loaded
        uint32_t s[128];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(t_s + cc * 32, &s[cc * 32]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[q]);
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1])));
        }
        alpha = advance_max(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])), mc, warp_need);
        if (j > 0) {  // P^q_{j-1} V_{j-1} must be complete before P^q (smem) is overwritten / O rescaled
          mbar_wait(&o_done[q], (j - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int i0 = 0; i0 < 128; i0 += 8) {
          float pv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pv[e] = ex2_approx(fmaf(__uint_as_float(s[i0 + e]), c, -mc));
            rs4[e & 3] += pv[e];
          }
          store_p8(i0, pv);
        }
      } else {
        // ---- last, partial tile (key padding): two passes over TMEM through a 32-column buffer (register-light)
        float mx = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          if (cc * 32 < nvalid) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(t_s + cc * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, cc * 32 + i < nvalid ? __uint_as_float(t[i]) : -INFINITY);
          }
        }
        alpha = advance_max(mx, mc, warp_need);
        if (j > 0) {
          mbar_wait(&o_done[q], (j - 1) & 1);
          tc_fence_after();
        }
        const int ncols_w = (nvalid + 15) & ~15;   // the P.V MMA reads whole 16-column groups
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          if (cc * 32 < ncols_w) {
            uint32_t t[32];
            tmem_ld_32x32b_x32(t_s + cc * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (cc * 32 + g * 8 < ncols_w) {
                float pv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float pe = ex2_approx(fmaf(__uint_as_float(t[g * 8 + e]), c, -mc));
                  pv[e] = cc * 32 + g * 8 + e < nvalid ? pe : 0.f;
                  rs4[e & 3] += pv[e];
                }
                store_p8(cc * 32 + g * 8, pv);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&s_free[q]);
      }
      if (j > 0 && warp_need) rescale_o(alpha);   // rare (lazy rescale)
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[q]);
    }
    // ---------------- epilogue: O / l -> fp16 -> global ----------------
    mbar_wait(&o_done[q], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const long lq = static_cast<long>(qt) * 256 + q * 128 + r;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(t_o + ch * 32, o);
      tmem_ld_wait();
      if (lq < p.Lq) {
        __half* dst = p.out + (static_cast<long>(b) * p.Lq + lq) * p.o_ld + h * 64 + ch * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 val;
          val.x = pack_half2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          val.y = pack_half2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          val.z = pack_half2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          val.w = pack_half2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * i) = val;
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

}  // namespace

int attention(const AttnArgs& a, cudaStream_t stream) {
  M324_REQUIRE(a.q && a.k && a.v && a.out, "attention: null pointer");
  M324_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "attention: empty problem B=%d H=%d Lq=%d Lk=%d", a.B, a.H, a.Lq, a.Lk);
  M324_REQUIRE(a.q_ld % 8 == 0 && a.k_ld % 8 == 0 && a.v_ld % 8 == 0 && a.o_ld % 8 == 0, "attention: row strides must be multiples of 8");
  M324_REQUIRE(a.q_ld >= a.H * 64 && a.k_ld >= a.H * 64 && a.v_ld >= a.H * 64 && a.o_ld >= a.H * 64, "attention: row stride < H*64");
  M324_REQUIRE(a.q_batch_div >= 1, "attention: q_batch_div must be >= 1");
  M324_REQUIRE(a.q_rows >= (long)((a.B - 1) / a.q_batch_div) * a.q_batch_rows + a.Lq && a.kv_rows >= (long)(a.B - 1) * a.kv_batch_rows + a.Lk,
               "attention: q_rows / kv_rows smaller than the addressed range");
  static bool configured = false;
  if (!configured) {
    M324_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    configured = true;
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
  dim3 grid((a.Lq + 255) / 256, a.H, a.B);
  attn_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tq, tk, tv, a);
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

}  // namespace m324
