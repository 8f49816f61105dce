// tcgen05 flash-attention backward, head dim 64, non-causal:  given dO, the forward's log-sum-exp and D = rowsum(dO * O),
//   P = exp(scale * Q K^T - lse),  dV = P^T dO,  dP = dO V^T,  dS = scale * P * (dP - D),  dQ = dS K,  dK = dS^T Q
// This is the BwOp of xformers.ops.memory_efficient_attention as called by the reference (model/transformer.py:134-139,
// 209-214) -- what torch.autograd runs for every attention of the trainable trunk under train.py:157-170.
//
// One CTA = one (batch, head, 128-key K/V tile); it walks the 128-query tiles of that batch.  Everything is computed
// TRANSPOSED (rows = keys), so that the per-query statistics (lse, D) are per-COLUMN broadcasts and no row reduction
// is ever needed in the backward:
//   warp 0     : TMA producer (K_j, V_j once; Q_i, dO_i through a 2-stage ring)
//   warp 1     : tcgen05.mma issuer.  Per query tile i, all accumulators in TMEM (448 of 512 columns):
//                  S^T  = K_j Q_i^T          (128 x 128 x 64)        dP^T = V_j dO_i^T        (128 x 128 x 64)
//                  dV  += P^T dO_i           (128 x 64 x 128, B = dO MN-major straight from its TMA tile)
//                  dK  += dS^T Q_i           (128 x 64 x 128, B = Q  MN-major)
//                  dQ_i = dS K_j             (128 x 64 x 128, A = the dS^T tile read MN-major, B = K MN-major)
//   warps 4-11 : 256 threads, two per key row (64 query columns each): P^T and dS^T -> fp16, 128B-swizzled shared memory
//                (the K-major A operand of dV / dK and, read MN-major, of dQ); then the dQ_i tile leaves through a staging
//                tile and a TMA reduce-add (fp32 adds in L2: dQ accumulates over the K/V tiles, i.e. over CTAs).
// dK / dV of the tile are written once at the end (fp32).  Out-of-range queries / keys get P = 0.
//
// Software pipeline (the tensor pipe and the softmax threads never wait for each other in steady state):
//   issuer : S^T/dP^T(0) | for i: [S^T, dP^T of tile i+1 as soon as both of tile i are in registers] ,
//                                  [dV, dK, dQ of tile i when P^T_i / dS^T_i are in shared memory]
//   threads: for i: exp phase of tile i in registers (runs while dV/dK/dQ of tile i-1 execute) ,
//                   dQ_{i-1} out of TMEM -> staging -> TMA reduce-add ,  P^T_i -> TMEM ,  dS phase of tile i
// P^T (fp16) lives in TENSOR memory (64 spare columns) and is the A operand of dV straight from there (tcgen05.mma with A
// in TMEM): the shared-memory port is the binding resource of this kernel (ncu: LSU + tensor-core wavefronts 73 % of the
// data pipe), and this removes the P^T stores and the P^T operand reads from it.  dS^T stays in shared memory because dQ
// needs it transposed (MN-major A), which tensor memory cannot provide.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace m324 {
namespace {

constexpr int BWD_THREADS = 384;
constexpr int TILE = 128 * 64 * 2;             // 16 KB: 128 rows x 64 fp16
constexpr int BOFF_K = 0;
constexpr int BOFF_V = BOFF_K + TILE;
constexpr int BOFF_Q = BOFF_V + TILE;          // 2 stages
constexpr int BOFF_DO = BOFF_Q + 2 * TILE;     // 2 stages
constexpr int BOFF_STG = BOFF_DO + 2 * TILE;   // dQ staging: 8 warps x 32 x 32 fp32 = 32 KB
constexpr int BOFF_DST = BOFF_STG + 2 * TILE;  // dS^T [128 keys][128 queries] fp16 as two 16 KB sub-blocks of 64 queries
constexpr int BOFF_STAT = BOFF_DST + 2 * TILE; // [stage 2][lse 128 | D 128] floats
constexpr int BOFF_BAR = BOFF_STAT + 2048;
constexpr int BWD_SMEM = BOFF_BAR + 256 + 1024;
static_assert(BWD_SMEM <= 232448, "attention backward shared memory");
constexpr uint32_t TB_ST = 0, TB_DPT = 128, TB_DV = 256, TB_DK = 320, TB_DQ = 384, TB_PT = 448;   // P^T: 128 queries x fp16 = 64 columns
constexpr float LOG2E = 1.4426950408889634f;
// P and dS are tensor-core operands in fp16: with thousands of keys a normalised probability (1e-4) times a gradient
// (1e-3) would fall into fp16's subnormal range.  Both are therefore carried times 2^kPShift (P' = 2^12 P <= 4096,
// dS' = P' (dP - D) scale) and dV / dK / dQ are scaled back by 2^-12 in fp32 when they leave TMEM.
constexpr float kPShift = 12.0f;
constexpr float kPUnshift = 1.0f / 4096.0f;

// Warp-level timeline, profiling builds only (-DM324_TIMELINE=1; see csrc/attention.cu and scripts/attn_timeline.py --bwd)
#if defined(M324_TIMELINE) && M324_TIMELINE
__device__ unsigned long long* g_tlb_buf = nullptr;
__device__ int g_tlb_cta = -1, g_tlb_cap = 0;
enum : int { TLB_TOP = 1, TLB_S_READY, TLB_EXP_DONE, TLB_MMA_DONE_OK, TLB_PT_STORED, TLB_DQ_OUT, TLB_DP_LOADED, TLB_P_READY,
             TLB_S_WAIT = 16, TLB_S_ISSUED, TLB_G_WAIT, TLB_G_ISSUED };
__device__ __forceinline__ void tlb_record(int ev) {
  if (static_cast<int>(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) != g_tlb_cta || (threadIdx.x & 31) != 0 || g_tlb_buf == nullptr) return;
  const unsigned long long i = atomicAdd(g_tlb_buf, 1ull);
  if (i + 1 < static_cast<unsigned long long>(g_tlb_cap))
    g_tlb_buf[i + 1] = (static_cast<unsigned long long>(clock64()) << 16) | (static_cast<unsigned long long>(threadIdx.x >> 5) << 8) | ev;
}
#define TLB(ev) tlb_record(ev)
#else
#define TLB(ev) ((void)0)
#endif

__device__ __forceinline__ void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// fp16 row r, columns [col0, col0 + 8) of a [128][128] tile stored as two 128B-swizzled K-major sub-blocks
__device__ __forceinline__ void store_h8(uint32_t tile_row_addr, int col0, const float (&x)[8]) {
  const int sb = col0 >> 6, c8 = (col0 & 63) >> 3;
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((tile_row_addr ^ static_cast<uint32_t>(c8 << 4)) + sb * TILE),
               "r"(pack_half2(x[0], x[1])), "r"(pack_half2(x[2], x[3])), "r"(pack_half2(x[4], x[5])), "r"(pack_half2(x[6], x[7]))
               : "memory");
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const AttnBwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BOFF_BAR);
  uint64_t* kv_full = bars;        // 1
  uint64_t* q_full = bars + 1;     // 2
  uint64_t* q_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;     // S^T and dP^T of tile i are in TMEM
  uint64_t* s_free = bars + 6;     // 256: both have been read into registers
  uint64_t* p_ready = bars + 7;    // 256: P^T and dS^T of tile i are in shared memory
  uint64_t* mma_done = bars + 8;   // dV / dK / dQ MMAs of tile i completed: P^T / dS^T reusable, dQ_i readable
  uint64_t* dq_free = bars + 9;    // 256: dQ_i has been read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.Lq + 127) / 128;
  const long q_row0 = static_cast<long>(b / p.q_batch_div) * p.q_batch_rows;     // row of query 0 in q / dQ
  const long o_row0 = static_cast<long>(b) * p.Lq;                               // row of query 0 in dO / lse / D
  const long kv_row0 = static_cast<long>(b) * p.kv_batch_rows + static_cast<long>(jt) * 128;
  const int nvalid_k = min(128, p.Lk - jt * 128);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_ready, 256);
    mbar_init(mma_done, 1);
    mbar_init(dq_free, 256);
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
      mbar_expect_tx(kv_full, 2 * TILE);
      tma_load_2d(smem + BOFF_K, &tmK, kv_full, h * 64, static_cast<int>(kv_row0));
      tma_load_2d(smem + BOFF_V, &tmV, kv_full, h * 64, static_cast<int>(kv_row0));
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * TILE);
        tma_load_2d(smem + BOFF_Q + st * TILE, &tmQ, &q_full[st], h * 64, static_cast<int>(q_row0 + i * 128));
        tma_load_2d(smem + BOFF_DO + st * TILE, &tmDO, &q_full[st], h * 64, static_cast<int>(o_row0 + i * 128));
      }
    }
  } else if (warp == 1) {
    const uint32_t id_s = umma_idesc_f16_ex(128, 128, false, false, false, false);   // K-major x K-major
    const uint32_t id_kv = umma_idesc_f16_ex(128, 64, false, false, false, true);    // A K-major (P^T / dS^T), B MN-major (dO / Q)
    const uint32_t id_dq = umma_idesc_f16_ex(128, 64, false, false, true, true);     // A = dS^T read MN-major, B = K MN-major
    const uint64_t dK_ = umma_desc_sw128(smem_u32(smem + BOFF_K), 16, 1024), dV_ = umma_desc_sw128(smem_u32(smem + BOFF_V), 16, 1024);
    const uint64_t dQ0 = umma_desc_sw128(smem_u32(smem + BOFF_Q), 16, 1024), dDO0 = umma_desc_sw128(smem_u32(smem + BOFF_DO), 16, 1024);
    const uint64_t dDST = umma_desc_sw128(smem_u32(smem + BOFF_DST), 16, 1024);
    const uint64_t dDST_mn = umma_desc_sw128(smem_u32(smem + BOFF_DST), TILE, 1024);   // MN atoms (64 queries) 16 KB apart
    constexpr uint64_t kTile = TILE >> 4;
    mbar_wait(kv_full, 0);
    auto issue_s = [&](int i) {     // S^T and dP^T of query tile i
      const int st = i & 1;
      TLB(TLB_S_WAIT);
      mbar_wait(&q_full[st], (i >> 1) & 1);
      if (i > 0) mbar_wait(s_free, (i - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + TB_ST, dK_ + 2 * k, dQ0 + st * kTile + 2 * k, id_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + TB_DPT, dV_ + 2 * k, dDO0 + st * kTile + 2 * k, id_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      TLB(TLB_S_ISSUED);
    };
    issue_s(0);
    for (int i = 0; i < n_q; ++i) {
      const int st = i & 1;
      if (i + 1 < n_q) issue_s(i + 1);
      TLB(TLB_G_WAIT);
      mbar_wait(p_ready, i & 1);
      if (i > 0) mbar_wait(dq_free, (i - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // contraction over the 128 queries, 16 at a time
          const uint64_t a_off = (kk >> 2) * kTile + (kk & 3) * 2, b_off = kk * (2048 >> 4);
          umma_f16_ts(tmem_base + TB_DV, tmem_base + TB_PT + kk * 8, dDO0 + st * kTile + b_off, id_kv, (i > 0 || kk > 0) ? 1u : 0u);
          umma_f16_ss(tmem_base + TB_DK, dDST + a_off, dQ0 + st * kTile + b_off, id_kv, (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)     // contraction over the 128 keys
          umma_f16_ss(tmem_base + TB_DQ, dDST_mn + kk * (2048 >> 4), dK_ + kk * (2048 >> 4), id_dq, kk > 0 ? 1u : 0u);
        umma_commit(mma_done);
        umma_commit(&q_empty[st]);
      }
      __syncwarp();
      TLB(TLB_G_ISSUED);
    }
  } else if (warp >= 4) {
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;                       // key row of this thread (S^T / dP^T / dV / dK); query row for dQ
    const uint32_t t_lane = static_cast<uint32_t>(quarter * 32) << 16;
    const int ctid = threadIdx.x - 128;                      // 0..255
    float* stat = reinterpret_cast<float*>(smem + BOFF_STAT);
    const uint32_t stat_addr = smem_u32(smem + BOFF_STAT);
    const uint32_t stg_addr = smem_u32(smem + BOFF_STG) + (warp - 4) * 4096;
    const uint32_t dst_row = smem_u32(smem + BOFF_DST) + r * 128 + ((r & 7) << 4);
    const float c = p.scale * LOG2E;
    const bool key_ok = r < nvalid_k;
    // statistic of query (ctid & 127): lse for the first 128 threads, D for the others; prefetched one tile ahead
    auto load_stat = [&](int i) -> float {
      const int qi = i * 128 + (ctid & 127);
      if (qi >= p.Lq) return 0.f;
      return ctid < 128 ? p.lse[(o_row0 + qi) * p.lse_ld + h] - kPShift : p.D[(o_row0 + qi) * p.d_ld + h];
    };
    // dQ of query tile i: TMEM -> (x 2^-12) -> staging in the idle P^T buffer -> TMA reduce-add into global dQ
    auto dq_out = [&](int i) {
      uint32_t dq[32];
      tmem_ld_32x32b_x32(tmem_base + t_lane + TB_DQ + half * 32, dq);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);
      if (lane == 0) tma_store_wait_read();     // the previous box of this warp has left its staging tile
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg_addr + lane * 128 + ((j ^ (lane & 7)) << 4)),
                     "f"(__uint_as_float(dq[4 * j]) * kPUnshift), "f"(__uint_as_float(dq[4 * j + 1]) * kPUnshift),
                     "f"(__uint_as_float(dq[4 * j + 2]) * kPUnshift), "f"(__uint_as_float(dq[4 * j + 3]) * kPUnshift)
                     : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && !(p.tune & 1)) {
        tma_reduce_add_2d(&tmDQ, smem + BOFF_STG + (warp - 4) * 4096, h * 64 + half * 32, static_cast<int>(q_row0 + i * 128 + quarter * 32));
        tma_store_commit();
      }
    };
    auto lds4 = [&](uint32_t addr) -> float4 {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
      return v;
    };
    float stat_next = load_stat(0);
    for (int i = 0; i < n_q; ++i) {
      float* st_lse = stat + (i & 1) * 256;
      TLB(TLB_TOP);
      st_lse[ctid] = stat_next;                             // [0,128) lse, [128,256) D
      bar_sync_named(1, 256);
      if (i + 1 < n_q) stat_next = load_stat(i + 1);
      const int q0 = i * 128 + half * 64;                   // first query column of this thread
      const uint32_t lse_addr = stat_addr + ((i & 1) * 256 + half * 64) * 4, d_addr = lse_addr + 512;
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      TLB(TLB_S_READY);
      float pv[64];
      {
        uint32_t* pu = reinterpret_cast<uint32_t*>(pv);
        tmem_ld_32x32b_x32(tmem_base + t_lane + TB_ST + half * 64, pu);
        tmem_ld_32x32b_x32(tmem_base + t_lane + TB_ST + half * 64 + 32, pu + 32);
        tmem_ld_wait();
      }
      uint32_t ppk[32];                                     // P^T of this thread's 64 queries, packed fp16 pairs
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 la = lds4(lse_addr + g * 32), lb = lds4(lse_addr + g * 32 + 16);
        const float l8[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = g * 8 + e;
          const float xx = fmaf(pv[j], c, -l8[e]);
          const float pe = (p.tune & 2) ? xx : ex2_approx(xx);
          pv[j] = (key_ok && q0 + j < p.Lq) ? pe : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) ppk[g * 4 + e] = pack_half2(pv[g * 8 + 2 * e], pv[g * 8 + 2 * e + 1]);
      }
      TLB(TLB_EXP_DONE);
      if (i > 0) {                   // dV / dK / dQ of tile i-1 have run behind the exponentials above: P^T (TMEM) is free, dQ ready
        mbar_wait(mma_done, (i - 1) & 1);
        tc_fence_after();
      }
      TLB(TLB_MMA_DONE_OK);
#pragma unroll
      for (int g = 0; g < 4; ++g) tmem_st_32x32b_x8(tmem_base + t_lane + TB_PT + half * 32 + g * 8, &ppk[g * 8]);
      TLB(TLB_PT_STORED);
      if (i > 0) dq_out(i - 1);
      TLB(TLB_DQ_OUT);
      {
        uint32_t dp[64];
        tmem_ld_32x32b_x32(tmem_base + t_lane + TB_DPT + half * 64, dp);
        tmem_ld_32x32b_x32(tmem_base + t_lane + TB_DPT + half * 64 + 32, dp + 32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free);
        TLB(TLB_DP_LOADED);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float x[8];
          const float4 da = lds4(d_addr + g * 32), db = lds4(d_addr + g * 32 + 16);
          const float d8[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = g * 8 + e;
            x[e] = pv[j] * (__uint_as_float(dp[j]) - d8[e]) * p.scale;
          }
          store_h8(dst_row, half * 64 + g * 8, x);
        }
      }
      fence_proxy_async_smem();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      TLB(TLB_P_READY);
    }
    mbar_wait(mma_done, (n_q - 1) & 1);
    tc_fence_after();
    dq_out(n_q - 1);
    // dK / dV of this K/V tile: row per thread, 32 columns per thread and tensor
    {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_base + t_lane + TB_DV + half * 32, o);
      tmem_ld_wait();
      if (key_ok) {
        float4* dst = reinterpret_cast<float4*>(p.dV + (kv_row0 + r) * p.dv_ld + h * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(o[4 * j]) * kPUnshift, __uint_as_float(o[4 * j + 1]) * kPUnshift,
                               __uint_as_float(o[4 * j + 2]) * kPUnshift, __uint_as_float(o[4 * j + 3]) * kPUnshift);
      }
      tmem_ld_32x32b_x32(tmem_base + t_lane + TB_DK + half * 32, o);
      tmem_ld_wait();
      if (key_ok) {
        float4* dst = reinterpret_cast<float4*>(p.dK + (kv_row0 + r) * p.dk_ld + h * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(o[4 * j]) * kPUnshift, __uint_as_float(o[4 * j + 1]) * kPUnshift,
                               __uint_as_float(o[4 * j + 2]) * kPUnshift, __uint_as_float(o[4 * j + 3]) * kPUnshift);
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int attention_bwd(const AttnBwdArgs& a, cudaStream_t stream) {
  M324_REQUIRE(a.q && a.k && a.v && a.dO && a.lse && a.D && a.dQ && a.dK && a.dV, "attention_bwd: null pointer");
  M324_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0, "attention_bwd: empty problem");
  M324_REQUIRE(a.q_ld % 8 == 0 && a.k_ld % 8 == 0 && a.v_ld % 8 == 0 && a.do_ld % 8 == 0, "attention_bwd: f16 row strides must be multiples of 8");
  M324_REQUIRE(a.dq_ld % 4 == 0 && a.dk_ld % 4 == 0 && a.dv_ld % 4 == 0, "attention_bwd: fp32 row strides must be multiples of 4");
  M324_REQUIRE((reinterpret_cast<uintptr_t>(a.dQ) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dK) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(a.dV) & 15) == 0, "attention_bwd: gradient buffers must be 16-byte aligned");
  M324_REQUIRE(a.q_batch_div >= 1, "attention_bwd: q_batch_div must be >= 1");
  M324_REQUIRE(a.kv_batch_rows != 0 || a.B == 1, "attention_bwd: a K/V operand shared by several batches is not supported");
  M324_REQUIRE(a.q_rows >= (long)((a.B - 1) / a.q_batch_div) * a.q_batch_rows + a.Lq && a.kv_rows >= (long)(a.B - 1) * a.kv_batch_rows + a.Lk,
               "attention_bwd: q_rows / kv_rows smaller than the addressed range");
  static PerDeviceOnce configured;
  if (configured.need()) {
    M324_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    configured.mark();
  }
  CUtensorMap tq, tk, tv, tdo, tdq;
  uint32_t box[2] = {64, 128};
  {
    uint64_t dims[2] = {static_cast<uint64_t>(a.H) * 64, static_cast<uint64_t>(a.q_rows)};
    uint64_t str[1] = {static_cast<uint64_t>(a.q_ld) * 2};
    int e = make_tmap_16b(&tq, a.q, 2, dims, str, box);
    if (e) return e;
    uint64_t strq[1] = {static_cast<uint64_t>(a.dq_ld) * 4};
    uint32_t box32[2] = {32, 32};
    e = make_tmap_f32(&tdq, a.dQ, 2, dims, strq, box32);
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
  {
    uint64_t dims[2] = {static_cast<uint64_t>(a.H) * 64, static_cast<uint64_t>(a.B) * a.Lq};
    uint64_t str[1] = {static_cast<uint64_t>(a.do_ld) * 2};
    int e = make_tmap_16b(&tdo, a.dO, 2, dims, str, box);
    if (e) return e;
  }
  dim3 grid((a.Lk + 127) / 128, a.H, a.B);
  M324_CUDA(launch_pdl(attn_bwd_kernel, grid, dim3(BWD_THREADS), BWD_SMEM, stream, tq, tk, tv, tdo, tdq, a));
  M324_CUDA(cudaGetLastError());
  return M324_OK;
}

#if defined(M324_TIMELINE) && M324_TIMELINE
// Profiling builds only: timeline buffer (word 0 = record counter) and the linear CTA index to trace, for attn_bwd_kernel.
extern "C" int m324_timeline_set_bwd(void* buf, int cap, int cta) {
  unsigned long long* b = static_cast<unsigned long long*>(buf);
  if (cudaMemcpyToSymbol(g_tlb_buf, &b, sizeof(b)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_tlb_cap, &cap, sizeof(cap)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_tlb_cta, &cta, sizeof(cta)) != cudaSuccess) return -1;
  return 0;
}
#endif

}  // namespace m324
