// Fused attention backward for sm_100a (head_dim 128, non-causal), three kernels:
//
//   attn_bwd_delta    D[b,h,q] = sum_d O*dO                       (HBM-bound, one warp per (token, head))
//   attn_bwd_dkv      one CTA per 128-key tile, loops over query tiles:
//                        St  = K Q^T          (SS)     -> TMEM            Pt  = exp2(St*c - LSE[q])
//                        dPt = V dO^T         (SS)     -> TMEM            dSt = Pt * (dPt - D[q])
//                        dV += Pt  dO         (TS: Pt / dSt are read from TMEM, dO / Q tiles re-used MN-major)
//                        dK += dSt Q
//   attn_bwd_dq       one CTA per 128-query tile, loops over key tiles:
//                        S = Q K^T, dP = dO V^T (SS) ; dS = P*(dP - D) ; dQ += dS K (TS, K tile re-used MN-major)
//
// Splitting dQ from dK/dV recomputes S and dP once more (7 instead of 5 tile GEMMs) but needs no atomics and no fp32
// dQ staging buffer.  The transposed formulation in dkv makes P^T / dS^T land row-per-thread in TMEM so they can feed
// tcgen05.mma as the A operand without touching shared memory.
//
// In both kernels 8 softmax warps share a tile: warp w owns TMEM lane quarter (w & 3) and column half (w-4)/4.
// bf16 P / dS for columns [64h, 64h+64) are written over columns [64h, 64h+32) of the fp32 tile they came from, so each
// warp only overwrites columns it has already read.
//
// LSE is the log2-domain logsumexp written by attn_fwd.  Outputs dQ/dK/dV are head-major [B,H,L,128] bf16.
// Replaces flash-attn / SDPA backward reached through autograd from models/flux.py:502,525.
#include <stdlib.h>

#include "host_util.h"
#include "sm100_common.cuh"

namespace dpipe {

constexpr int AB_THREADS = 384;
constexpr int BT = 128;  // tile edge (queries and keys)
constexpr int BHALF = 128 * 64 * 2;
constexpr int BTILE = 2 * BHALF;  // 32 KiB
constexpr int AB_STATS_BYTES = 4096;   // v1/v2: [2][lse 128 | D 128] floats; v3: 4-deep ring of the same
constexpr int AB_SMEM_BYTES = 6 * BTILE + AB_STATS_BYTES + 1024 + 256;

struct AttnBwdParams {
  const float* lse;    // [B,H,Lq] log2 domain
  const float* delta;  // [B,H,Lq]
  __nv_bfloat16* dq;   // [B,H,Lq,128]
  __nv_bfloat16* dk;   // [B,H,Lk,128]
  __nv_bfloat16* dv;   // [B,H,Lk,128]
  int batch, heads, seq_q, seq_k;
  float scale_log2, scale;
  const __nv_bfloat16* k;     // [B,H,Lk,128]  (v3 dK/dV kernel: K rows go straight from global memory into TMEM)
  const __nv_bfloat16* q;     // [B,H,Lq,128]  (v4 dQ kernel: likewise)
  const __nv_bfloat16* d_o;   // [B*Lq, lddo] token-major
  int64_t lddo;
};

// K-major operand tile [128 rows][128 d] as two 64-wide halves: byte offset of k-step `k` (16 d per step)
__device__ __forceinline__ uint32_t kmajor_off(int k) { return (k >> 2) * BHALF + (k & 3) * 32; }
// TMEM column of the packed bf16 pair for logical column c (see header comment)
__device__ __forceinline__ uint32_t packed_col(int kstep) { return (kstep < 4) ? kstep * 8 : 64 + (kstep - 4) * 8; }

// ---------------------------------------------------------------------------------------------
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t ldo,
                                      const __nv_bfloat16* __restrict__ d_o, int64_t lddo, float* __restrict__ delta,
                                      int batch, int heads, int seq) {
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t total = (int64_t)batch * seq * heads;
  if (gw >= total) return;
  const int h = gw % heads;
  const int64_t tok = gw / heads;  // b*seq + q
  const uint2 a = *reinterpret_cast<const uint2*>(o + tok * ldo + h * 128 + lane * 4);
  const uint2 b = *reinterpret_cast<const uint2*>(d_o + tok * lddo + h * 128 + lane * 4);
  float s = bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) +
            bf16_hi(a.y) * bf16_hi(b.y);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const int bb = tok / seq, q = tok % seq;
    delta[((int64_t)bb * heads + h) * seq + q] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// dK / dV
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                    const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                    const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* k_smem = smem;
  uint8_t* v_smem = smem + BTILE;
  uint8_t* q_smem = smem + 2 * BTILE;   // 2 stages
  uint8_t* do_smem = smem + 4 * BTILE;  // 2 stages
  float* lse_smem = reinterpret_cast<float*>(smem + 6 * BTILE);  // [2][128]
  float* dl_smem = lse_smem + 256;                               // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * BTILE + AB_STATS_BYTES);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* do_full = bars + 5;    // [2]
  uint64_t* do_empty = bars + 7;   // [2]
  uint64_t* st_full = bars + 9;    // MMA -> softmax
  uint64_t* dpt_full = bars + 10;  // MMA -> softmax
  uint64_t* p_ready = bars + 11;   // softmax -> MMA (8 warps)
  uint64_t* ds_ready = bars + 12;  // softmax -> MMA (8 warps)
  uint64_t* acc_done = bars + 13;  // MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int kv0 = blockIdx.x * BT;
  const int head = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int nq = (p.seq_q + BT - 1) / BT;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(smem_u32(kv_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&q_full[s]), 1); mbar_init(smem_u32(&q_empty[s]), 1);
      mbar_init(smem_u32(&do_full[s]), 1); mbar_init(smem_u32(&do_empty[s]), 1);
    }
    mbar_init(smem_u32(st_full), 1); mbar_init(smem_u32(dpt_full), 1);
    mbar_init(smem_u32(p_ready), 8); mbar_init(smem_u32(ds_ready), 8);
    mbar_init(smem_u32(acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // TMEM: St [0,128)  dPt [128,256)  dV [256,384)  dK [384,512)
  const uint32_t T_ST = tmem_base, T_DPT = tmem_base + 128, T_DV = tmem_base + 256, T_DK = tmem_base + 384;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t kvb = smem_u32(kv_full);
      mbar_expect_tx(kvb, 2 * BTILE);
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(&tma_k, kvb, smem_u32(k_smem + h * BHALF), h * 64, kv0, bh, kEvictFirst);
        tma_load_3d(&tma_v, kvb, smem_u32(v_smem + h * BHALF), h * 64, kv0, bh, kEvictFirst);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nq; ++j) {
        const int q0 = j * BT;
        mbar_wait(smem_u32(&q_empty[stage]), phase ^ 1);
        const uint32_t qb = smem_u32(&q_full[stage]);
        mbar_expect_tx(qb, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_q, qb, smem_u32(q_smem + stage * BTILE + h * BHALF), h * 64, q0, bh, kEvictLast);
        mbar_wait(smem_u32(&do_empty[stage]), phase ^ 1);
        const uint32_t db = smem_u32(&do_full[stage]);
        mbar_expect_tx(db, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_4d(&tma_do, db, smem_u32(do_smem + stage * BTILE + h * BHALF), h * 64, q0, head, b, kEvictLast);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_ss = make_idesc_bf16(BT, BT, false, false);
      constexpr uint32_t idesc_ts = make_idesc_bf16(BT, 128, false, true);
      mbar_wait(smem_u32(kv_full), 0);
      const uint32_t kb = smem_u32(k_smem), vb = smem_u32(v_smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nq; ++j) {
        const uint32_t qb = smem_u32(q_smem + stage * BTILE), dob = smem_u32(do_smem + stage * BTILE);
        mbar_wait(smem_u32(&q_full[stage]), phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // St[kv, q] = K Q^T
          umma_ss<1>(T_ST, make_smem_desc(kb + kmajor_off(k), 16, 1024), make_smem_desc(qb + kmajor_off(k), 16, 1024),
                     idesc_ss, k != 0);
        umma_commit<1>(smem_u32(st_full));
        mbar_wait(smem_u32(&do_full[stage]), phase);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dPt[kv, q] = V dO^T
          umma_ss<1>(T_DPT, make_smem_desc(vb + kmajor_off(k), 16, 1024), make_smem_desc(dob + kmajor_off(k), 16, 1024),
                     idesc_ss, k != 0);
        umma_commit<1>(smem_u32(dpt_full));
        mbar_wait(smem_u32(p_ready), j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV[kv, d] += Pt[kv, q] dO[q, d]   (dO tile read MN-major: 16 queries per step)
          umma_ts(T_DV, T_ST + packed_col(k), make_smem_desc(dob + k * 2048, BHALF, 1024), idesc_ts, (j | k) != 0);
        umma_commit<1>(smem_u32(&do_empty[stage]));
        mbar_wait(smem_u32(ds_ready), j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK[kv, d] += dSt[kv, q] Q[q, d]
          umma_ts(T_DK, T_DPT + packed_col(k), make_smem_desc(qb + k * 2048, BHALF, 1024), idesc_ts, (j | k) != 0);
        umma_commit<1>(smem_u32(&q_empty[stage]));
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
      umma_commit<1>(smem_u32(acc_done));
    }
  } else if (warp >= 4) {
    const uint32_t quad = warp & 3;
    const int half = (warp - 4) >> 2;          // column half of the tile this warp handles
    const int sm_tid = (warp - 4) * 32 + lane;  // 0..255 among softmax threads
    const uint32_t lane_base = (quad * 32u) << 16;
    const float c = p.scale_log2;
    for (int j = 0; j < nq; ++j) {
      const int q0 = j * BT;
      const int st = j & 1;
      // stage LSE / D of this query tile (invalid queries get +inf so their probabilities vanish)
      {
        const int t = sm_tid & 127;
        const int q = q0 + t;
        if (sm_tid < 128) lse_smem[st * 128 + t] = (q < p.seq_q) ? p.lse[(int64_t)bh * p.seq_q + q] : INFINITY;
        else dl_smem[st * 128 + t] = (q < p.seq_q) ? p.delta[(int64_t)bh * p.seq_q + q] : 0.f;
      }
      named_bar_sync(1, 256);
      const float* lse_s = lse_smem + st * 128 + half * 64;
      const float* dl_s = dl_smem + st * 128 + half * 64;
      float pv[64];
      mbar_wait(smem_u32(st_full), j & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32], pk[16];
        tmem_ld_x32(T_ST + lane_base + half * 64 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(r[x]), c, -lse_s[cc * 32 + x]));
          const float p1 = fast_exp2(fmaf(__uint_as_float(r[x + 1]), c, -lse_s[cc * 32 + x + 1]));
          pv[cc * 32 + x] = p0; pv[cc * 32 + x + 1] = p1;
          pk[x >> 1] = pack_bf16(p0, p1);
        }
        tmem_st_x16(T_ST + lane_base + half * 64 + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(p_ready));
      mbar_wait(smem_u32(dpt_full), j & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32], pk[16];
        tmem_ld_x32(T_DPT + lane_base + half * 64 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const float d0 = pv[cc * 32 + x] * (__uint_as_float(r[x]) - dl_s[cc * 32 + x]);
          const float d1 = pv[cc * 32 + x + 1] * (__uint_as_float(r[x + 1]) - dl_s[cc * 32 + x + 1]);
          pk[x >> 1] = pack_bf16(d0, d1);
        }
        tmem_st_x16(T_DPT + lane_base + half * 64 + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(ds_ready));
    }
    // epilogue: each thread stores 64 columns of its key row of dV and dK
    mbar_wait(smem_u32(acc_done), 0);
    tc_fence_after();
    const int kv = kv0 + quad * 32 + lane;
    const bool ok = kv < p.seq_k;
    const int64_t o = ((int64_t)bh * p.seq_k + kv) * 128 + half * 64;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t t = (which == 0 ? T_DV : T_DK) + lane_base + half * 64;
      __nv_bfloat16* dst = (which == 0 ? p.dv : p.dk) + o;
      const float mul = which == 0 ? 1.0f : p.scale;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld_x32(t + cc * 32, r);
        tmem_ld_wait();
        if (ok) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            uint4 qv;
            qv.x = pack_bf16(__uint_as_float(r[x * 8 + 0]) * mul, __uint_as_float(r[x * 8 + 1]) * mul);
            qv.y = pack_bf16(__uint_as_float(r[x * 8 + 2]) * mul, __uint_as_float(r[x * 8 + 3]) * mul);
            qv.z = pack_bf16(__uint_as_float(r[x * 8 + 4]) * mul, __uint_as_float(r[x * 8 + 5]) * mul);
            qv.w = pack_bf16(__uint_as_float(r[x * 8 + 6]) * mul, __uint_as_float(r[x * 8 + 7]) * mul);
            d4[x] = qv;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                   const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                   const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_smem = smem;
  uint8_t* do_smem = smem + BTILE;
  uint8_t* k_smem = smem + 2 * BTILE;  // 2 stages
  uint8_t* v_smem = smem + 4 * BTILE;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * BTILE + AB_STATS_BYTES);
  uint64_t* qdo_full = bars;       // [1]
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;     // [2] double-buffered S
  uint64_t* dp_full = bars + 11;
  uint64_t* ds_ready = bars + 12;  // 8 warps
  uint64_t* acc_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * BT;
  const int head = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int nkv = (p.seq_k + BT - 1) / BT;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(smem_u32(qdo_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1); mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1); mbar_init(smem_u32(&v_empty[s]), 1);
      mbar_init(smem_u32(&s_full[s]), 1);
    }
    mbar_init(smem_u32(dp_full), 1);
    mbar_init(smem_u32(ds_ready), 8);
    mbar_init(smem_u32(acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // TMEM: S0 [0,128) S1 [128,256) dP [256,384) dQ [384,512)
  const uint32_t T_DP = tmem_base + 256, T_DQ = tmem_base + 384;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t qb = smem_u32(qdo_full);
      mbar_expect_tx(qb, 2 * BTILE);
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(&tma_q, qb, smem_u32(q_smem + h * BHALF), h * 64, q0, bh, kEvictFirst);
        tma_load_4d(&tma_do, qb, smem_u32(do_smem + h * BHALF), h * 64, q0, head, b, kEvictFirst);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        const int kv0 = j * BT;
        mbar_wait(smem_u32(&k_empty[stage]), phase ^ 1);
        const uint32_t kb = smem_u32(&k_full[stage]);
        mbar_expect_tx(kb, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_k, kb, smem_u32(k_smem + stage * BTILE + h * BHALF), h * 64, kv0, bh, kEvictLast);
        mbar_wait(smem_u32(&v_empty[stage]), phase ^ 1);
        const uint32_t vb = smem_u32(&v_full[stage]);
        mbar_expect_tx(vb, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_v, vb, smem_u32(v_smem + stage * BTILE + h * BHALF), h * 64, kv0, bh, kEvictLast);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_ss = make_idesc_bf16(BT, BT, false, false);
      constexpr uint32_t idesc_ts = make_idesc_bf16(BT, 128, false, true);
      const uint32_t qb = smem_u32(q_smem), dob = smem_u32(do_smem);
      auto issue_s = [&](int jj, int stg) {   // S_{jj&1}[q, kv] = Q K^T
        const uint32_t kb = smem_u32(k_smem + stg * BTILE);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(tmem_base + (jj & 1) * 128, make_smem_desc(qb + kmajor_off(k), 16, 1024),
                     make_smem_desc(kb + kmajor_off(k), 16, 1024), idesc_ss, k != 0);
        umma_commit<1>(smem_u32(&s_full[jj & 1]));
      };
      auto issue_dp = [&](int stg) {          // dP[q, kv] = dO V^T
        const uint32_t vb = smem_u32(v_smem + stg * BTILE);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_DP, make_smem_desc(dob + kmajor_off(k), 16, 1024), make_smem_desc(vb + kmajor_off(k), 16, 1024),
                     idesc_ss, k != 0);
        umma_commit<1>(smem_u32(dp_full));
      };
      mbar_wait(smem_u32(qdo_full), 0);
      mbar_wait(smem_u32(&k_full[0]), 0);
      tc_fence_after();
      issue_s(0, 0);
      mbar_wait(smem_u32(&v_full[0]), 0);
      tc_fence_after();
      issue_dp(0);
      umma_commit<1>(smem_u32(&v_empty[0]));
      for (int j = 0; j < nkv; ++j) {
        const int stg = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const bool more = j + 1 < nkv;
        const int nstg = (j + 1) & 1;
        const uint32_t nph = ((j + 1) >> 1) & 1;
        if (more) {   // S(j+1) into the other S buffer while the softmax warps work on tile j
          mbar_wait(smem_u32(&k_full[nstg]), nph);
          tc_fence_after();
          issue_s(j + 1, nstg);
        }
        mbar_wait(smem_u32(ds_ready), j & 1);
        tc_fence_after();
        const uint32_t kb = smem_u32(k_smem + stg * BTILE);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ[q, d] += dS[q, kv] K[kv, d]   (K tile read MN-major)
          umma_ts(T_DQ, T_DP + packed_col(k), make_smem_desc(kb + k * 2048, BHALF, 1024), idesc_ts, (j | k) != 0);
        umma_commit<1>(smem_u32(&k_empty[stg]));
        (void)ph;
        if (more) {
          mbar_wait(smem_u32(&v_full[nstg]), nph);
          tc_fence_after();
          issue_dp(nstg);
          umma_commit<1>(smem_u32(&v_empty[nstg]));
        }
      }
      umma_commit<1>(smem_u32(acc_done));
    }
  } else if (warp >= 4) {
    const uint32_t quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t lane_base = (quad * 32u) << 16;
    const int q = q0 + quad * 32 + lane;
    const bool ok = q < p.seq_q;
    const float lse = ok ? p.lse[(int64_t)bh * p.seq_q + q] : INFINITY;
    const float dl = ok ? p.delta[(int64_t)bh * p.seq_q + q] : 0.f;
    const float c = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int valid = min(BT, p.seq_k - j * BT) - half * 64;  // valid keys within this warp's 64 columns
      float pv[64];
      mbar_wait(smem_u32(&s_full[j & 1]), (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld_x32(tmem_base + (j & 1) * 128 + lane_base + half * 64 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          float pp = fast_exp2(fmaf(__uint_as_float(r[x]), c, -lse));
          pv[cc * 32 + x] = (cc * 32 + x < valid) ? pp : 0.f;
        }
      }
      mbar_wait(smem_u32(dp_full), j & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32], pk[16];
        tmem_ld_x32(T_DP + lane_base + half * 64 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 2)
          pk[x >> 1] = pack_bf16(pv[cc * 32 + x] * (__uint_as_float(r[x]) - dl),
                                 pv[cc * 32 + x + 1] * (__uint_as_float(r[x + 1]) - dl));
        tmem_st_x16(T_DP + lane_base + half * 64 + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(ds_ready));
    }
    mbar_wait(smem_u32(acc_done), 0);
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + ((int64_t)bh * p.seq_q + q) * 128 + half * 64;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t r[32];
      tmem_ld_x32(T_DQ + lane_base + half * 64 + cc * 32, r);
      tmem_ld_wait();
      if (ok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 qv;
          qv.x = pack_bf16(__uint_as_float(r[x * 8 + 0]) * p.scale, __uint_as_float(r[x * 8 + 1]) * p.scale);
          qv.y = pack_bf16(__uint_as_float(r[x * 8 + 2]) * p.scale, __uint_as_float(r[x * 8 + 3]) * p.scale);
          qv.z = pack_bf16(__uint_as_float(r[x * 8 + 4]) * p.scale, __uint_as_float(r[x * 8 + 5]) * p.scale);
          qv.w = pack_bf16(__uint_as_float(r[x * 8 + 6]) * p.scale, __uint_as_float(r[x * 8 + 7]) * p.scale);
          d4[x] = qv;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// =============================================================================================
// v2: software-pipelined variants.  The 128-wide score tile is processed as two 64-wide sub-tiles with double-buffered
// S / dP accumulators in TMEM, so the tensor core computes the scores of sub-tile t+1 while the softmax warps turn
// sub-tile t into P / dS, and the dV / dK (or dQ) MMAs of sub-tile t run while the softmax warps are already on t+1.
//   dkv TMEM: St[2] 2x64 | dPt[2] 2x64 | dV 128 | dK 128          dq TMEM: S[2] 2x64 | dP[2] 2x64 | dQ 128
// bf16 P / dS for the 32 columns a warp owns are written over the first 16 of those same columns.
// =============================================================================================

// STATS_WARPS: the two otherwise idle control warps (2, 3) stream LSE / D of every query tile into a 4-deep shared-memory
// ring ahead of the softmax warpgroups (mbarrier full/empty), so the softmax loop holds no global load and no named barrier
// (r01b capture: `barrier` 2.47 + `long_scoreboard` 2.33 warps per issue were the top stalls of this kernel).
template <bool STATS_WARPS>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                     const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                     const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* k_smem = smem;
  uint8_t* v_smem = smem + BTILE;
  uint8_t* q_smem = smem + 2 * BTILE;   // 2 stages of [128 q][128 d]
  uint8_t* do_smem = smem + 4 * BTILE;  // 2 stages
  float* lse_smem = reinterpret_cast<float*>(smem + 6 * BTILE);  // !STATS_WARPS: [2][128]; STATS_WARPS: ring [4][lse 128 | D 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * BTILE + AB_STATS_BYTES);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* do_full = bars + 5;    // [2]
  uint64_t* do_empty = bars + 7;   // [2]
  uint64_t* sd_full = bars + 9;    // [2] MMA -> softmax: St and dPt of buffer b complete
  uint64_t* p_ready = bars + 11;   // [2] softmax -> MMA (8 warps)
  uint64_t* ds_ready = bars + 13;  // [2] softmax -> MMA (8 warps)
  uint64_t* acc_done = bars + 15;
  uint64_t* stats_full = bars + 16;   // [4] loader warps -> softmax
  uint64_t* stats_empty = bars + 20;  // [4] softmax (8 warps) -> loader warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int kv0 = blockIdx.x * BT;
  const int head = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int nq = (p.seq_q + BT - 1) / BT;   // 128-row query tiles
  const int nsub = 2 * nq;                  // 64-row sub-tiles (the last one may be entirely out of range: harmless)

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(smem_u32(kv_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&q_full[s]), 1); mbar_init(smem_u32(&q_empty[s]), 1);
      mbar_init(smem_u32(&do_full[s]), 1); mbar_init(smem_u32(&do_empty[s]), 1);
      mbar_init(smem_u32(&sd_full[s]), 1);
      mbar_init(smem_u32(&p_ready[s]), 4); mbar_init(smem_u32(&ds_ready[s]), 4);   // one softmax warpgroup per buffer
    }
    mbar_init(smem_u32(acc_done), 1);
    for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&stats_full[s]), 2); mbar_init(smem_u32(&stats_empty[s]), 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t T_ST = tmem_base, T_DPT = tmem_base + 128, T_DV = tmem_base + 256, T_DK = tmem_base + 384;

  if (STATS_WARPS && (warp == 2 || warp == 3)) {
    // warp 2: LSE, warp 3: D = rowsum(O * dO); rows beyond seq_q get LSE = +inf (P = 0) and D = 0
    const float* src = (warp == 2 ? p.lse : p.delta) + (int64_t)bh * p.seq_q;
    const float fill = warp == 2 ? -INFINITY : 0.f;     // the ring holds -LSE (one packed FFMA2 operand per column pair)
    const float sgn = warp == 2 ? -1.f : 1.f;
    for (int j = 0; j < nq; ++j) {
      const int stage = j & 3;
      mbar_wait(smem_u32(&stats_empty[stage]), ((j >> 2) & 1) ^ 1);
      float* dst = lse_smem + stage * 256 + (warp == 2 ? 0 : 128);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int q = j * BT + r * 32 + (int)lane;
        dst[r * 32 + lane] = q < p.seq_q ? sgn * __ldg(src + q) : fill;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&stats_full[stage]));
    }
  }

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t kvb = smem_u32(kv_full);
      mbar_expect_tx(kvb, 2 * BTILE);
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(&tma_k, kvb, smem_u32(k_smem + h * BHALF), h * 64, kv0, bh, kEvictFirst);
        tma_load_3d(&tma_v, kvb, smem_u32(v_smem + h * BHALF), h * 64, kv0, bh, kEvictFirst);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nq; ++j) {
        const int q0 = j * BT;
        mbar_wait(smem_u32(&q_empty[stage]), phase ^ 1);
        const uint32_t qb = smem_u32(&q_full[stage]);
        mbar_expect_tx(qb, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_q, qb, smem_u32(q_smem + stage * BTILE + h * BHALF), h * 64, q0, bh, kEvictLast);
        mbar_wait(smem_u32(&do_empty[stage]), phase ^ 1);
        const uint32_t db = smem_u32(&do_full[stage]);
        mbar_expect_tx(db, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_4d(&tma_do, db, smem_u32(do_smem + stage * BTILE + h * BHALF), h * 64, q0, head, b, kEvictLast);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BT, 64, false, false);   // [128 kv] x [64 q]
      constexpr uint32_t idesc_g = make_idesc_bf16(BT, 128, false, true);   // [128 kv] x [128 d], K = 64 queries
      mbar_wait(smem_u32(kv_full), 0);
      const uint32_t kb = smem_u32(k_smem), vb = smem_u32(v_smem);
      // scores of sub-tile t: 64 query rows starting at row 64*(t&1) of query tile t>>1
      auto issue_scores = [&](int t) {
        const int j = t >> 1, stage = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        if ((t & 1) == 0) { mbar_wait(smem_u32(&q_full[stage]), ph); mbar_wait(smem_u32(&do_full[stage]), ph); }
        tc_fence_after();
        const uint32_t qb = smem_u32(q_smem + stage * BTILE) + (t & 1) * 8192;
        const uint32_t dob = smem_u32(do_smem + stage * BTILE) + (t & 1) * 8192;
        const uint32_t buf = t & 1;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_ST + buf * 64, make_smem_desc(kb + kmajor_off(k), 16, 1024), make_smem_desc(qb + kmajor_off(k), 16, 1024),
                     idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_DPT + buf * 64, make_smem_desc(vb + kmajor_off(k), 16, 1024), make_smem_desc(dob + kmajor_off(k), 16, 1024),
                     idesc_s, k != 0);
        umma_commit<1>(smem_u32(&sd_full[buf]));
      };
      issue_scores(0);
      for (int t = 0; t < nsub; ++t) {
        if (t + 1 < nsub) issue_scores(t + 1);
        const int j = t >> 1, stage = j & 1;
        const uint32_t buf = t & 1, rph = (t >> 1) & 1;
        const uint32_t qb = smem_u32(q_smem + stage * BTILE) + (t & 1) * 8192;
        const uint32_t dob = smem_u32(do_smem + stage * BTILE) + (t & 1) * 8192;
        mbar_wait(smem_u32(&p_ready[buf]), rph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dV[kv, d] += Pt[kv, 64 q] dO[64 q, d]
          umma_ts(T_DV, T_ST + buf * 64 + k * 8, make_smem_desc(dob + k * 2048, BHALF, 1024), idesc_g, (t | k) != 0);
        if (t & 1) umma_commit<1>(smem_u32(&do_empty[stage]));
        mbar_wait(smem_u32(&ds_ready[buf]), rph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dK[kv, d] += dSt[kv, 64 q] Q[64 q, d]
          umma_ts(T_DK, T_DPT + buf * 64 + k * 8, make_smem_desc(qb + k * 2048, BHALF, 1024), idesc_g, (t | k) != 0);
        if (t & 1) umma_commit<1>(smem_u32(&q_empty[stage]));
      }
      umma_commit<1>(smem_u32(acc_done));
    }
  } else if (warp >= 4) {
    // Warpgroup wg (warps 4-7 / 8-11) owns TMEM buffer wg and therefore every sub-tile t with t % 2 == wg: the two
    // warpgroups work on different sub-tiles at any time, so their TMEM waits and MUFU bursts interleave.
    const uint32_t quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int wg_tid = ((warp - 4) & 3) * 32 + lane;   // 0..127 inside the warpgroup
    const uint32_t lane_base = (quad * 32u) << 16;
    const float c = p.scale_log2;
    float* lse_s = lse_smem + wg * 128;                // [64] LSE then [64] D of the current sub-tile
    float* dl_s = lse_s + 64;
    const uint32_t buf = wg;
    for (int t = wg, it = 0; t < nsub; t += 2, ++it) {
      const uint32_t rph = it & 1;
      if (STATS_WARPS) {   // the loader warps are several query tiles ahead: normally no wait at all
        const int j = t >> 1, stage = j & 3;
        mbar_wait(smem_u32(&stats_full[stage]), (j >> 2) & 1);
        lse_s = lse_smem + stage * 256 + (t & 1) * 64;
        dl_s = lse_s + 128;
      } else {   // stage LSE / D of the 64 query rows of this sub-tile (rows beyond seq_q: +inf -> P = 0)
        const int q = t * 64 + (wg_tid & 63);
        named_bar_sync(1 + wg, 128);                   // previous sub-tile's readers are done
        if (wg_tid < 64) lse_s[wg_tid] = (q < p.seq_q) ? -p.lse[(int64_t)bh * p.seq_q + q] : -INFINITY;
        else dl_s[wg_tid - 64] = (q < p.seq_q) ? p.delta[(int64_t)bh * p.seq_q + q] : 0.f;
        named_bar_sync(1 + wg, 128);
      }
      mbar_wait(smem_u32(&sd_full[buf]), rph);
      tc_fence_after();
      f32x2 pv[32];                                   // P of this thread's key row, two query columns per entry
      const f32x2 c2 = f2_pack(c, c);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], pk[16];
        tmem_ld_x32(T_ST + lane_base + buf * 64 + hh * 32, r);
        tmem_ld_wait();
        const f32x2* nl = reinterpret_cast<const f32x2*>(lse_s + hh * 32);   // -LSE of two adjacent query columns
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const f32x2 e = f2_exp2_mufu(f2_fma(f2_pack_bits(r[x], r[x + 1]), c2, nl[x >> 1]));
          pv[hh * 16 + (x >> 1)] = e;
          float p0, p1;
          f2_unpack(e, p0, p1);
          pk[x >> 1] = pack_bf16(p0, p1);
        }
        tmem_st_x16(T_ST + lane_base + buf * 64 + hh * 16, pk);   // packed P over columns this thread has already read
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_ready[buf]));
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], pk[16];
        tmem_ld_x32(T_DPT + lane_base + buf * 64 + hh * 32, r);
        tmem_ld_wait();
        const f32x2* dl2 = reinterpret_cast<const f32x2*>(dl_s + hh * 32);
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const f32x2 d = f2_mul(pv[hh * 16 + (x >> 1)], f2_sub(f2_pack_bits(r[x], r[x + 1]), dl2[x >> 1]));
          float d0, d1;
          f2_unpack(d, d0, d1);
          pk[x >> 1] = pack_bf16(d0, d1);
        }
        tmem_st_x16(T_DPT + lane_base + buf * 64 + hh * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&ds_ready[buf]));
        if (STATS_WARPS) mbar_arrive(smem_u32(&stats_empty[(t >> 1) & 3]));
      }
    }
    mbar_wait(smem_u32(acc_done), 0);
    tc_fence_after();
    const int kv = kv0 + quad * 32 + lane;
    const bool ok = kv < p.seq_k;
    const int chalf = (warp - 4) >> 2;
    const int64_t o = ((int64_t)bh * p.seq_k + kv) * 128 + chalf * 64;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tt = (which == 0 ? T_DV : T_DK) + lane_base + chalf * 64;
      __nv_bfloat16* dst = (which == 0 ? p.dv : p.dk) + o;
      const float mul = which == 0 ? 1.0f : p.scale;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t rr[32];
        tmem_ld_x32(tt + cc * 32, rr);
        tmem_ld_wait();
        if (ok) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            uint4 qv;
            qv.x = pack_bf16(__uint_as_float(rr[x * 8 + 0]) * mul, __uint_as_float(rr[x * 8 + 1]) * mul);
            qv.y = pack_bf16(__uint_as_float(rr[x * 8 + 2]) * mul, __uint_as_float(rr[x * 8 + 3]) * mul);
            qv.z = pack_bf16(__uint_as_float(rr[x * 8 + 4]) * mul, __uint_as_float(rr[x * 8 + 5]) * mul);
            qv.w = pack_bf16(__uint_as_float(rr[x * 8 + 6]) * mul, __uint_as_float(rr[x * 8 + 7]) * mul);
            d4[x] = qv;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq2_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                    const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                    const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_smem = smem;
  uint8_t* do_smem = smem + BTILE;
  uint8_t* k_smem = smem + 2 * BTILE;  // 2 stages of [128 kv][128 d]
  uint8_t* v_smem = smem + 4 * BTILE;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * BTILE + AB_STATS_BYTES);
  uint64_t* qdo_full = bars;       // [1]
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* sd_full = bars + 9;    // [2]
  uint64_t* ds_ready = bars + 11;  // [2] (8 warps)
  uint64_t* acc_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * BT;
  const int head = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int nkv = (p.seq_k + BT - 1) / BT;
  const int nsub = 2 * nkv;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(smem_u32(qdo_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1); mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1); mbar_init(smem_u32(&v_empty[s]), 1);
      mbar_init(smem_u32(&sd_full[s]), 1);
      mbar_init(smem_u32(&ds_ready[s]), 4);
    }
    mbar_init(smem_u32(acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 128, T_DQ = tmem_base + 256;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t qb = smem_u32(qdo_full);
      mbar_expect_tx(qb, 2 * BTILE);
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(&tma_q, qb, smem_u32(q_smem + h * BHALF), h * 64, q0, bh, kEvictFirst);
        tma_load_4d(&tma_do, qb, smem_u32(do_smem + h * BHALF), h * 64, q0, head, b, kEvictFirst);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        const int kv0 = j * BT;
        mbar_wait(smem_u32(&k_empty[stage]), phase ^ 1);
        const uint32_t kb = smem_u32(&k_full[stage]);
        mbar_expect_tx(kb, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_k, kb, smem_u32(k_smem + stage * BTILE + h * BHALF), h * 64, kv0, bh, kEvictLast);
        mbar_wait(smem_u32(&v_empty[stage]), phase ^ 1);
        const uint32_t vb = smem_u32(&v_full[stage]);
        mbar_expect_tx(vb, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_v, vb, smem_u32(v_smem + stage * BTILE + h * BHALF), h * 64, kv0, bh, kEvictLast);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BT, 64, false, false);   // [128 q] x [64 kv]
      constexpr uint32_t idesc_g = make_idesc_bf16(BT, 128, false, true);   // [128 q] x [128 d], K = 64 keys
      const uint32_t qb = smem_u32(q_smem), dob = smem_u32(do_smem);
      mbar_wait(smem_u32(qdo_full), 0);
      auto issue_scores = [&](int t) {
        const int j = t >> 1, stage = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        if ((t & 1) == 0) { mbar_wait(smem_u32(&k_full[stage]), ph); mbar_wait(smem_u32(&v_full[stage]), ph); }
        tc_fence_after();
        const uint32_t kb = smem_u32(k_smem + stage * BTILE) + (t & 1) * 8192;
        const uint32_t vb = smem_u32(v_smem + stage * BTILE) + (t & 1) * 8192;
        const uint32_t buf = t & 1;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_S + buf * 64, make_smem_desc(qb + kmajor_off(k), 16, 1024), make_smem_desc(kb + kmajor_off(k), 16, 1024),
                     idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_DP + buf * 64, make_smem_desc(dob + kmajor_off(k), 16, 1024), make_smem_desc(vb + kmajor_off(k), 16, 1024),
                     idesc_s, k != 0);
        umma_commit<1>(smem_u32(&sd_full[buf]));
        if (t & 1) umma_commit<1>(smem_u32(&v_empty[stage]));   // V tile fully consumed by the two dP sub-tiles
      };
      issue_scores(0);
      for (int t = 0; t < nsub; ++t) {
        if (t + 1 < nsub) issue_scores(t + 1);
        const int j = t >> 1, stage = j & 1;
        const uint32_t buf = t & 1, rph = (t >> 1) & 1;
        const uint32_t kb = smem_u32(k_smem + stage * BTILE) + (t & 1) * 8192;
        mbar_wait(smem_u32(&ds_ready[buf]), rph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dQ[q, d] += dS[q, 64 kv] K[64 kv, d]
          umma_ts(T_DQ, T_DP + buf * 64 + k * 8, make_smem_desc(kb + k * 2048, BHALF, 1024), idesc_g, (t | k) != 0);
        if (t & 1) umma_commit<1>(smem_u32(&k_empty[stage]));
      }
      umma_commit<1>(smem_u32(acc_done));
    }
  } else if (warp >= 4) {
    const uint32_t quad = warp & 3;
    const int wg = (warp - 4) >> 2;          // warpgroup wg owns TMEM buffer wg, i.e. key sub-tiles t with t % 2 == wg
    const int half = wg;                     // (epilogue: which 64 output columns this warp stores)
    const uint32_t lane_base = (quad * 32u) << 16;
    const int q = q0 + quad * 32 + lane;
    const bool ok = q < p.seq_q;
    const float lse = ok ? p.lse[(int64_t)bh * p.seq_q + q] : INFINITY;
    const float dl = ok ? p.delta[(int64_t)bh * p.seq_q + q] : 0.f;
    const float c = p.scale_log2;
    const uint32_t buf = wg;
    for (int t = wg, it = 0; t < nsub; t += 2, ++it) {
      const uint32_t rph = it & 1;
      const int valid = p.seq_k - t * 64;    // keys of this sub-tile that exist (may be <= 0 or >= 64)
      mbar_wait(smem_u32(&sd_full[buf]), rph);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], rd[32], pk[16];
        tmem_ld_x32(T_S + lane_base + buf * 64 + hh * 32, r);
        tmem_ld_x32(T_DP + lane_base + buf * 64 + hh * 32, rd);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[x]), c, -lse));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[x + 1]), c, -lse));
          if (hh * 32 + x >= valid) p0 = 0.f;
          if (hh * 32 + x + 1 >= valid) p1 = 0.f;
          pk[x >> 1] = pack_bf16(p0 * (__uint_as_float(rd[x]) - dl), p1 * (__uint_as_float(rd[x + 1]) - dl));
        }
        tmem_st_x16(T_DP + lane_base + buf * 64 + hh * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ds_ready[buf]));
    }
    mbar_wait(smem_u32(acc_done), 0);
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + ((int64_t)bh * p.seq_q + q) * 128 + half * 64;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t rr[32];
      tmem_ld_x32(T_DQ + lane_base + half * 64 + cc * 32, rr);
      tmem_ld_wait();
      if (ok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 qv;
          qv.x = pack_bf16(__uint_as_float(rr[x * 8 + 0]) * p.scale, __uint_as_float(rr[x * 8 + 1]) * p.scale);
          qv.y = pack_bf16(__uint_as_float(rr[x * 8 + 2]) * p.scale, __uint_as_float(rr[x * 8 + 3]) * p.scale);
          qv.z = pack_bf16(__uint_as_float(rr[x * 8 + 4]) * p.scale, __uint_as_float(rr[x * 8 + 5]) * p.scale);
          qv.w = pack_bf16(__uint_as_float(rr[x * 8 + 6]) * p.scale, __uint_as_float(rr[x * 8 + 7]) * p.scale);
          d4[x] = qv;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}


// dS = P * (dP - D) for 32 score columns of one query row: P = exp2(S * c - LSE); packed fp32 pairs throughout
// (FFMA2 / FADD2 / FMUL2: the FMA pipe, not MUFU and not the tensor core, bounded the round-1 kernels)
template <bool MASK>
__device__ __forceinline__ void dq_softmax_chunk(const uint32_t (&r)[32], const uint32_t (&rd)[32], uint32_t (&pk)[16],
                                                 f32x2 c2, f32x2 nlse2, f32x2 dl2, int valid) {
#pragma unroll
  for (int x = 0; x < 32; x += 2) {
    f32x2 e = f2_exp2_mufu(f2_fma(f2_pack_bits(r[x], r[x + 1]), c2, nlse2));
    if (MASK && x + 1 >= valid) {       // keys beyond seq_k (last, partial sub-tile only)
      float p0, p1;
      f2_unpack(e, p0, p1);
      e = f2_pack(x >= valid ? 0.f : p0, 0.f);
    }
    const f32x2 d = f2_mul(e, f2_sub(f2_pack_bits(rd[x], rd[x + 1]), dl2));
    float d0, d1;
    f2_unpack(d, d0, d1);
    pk[x >> 1] = pack_bf16(d0, d1);
  }
}

// ---------------------------------------------------------------------------------------------
// v3 dQ: three S / dP accumulator buffers in TMEM instead of two, and K / V streamed as 64-row half tiles.
//   TMEM: S[3] 3x64 | dP[3] 3x64 | dQ 128  = 512 columns.
// With two buffers the scores of sub-tile t+2 can only be issued once dS(t) has been consumed, so each softmax warpgroup
// sits in the chain  dS(t) -> dQ(t) MMA -> scores(t+2) MMA -> softmax(t+2)  and the tensor pipe idles for the softmax
// latency of every sub-tile (r01b: 52 % tensor-pipe active).  With three, scores(t+2) go out BEFORE the wait for dS(t):
// when a warpgroup finishes sub-tile t its next one (t+2) is already in TMEM, and the pipe is fed as long as one softmax
// pass (2 warpgroups alternating) is shorter than two sub-tiles of MMA work.  K / V travel as four 64-row slots each
// (slot = t & 3, released per sub-tile), so a slot has two sub-tile periods to land before its scores are issued and the
// MMA thread does not block on a full-tile load behind a pending dQ.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq3_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k64,
                    const __grid_constant__ CUtensorMap tma_v64, const __grid_constant__ CUtensorMap tma_do,
                    const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_smem = smem;
  uint8_t* do_smem = smem + BTILE;
  uint8_t* k_smem = smem + 2 * BTILE;  // 2 tiles of [128 kv][128 d] = 4 half-tile slots
  uint8_t* v_smem = smem + 4 * BTILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * BTILE + AB_STATS_BYTES);
  uint64_t* qdo_full = bars;        // [1]
  uint64_t* k_full = bars + 1;      // [4]
  uint64_t* k_empty = bars + 5;     // [4]
  uint64_t* v_full = bars + 9;      // [4]
  uint64_t* v_empty = bars + 13;    // [4]
  uint64_t* sd_full = bars + 17;    // [3]
  uint64_t* ds_ready = bars + 20;   // [3] (4 warps)
  uint64_t* acc_done = bars + 23;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  constexpr int NB = 3;

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * BT;
  const int head = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int nsub = (p.seq_k + 63) / 64;   // 64-key sub-tiles

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k64); tma_prefetch_desc(&tma_v64); tma_prefetch_desc(&tma_do);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(smem_u32(qdo_full), 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1); mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1); mbar_init(smem_u32(&v_empty[s]), 1);
    }
    for (int s = 0; s < NB; ++s) { mbar_init(smem_u32(&sd_full[s]), 1); mbar_init(smem_u32(&ds_ready[s]), 4); }
    mbar_init(smem_u32(acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t T_S = tmem_base, T_DP = tmem_base + NB * 64, T_DQ = tmem_base + 2 * NB * 64;
  // byte offset of half-tile slot `s` inside a 2-tile operand buffer: tile (s >> 1), rows 64 * (s & 1)
  auto slot_off = [](int s) { return (uint32_t)((s >> 1) * BTILE + (s & 1) * 8192); };

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t qb = smem_u32(qdo_full);
      mbar_expect_tx(qb, 2 * BTILE);
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(&tma_q, qb, smem_u32(q_smem + h * BHALF), h * 64, q0, bh, kEvictFirst);
        tma_load_4d(&tma_do, qb, smem_u32(do_smem + h * BHALF), h * 64, q0, head, b, kEvictFirst);
      }
      for (int t = 0; t < nsub; ++t) {
        const int slot = t & 3;
        const uint32_t ph = ((t >> 2) & 1) ^ 1;
        mbar_wait(smem_u32(&k_empty[slot]), ph);
        const uint32_t kb = smem_u32(&k_full[slot]);
        mbar_expect_tx(kb, BTILE / 2);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_k64, kb, smem_u32(k_smem + slot_off(slot) + h * BHALF), h * 64, t * 64, bh, kEvictLast);
        mbar_wait(smem_u32(&v_empty[slot]), ph);
        const uint32_t vb = smem_u32(&v_full[slot]);
        mbar_expect_tx(vb, BTILE / 2);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_v64, vb, smem_u32(v_smem + slot_off(slot) + h * BHALF), h * 64, t * 64, bh, kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BT, 64, false, false);   // [128 q] x [64 kv]
      constexpr uint32_t idesc_g = make_idesc_bf16(BT, 128, false, true);   // [128 q] x [128 d], K = 64 keys
      const uint32_t qb = smem_u32(q_smem), dob = smem_u32(do_smem);
      mbar_wait(smem_u32(qdo_full), 0);
      auto issue_scores = [&](int t) {
        const int slot = t & 3;
        const uint32_t ph = (t >> 2) & 1;
        mbar_wait(smem_u32(&k_full[slot]), ph);
        mbar_wait(smem_u32(&v_full[slot]), ph);
        tc_fence_after();
        const uint32_t kb = smem_u32(k_smem + slot_off(slot));
        const uint32_t vb = smem_u32(v_smem + slot_off(slot));
        const uint32_t buf = t % NB;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_S + buf * 64, make_smem_desc(qb + kmajor_off(k), 16, 1024), make_smem_desc(kb + kmajor_off(k), 16, 1024),
                     idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_DP + buf * 64, make_smem_desc(dob + kmajor_off(k), 16, 1024), make_smem_desc(vb + kmajor_off(k), 16, 1024),
                     idesc_s, k != 0);
        umma_commit<1>(smem_u32(&sd_full[buf]));
        umma_commit<1>(smem_u32(&v_empty[slot]));      // this V half tile is consumed by its dP
      };
      for (int t = 0; t < NB - 1 && t < nsub; ++t) issue_scores(t);
      for (int t = 0; t < nsub; ++t) {
        if (t + NB - 1 < nsub) issue_scores(t + NB - 1);
        const int slot = t & 3;
        const uint32_t buf = t % NB, rph = (t / NB) & 1;
        const uint32_t kb = smem_u32(k_smem + slot_off(slot));
        mbar_wait(smem_u32(&ds_ready[buf]), rph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dQ[q, d] += dS[q, 64 kv] K[64 kv, d]
          umma_ts(T_DQ, T_DP + buf * 64 + k * 8, make_smem_desc(kb + k * 2048, BHALF, 1024), idesc_g, (t | k) != 0);
        umma_commit<1>(smem_u32(&k_empty[slot]));
      }
      umma_commit<1>(smem_u32(acc_done));
    }
  } else if (warp >= 4) {
    const uint32_t quad = warp & 3;
    const int wg = (warp - 4) >> 2;          // warpgroup wg takes the key sub-tiles t with t % 2 == wg (buffer t % 3)
    const int half = wg;                     // (epilogue: which 64 output columns this warp stores)
    const uint32_t lane_base = (quad * 32u) << 16;
    const int q = q0 + quad * 32 + lane;
    const bool ok = q < p.seq_q;
    const float lse = ok ? p.lse[(int64_t)bh * p.seq_q + q] : INFINITY;
    const float dl = ok ? p.delta[(int64_t)bh * p.seq_q + q] : 0.f;
    const float c = p.scale_log2;
    const f32x2 c2 = f2_pack(c, c), nlse2 = f2_pack(-lse, -lse), dl2 = f2_pack(dl, dl);
    for (int t = wg; t < nsub; t += 2) {
      const uint32_t buf = t % NB, rph = (t / NB) & 1;
      const int valid = p.seq_k - t * 64;    // keys of this sub-tile that exist (>= 64 except in the last one)
      const bool partial = valid < 64;
      mbar_wait(smem_u32(&sd_full[buf]), rph);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], rd[32], pk[16];
        tmem_ld_x32(T_S + lane_base + buf * 64 + hh * 32, r);
        tmem_ld_x32(T_DP + lane_base + buf * 64 + hh * 32, rd);
        tmem_ld_wait();
        if (!partial) dq_softmax_chunk<false>(r, rd, pk, c2, nlse2, dl2, 0);
        else dq_softmax_chunk<true>(r, rd, pk, c2, nlse2, dl2, valid - hh * 32);
        tmem_st_x16(T_DP + lane_base + buf * 64 + hh * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ds_ready[buf]));
    }
    mbar_wait(smem_u32(acc_done), 0);
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + ((int64_t)bh * p.seq_q + q) * 128 + half * 64;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t rr[32];
      tmem_ld_x32(T_DQ + lane_base + half * 64 + cc * 32, rr);
      tmem_ld_wait();
      if (ok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 qv;
          qv.x = pack_bf16(__uint_as_float(rr[x * 8 + 0]) * p.scale, __uint_as_float(rr[x * 8 + 1]) * p.scale);
          qv.y = pack_bf16(__uint_as_float(rr[x * 8 + 2]) * p.scale, __uint_as_float(rr[x * 8 + 3]) * p.scale);
          qv.z = pack_bf16(__uint_as_float(rr[x * 8 + 4]) * p.scale, __uint_as_float(rr[x * 8 + 5]) * p.scale);
          qv.w = pack_bf16(__uint_as_float(rr[x * 8 + 6]) * p.scale, __uint_as_float(rr[x * 8 + 7]) * p.scale);
          d4[x] = qv;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}


// ---------------------------------------------------------------------------------------------
// v4 dQ: Q and dO live in TENSOR MEMORY as the A operands of the score MMAs (TS form), not in shared memory.
//   TMEM: S[2] 2x64 | dP[2] 2x64 | dQ 128 | Q 64 | dO 64   = 512 columns (a [128 x 128] bf16 tile is 64 32-bit columns)
// The r02 capture of the v3 kernel showed what bounds these kernels: not MUFU, not the FMA pipe, but the shared-memory
// data pipe feeding the tensor core (l1tex__data_pipe_tc_wavefronts_mem_shared at 60 % of peak with the MMA queue full and the
// tensor pipe 54 % active).  An SS-form MMA of N = 64 reads 4 KB of A and 2 KB of B from shared memory for 32 cycles of
// math; the A operand of both score MMAs (the Q and dO tile of the CTA) never changes, so reading it from TMEM removes
// two thirds of the operand traffic: 112 KB -> 48 KB per 64-key sub-tile.  Each query row is written to TMEM once by the
// thread that owns it (global -> registers -> tcgen05.st, no shared-memory staging), and the 64 KB of shared memory this
// frees hold a third K / V tile (6 half-tile slots each).
// ---------------------------------------------------------------------------------------------
constexpr int DQ4_SLOTS = 6;
constexpr int DQ4_SMEM_BYTES = 2 * (DQ4_SLOTS / 2) * BTILE + 1024 + 256;

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dq4_kernel(const __grid_constant__ CUtensorMap tma_k64, const __grid_constant__ CUtensorMap tma_v64,
                    const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* k_smem = smem;                                  // 3 tiles of [128 kv][128 d] = 6 half-tile slots
  uint8_t* v_smem = smem + (DQ4_SLOTS / 2) * BTILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (DQ4_SLOTS / 2) * BTILE);
  uint64_t* k_full = bars;                    // [6]
  uint64_t* k_empty = bars + DQ4_SLOTS;       // [6]
  uint64_t* v_full = bars + 2 * DQ4_SLOTS;    // [6]
  uint64_t* v_empty = bars + 3 * DQ4_SLOTS;   // [6]
  uint64_t* sd_full = bars + 4 * DQ4_SLOTS;   // [2]
  uint64_t* ds_ready = sd_full + 2;           // [2] (4 warps)
  uint64_t* qa_ready = ds_ready + 2;          // [1] (8 warps): Q and dO rows are in TMEM
  uint64_t* acc_done = qa_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int q0 = blockIdx.x * BT;
  const int head = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int nsub = (p.seq_k + 63) / 64;   // 64-key sub-tiles

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tma_k64); tma_prefetch_desc(&tma_v64); }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < DQ4_SLOTS; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1); mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1); mbar_init(smem_u32(&v_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&sd_full[s]), 1); mbar_init(smem_u32(&ds_ready[s]), 4); }
    mbar_init(smem_u32(qa_ready), 8);
    mbar_init(smem_u32(acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t T_S = tmem_base, T_DP = tmem_base + 128, T_DQ = tmem_base + 256, T_QA = tmem_base + 384, T_DOA = tmem_base + 448;
  auto slot_off = [](int s) { return (uint32_t)((s >> 1) * BTILE + (s & 1) * 8192); };

  if (warp == 0) {
    if (elect_one()) {
      int slot = 0;
      uint32_t ph = 1;
      for (int t = 0; t < nsub; ++t) {
        mbar_wait(smem_u32(&k_empty[slot]), ph);
        const uint32_t kb = smem_u32(&k_full[slot]);
        mbar_expect_tx(kb, BTILE / 2);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_k64, kb, smem_u32(k_smem + slot_off(slot) + h * BHALF), h * 64, t * 64, bh, kEvictLast);
        mbar_wait(smem_u32(&v_empty[slot]), ph);
        const uint32_t vb = smem_u32(&v_full[slot]);
        mbar_expect_tx(vb, BTILE / 2);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_v64, vb, smem_u32(v_smem + slot_off(slot) + h * BHALF), h * 64, t * 64, bh, kEvictLast);
        if (++slot == DQ4_SLOTS) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BT, 64, false, false);   // [128 q] x [64 kv], A from TMEM
      constexpr uint32_t idesc_g = make_idesc_bf16(BT, 128, false, true);   // [128 q] x [128 d], K = 64 keys
      mbar_wait(smem_u32(qa_ready), 0);
      tc_fence_after();
      int is_slot = 0;            // slot / phase of the next issue_scores
      uint32_t is_ph = 0;
      auto issue_scores = [&](int t) {
        mbar_wait(smem_u32(&k_full[is_slot]), is_ph);
        mbar_wait(smem_u32(&v_full[is_slot]), is_ph);
        tc_fence_after();
        const uint32_t kb = smem_u32(k_smem + slot_off(is_slot));
        const uint32_t vb = smem_u32(v_smem + slot_off(is_slot));
        const uint32_t buf = t & 1;
#pragma unroll
        for (int k = 0; k < 8; ++k)   // S[q, 64 kv] = Q K^T
          umma_ts(T_S + buf * 64, T_QA + k * 8, make_smem_desc(kb + kmajor_off(k), 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dP[q, 64 kv] = dO V^T
          umma_ts(T_DP + buf * 64, T_DOA + k * 8, make_smem_desc(vb + kmajor_off(k), 16, 1024), idesc_s, k != 0);
        umma_commit<1>(smem_u32(&sd_full[buf]));
        umma_commit<1>(smem_u32(&v_empty[is_slot]));      // this V half tile is consumed by its dP
        if (++is_slot == DQ4_SLOTS) { is_slot = 0; is_ph ^= 1; }
      };
      issue_scores(0);
      int slot = 0;
      for (int t = 0; t < nsub; ++t) {
        if (t + 1 < nsub) issue_scores(t + 1);
        const uint32_t buf = t & 1, rph = (t >> 1) & 1;
        const uint32_t kb = smem_u32(k_smem + slot_off(slot));
        mbar_wait(smem_u32(&ds_ready[buf]), rph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dQ[q, d] += dS[q, 64 kv] K[64 kv, d]
          umma_ts(T_DQ, T_DP + buf * 64 + k * 8, make_smem_desc(kb + k * 2048, BHALF, 1024), idesc_g, (t | k) != 0);
        umma_commit<1>(smem_u32(&k_empty[slot]));
        if (++slot == DQ4_SLOTS) slot = 0;
      }
      umma_commit<1>(smem_u32(acc_done));
    }
  } else if (warp >= 4) {
    const uint32_t quad = warp & 3;
    const int wg = (warp - 4) >> 2;          // warpgroup wg owns TMEM score buffer wg = the key sub-tiles t with t % 2 == wg
    const int half = wg;                     // (epilogue: which 64 output columns this warp stores)
    const uint32_t lane_base = (quad * 32u) << 16;
    const int q = q0 + quad * 32 + lane;
    const bool ok = q < p.seq_q;
    {   // this thread's row of Q (warpgroup 0) or dO (warpgroup 1): 256 contiguous bytes -> 64 TMEM columns of its lane
      const __nv_bfloat16* row = wg == 0 ? p.q + ((int64_t)bh * p.seq_q + q) * 128
                                         : p.d_o + ((int64_t)b * p.seq_q + q) * p.lddo + head * 128;
      const uint32_t dst = (wg == 0 ? T_QA : T_DOA) + lane_base;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
          if (ok) v4 = __ldg(reinterpret_cast<const uint4*>(row) + cc * 8 + x);
          r[x * 4 + 0] = v4.x; r[x * 4 + 1] = v4.y; r[x * 4 + 2] = v4.z; r[x * 4 + 3] = v4.w;
        }
        tmem_st_x32(dst + cc * 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(qa_ready));
    }
    const float lse = ok ? p.lse[(int64_t)bh * p.seq_q + q] : INFINITY;
    const float dl = ok ? p.delta[(int64_t)bh * p.seq_q + q] : 0.f;
    const float c = p.scale_log2;
    const f32x2 c2 = f2_pack(c, c), nlse2 = f2_pack(-lse, -lse), dl2 = f2_pack(dl, dl);
    const uint32_t buf = wg;
    for (int t = wg, it = 0; t < nsub; t += 2, ++it) {
      const uint32_t rph = it & 1;
      const int valid = p.seq_k - t * 64;    // keys of this sub-tile that exist (>= 64 except in the last one)
      const bool partial = valid < 64;
      mbar_wait(smem_u32(&sd_full[buf]), rph);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], rd[32], pk[16];
        tmem_ld_x32(T_S + lane_base + buf * 64 + hh * 32, r);
        tmem_ld_x32(T_DP + lane_base + buf * 64 + hh * 32, rd);
        tmem_ld_wait();
        if (!partial) dq_softmax_chunk<false>(r, rd, pk, c2, nlse2, dl2, 0);
        else dq_softmax_chunk<true>(r, rd, pk, c2, nlse2, dl2, valid - hh * 32);
        tmem_st_x16(T_DP + lane_base + buf * 64 + hh * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ds_ready[buf]));
    }
    mbar_wait(smem_u32(acc_done), 0);
    tc_fence_after();
    __nv_bfloat16* dst = p.dq + ((int64_t)bh * p.seq_q + q) * 128 + half * 64;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t rr[32];
      tmem_ld_x32(T_DQ + lane_base + half * 64 + cc * 32, rr);
      tmem_ld_wait();
      if (ok) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 qv;
          qv.x = pack_bf16(__uint_as_float(rr[x * 8 + 0]) * p.scale, __uint_as_float(rr[x * 8 + 1]) * p.scale);
          qv.y = pack_bf16(__uint_as_float(rr[x * 8 + 2]) * p.scale, __uint_as_float(rr[x * 8 + 3]) * p.scale);
          qv.z = pack_bf16(__uint_as_float(rr[x * 8 + 4]) * p.scale, __uint_as_float(rr[x * 8 + 5]) * p.scale);
          qv.w = pack_bf16(__uint_as_float(rr[x * 8 + 6]) * p.scale, __uint_as_float(rr[x * 8 + 7]) * p.scale);
          d4[x] = qv;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}


// ---------------------------------------------------------------------------------------------
// v3 dK / dV: the K tile of the CTA is the A operand of St = K Q^T from TENSOR MEMORY (TS form) — same reasoning as the v4 dQ
// kernel: the shared-memory data pipe (tensor-core operand reads 53 % + LSE / D broadcasts 32 % of its peak in the r02
// capture of v2) bounds this kernel, and the stationary A operand is two thirds of the score MMAs' operand bytes.
//   TMEM: St[2] 2x64 | dPt 64 (single buffer) | K 64 | dV 128 | dK 128 = 512 columns
// There is no room for V as well, so dPt = V dO^T stays an SS MMA and gets ONE buffer: dPt(t+1) is issued right after the
// dK MMA that consumes dSt(t) (in-order tensor pipe: no extra barrier), while St keeps two buffers so that the exponentials of
// sub-tile t+1 run under the dV / dK MMAs of sub-tile t.  Warpgroup t & 1 owns sub-tile t.
// ---------------------------------------------------------------------------------------------
constexpr int DKV3_SMEM_BYTES = 5 * BTILE + AB_STATS_BYTES + 1024 + 256;

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_dkv3_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_v,
                     const __grid_constant__ CUtensorMap tma_do, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* v_smem = smem;
  uint8_t* q_smem = smem + BTILE;       // 2 stages of [128 q][128 d]
  uint8_t* do_smem = smem + 3 * BTILE;  // 2 stages
  float* stats = reinterpret_cast<float*>(smem + 5 * BTILE);   // ring [4][-LSE 128 | D 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * BTILE + AB_STATS_BYTES);
  uint64_t* v_full = bars;          // [1]
  uint64_t* q_full = bars + 1;      // [2]
  uint64_t* q_empty = bars + 3;     // [2]
  uint64_t* do_full = bars + 5;     // [2]
  uint64_t* do_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;      // [2] MMA -> softmax: St of buffer b complete
  uint64_t* dp_full = bars + 11;    // [1] MMA -> softmax: dPt complete
  uint64_t* p_ready = bars + 12;    // [2] softmax -> MMA (4 warps)
  uint64_t* ds_ready = bars + 14;   // [2] softmax -> MMA (4 warps)
  uint64_t* ka_ready = bars + 16;   // [1] (8 warps): the K tile is in TMEM
  uint64_t* acc_done = bars + 17;
  uint64_t* stats_full = bars + 18;   // [4]
  uint64_t* stats_empty = bars + 22;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int kv0 = blockIdx.x * BT;
  const int head = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int nq = (p.seq_q + BT - 1) / BT;
  const int nsub = 2 * nq;

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do); }
  if (warp == 1 && elect_one()) {
    mbar_init(smem_u32(v_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&q_full[s]), 1); mbar_init(smem_u32(&q_empty[s]), 1);
      mbar_init(smem_u32(&do_full[s]), 1); mbar_init(smem_u32(&do_empty[s]), 1);
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&p_ready[s]), 4); mbar_init(smem_u32(&ds_ready[s]), 4);
    }
    mbar_init(smem_u32(dp_full), 1);
    mbar_init(smem_u32(ka_ready), 8);
    mbar_init(smem_u32(acc_done), 1);
    for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&stats_full[s]), 2); mbar_init(smem_u32(&stats_empty[s]), 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t T_ST = tmem_base, T_DPT = tmem_base + 128, T_KA = tmem_base + 192, T_DV = tmem_base + 256, T_DK = tmem_base + 384;

  if (warp == 2 || warp == 3) {
    // warp 2: -LSE, warp 3: D = rowsum(O * dO); rows beyond seq_q get -LSE = -inf (P = 0) and D = 0
    const float* src = (warp == 2 ? p.lse : p.delta) + (int64_t)bh * p.seq_q;
    const float fill = warp == 2 ? -INFINITY : 0.f;
    const float sgn = warp == 2 ? -1.f : 1.f;
    for (int j = 0; j < nq; ++j) {
      const int stage = j & 3;
      mbar_wait(smem_u32(&stats_empty[stage]), ((j >> 2) & 1) ^ 1);
      float* dst = stats + stage * 256 + (warp == 2 ? 0 : 128);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int q = j * BT + r * 32 + (int)lane;
        dst[r * 32 + lane] = q < p.seq_q ? sgn * __ldg(src + q) : fill;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&stats_full[stage]));
    }
  }

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t vb = smem_u32(v_full);
      mbar_expect_tx(vb, BTILE);
      for (int h = 0; h < 2; ++h) tma_load_3d(&tma_v, vb, smem_u32(v_smem + h * BHALF), h * 64, kv0, bh, kEvictFirst);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nq; ++j) {
        const int q0 = j * BT;
        mbar_wait(smem_u32(&q_empty[stage]), phase ^ 1);
        const uint32_t qb = smem_u32(&q_full[stage]);
        mbar_expect_tx(qb, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_q, qb, smem_u32(q_smem + stage * BTILE + h * BHALF), h * 64, q0, bh, kEvictLast);
        mbar_wait(smem_u32(&do_empty[stage]), phase ^ 1);
        const uint32_t db = smem_u32(&do_full[stage]);
        mbar_expect_tx(db, BTILE);
        for (int h = 0; h < 2; ++h)
          tma_load_4d(&tma_do, db, smem_u32(do_smem + stage * BTILE + h * BHALF), h * 64, q0, head, b, kEvictLast);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BT, 64, false, false);   // [128 kv] x [64 q]
      constexpr uint32_t idesc_g = make_idesc_bf16(BT, 128, false, true);   // [128 kv] x [128 d], K = 64 queries
      const uint32_t vb = smem_u32(v_smem);
      mbar_wait(smem_u32(ka_ready), 0);
      mbar_wait(smem_u32(v_full), 0);
      tc_fence_after();
      auto issue_s = [&](int t) {    // St[kv, 64 q] = K Q^T, A = K from TMEM
        const int j = t >> 1, stage = j & 1;
        if ((t & 1) == 0) { mbar_wait(smem_u32(&q_full[stage]), (j >> 1) & 1); tc_fence_after(); }
        const uint32_t qb = smem_u32(q_smem + stage * BTILE) + (t & 1) * 8192;
        const uint32_t buf = t & 1;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(T_ST + buf * 64, T_KA + k * 8, make_smem_desc(qb + kmajor_off(k), 16, 1024), idesc_s, k != 0);
        umma_commit<1>(smem_u32(&s_full[buf]));
      };
      auto issue_dp = [&](int t) {   // dPt[kv, 64 q] = V dO^T (SS)
        const int j = t >> 1, stage = j & 1;
        if ((t & 1) == 0) { mbar_wait(smem_u32(&do_full[stage]), (j >> 1) & 1); tc_fence_after(); }
        const uint32_t dob = smem_u32(do_smem + stage * BTILE) + (t & 1) * 8192;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss<1>(T_DPT, make_smem_desc(vb + kmajor_off(k), 16, 1024), make_smem_desc(dob + kmajor_off(k), 16, 1024),
                     idesc_s, k != 0);
        umma_commit<1>(smem_u32(dp_full));
      };
      issue_s(0);
      issue_dp(0);
      for (int t = 0; t < nsub; ++t) {
        if (t + 1 < nsub) issue_s(t + 1);
        const int j = t >> 1, stage = j & 1;
        const uint32_t buf = t & 1, rph = (t >> 1) & 1;
        const uint32_t qb = smem_u32(q_smem + stage * BTILE) + (t & 1) * 8192;
        const uint32_t dob = smem_u32(do_smem + stage * BTILE) + (t & 1) * 8192;
        mbar_wait(smem_u32(&p_ready[buf]), rph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dV[kv, d] += Pt[kv, 64 q] dO[64 q, d]
          umma_ts(T_DV, T_ST + buf * 64 + k * 8, make_smem_desc(dob + k * 2048, BHALF, 1024), idesc_g, (t | k) != 0);
        if (t & 1) umma_commit<1>(smem_u32(&do_empty[stage]));
        mbar_wait(smem_u32(&ds_ready[buf]), rph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dK[kv, d] += dSt[kv, 64 q] Q[64 q, d]
          umma_ts(T_DK, T_DPT + k * 8, make_smem_desc(qb + k * 2048, BHALF, 1024), idesc_g, (t | k) != 0);
        if (t & 1) umma_commit<1>(smem_u32(&q_empty[stage]));
        if (t + 1 < nsub) issue_dp(t + 1);   // the single dPt buffer is free once the dK MMA above has read dSt (in-order pipe)
      }
      umma_commit<1>(smem_u32(acc_done));
    }
  } else if (warp >= 4) {
    const uint32_t quad = warp & 3;
    const int wg = (warp - 4) >> 2;
    const uint32_t lane_base = (quad * 32u) << 16;
    const float c = p.scale_log2;
    const f32x2 c2 = f2_pack(c, c);
    const int kv = kv0 + quad * 32 + lane;
    const bool ok = kv < p.seq_k;
    {   // this thread's K row, d half `wg`: 128 contiguous bytes -> 32 TMEM columns of its lane
      const __nv_bfloat16* row = p.k + ((int64_t)bh * p.seq_k + kv) * 128 + wg * 64;
      uint32_t r[32];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        uint4 v4 = make_uint4(0u, 0u, 0u, 0u);
        if (ok) v4 = __ldg(reinterpret_cast<const uint4*>(row) + x);
        r[x * 4 + 0] = v4.x; r[x * 4 + 1] = v4.y; r[x * 4 + 2] = v4.z; r[x * 4 + 3] = v4.w;
      }
      tmem_st_x32(T_KA + lane_base + wg * 32, r);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(ka_ready));
    }
    const uint32_t buf = wg;
    for (int t = wg, it = 0; t < nsub; t += 2, ++it) {
      const uint32_t rph = it & 1;
      const int j = t >> 1, stage = j & 3;
      mbar_wait(smem_u32(&stats_full[stage]), (j >> 2) & 1);
      const float* lse_s = stats + stage * 256 + (t & 1) * 64;
      const float* dl_s = lse_s + 128;
      mbar_wait(smem_u32(&s_full[buf]), rph);
      tc_fence_after();
      f32x2 pv[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], pk[16];
        tmem_ld_x32(T_ST + lane_base + buf * 64 + hh * 32, r);
        tmem_ld_wait();
        const f32x2* nl = reinterpret_cast<const f32x2*>(lse_s + hh * 32);
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const f32x2 e = f2_exp2_mufu(f2_fma(f2_pack_bits(r[x], r[x + 1]), c2, nl[x >> 1]));
          pv[hh * 16 + (x >> 1)] = e;
          float p0, p1;
          f2_unpack(e, p0, p1);
          pk[x >> 1] = pack_bf16(p0, p1);
        }
        tmem_st_x16(T_ST + lane_base + buf * 64 + hh * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_ready[buf]));
      mbar_wait(smem_u32(dp_full), t & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32], pk[16];
        tmem_ld_x32(T_DPT + lane_base + hh * 32, r);
        tmem_ld_wait();
        const f32x2* dl2 = reinterpret_cast<const f32x2*>(dl_s + hh * 32);
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const f32x2 d = f2_mul(pv[hh * 16 + (x >> 1)], f2_sub(f2_pack_bits(r[x], r[x + 1]), dl2[x >> 1]));
          float d0, d1;
          f2_unpack(d, d0, d1);
          pk[x >> 1] = pack_bf16(d0, d1);
        }
        tmem_st_x16(T_DPT + lane_base + hh * 16, pk);   // packed dSt over columns this thread has already read
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&ds_ready[buf]));
        mbar_arrive(smem_u32(&stats_empty[stage]));
      }
    }
    mbar_wait(smem_u32(acc_done), 0);
    tc_fence_after();
    const int chalf = wg;
    const int64_t o = ((int64_t)bh * p.seq_k + kv) * 128 + chalf * 64;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const uint32_t tt = (which == 0 ? T_DV : T_DK) + lane_base + chalf * 64;
      __nv_bfloat16* dst = (which == 0 ? p.dv : p.dk) + o;
      const float mul = which == 0 ? 1.0f : p.scale;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t rr[32];
        tmem_ld_x32(tt + cc * 32, rr);
        tmem_ld_wait();
        if (ok) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + cc * 32);
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            uint4 qv;
            qv.x = pack_bf16(__uint_as_float(rr[x * 8 + 0]) * mul, __uint_as_float(rr[x * 8 + 1]) * mul);
            qv.y = pack_bf16(__uint_as_float(rr[x * 8 + 2]) * mul, __uint_as_float(rr[x * 8 + 3]) * mul);
            qv.z = pack_bf16(__uint_as_float(rr[x * 8 + 4]) * mul, __uint_as_float(rr[x * 8 + 5]) * mul);
            qv.w = pack_bf16(__uint_as_float(rr[x * 8 + 6]) * mul, __uint_as_float(rr[x * 8 + 7]) * mul);
            d4[x] = qv;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace dpipe

extern "C" int dpipe_attn_bwd(const dpipe_attn_bwd_args* a, void* stream) {
  using namespace dpipe;
  if (!a || !a->q || !a->k || !a->v || !a->o || !a->d_o || !a->lse || !a->delta || !a->dq || !a->dk || !a->dv)
    return fail(DPIPE_EINVAL, "dpipe_attn_bwd: null argument");
  if (a->batch <= 0 || a->heads <= 0 || a->seq_q <= 0 || a->seq_k <= 0) return fail(DPIPE_EINVAL, "dpipe_attn_bwd: empty problem");
  if (a->ldo % 8 || a->lddo % 8) return fail(DPIPE_EINVAL, "dpipe_attn_bwd: ldo/lddo must be multiples of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int B = a->batch, H = a->heads, Lq = a->seq_q, Lk = a->seq_k;
  {
    const int64_t warps = (int64_t)B * Lq * H;
    const int wpb = 8;
    attn_bwd_delta_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(a->o), a->ldo, reinterpret_cast<const __nv_bfloat16*>(a->d_o), a->lddo,
        a->delta, B, H, Lq);
  }
  CUtensorMap tq, tk, tv, tdo;
  const uint64_t bh = (uint64_t)B * H;
  int rc;
  if ((rc = make_tmap_3d_bf16(&tq, a->q, 128, Lq, bh, 128, (uint64_t)Lq * 128, 64, BT, 1))) return rc;
  if ((rc = make_tmap_3d_bf16(&tk, a->k, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, BT, 1))) return rc;
  if ((rc = make_tmap_3d_bf16(&tv, a->v, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, BT, 1))) return rc;
  {
    const uint64_t dims[4] = {128, (uint64_t)Lq, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)a->lddo, 128, (uint64_t)Lq * a->lddo};
    const uint32_t box[4] = {64, BT, 1, 1};
    if ((rc = make_tmap_4d_bf16(&tdo, a->d_o, dims, strides, box))) return rc;
  }
  static bool configured = false;
  if (!configured) {
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ4_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV3_SMEM_BYTES));
    configured = true;
  }
  AttnBwdParams p;
  p.lse = a->lse; p.delta = a->delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv);
  p.batch = B; p.heads = H; p.seq_q = Lq; p.seq_k = Lk;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  p.k = reinterpret_cast<const __nv_bfloat16*>(a->k);
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  p.d_o = reinterpret_cast<const __nv_bfloat16*>(a->d_o);
  p.lddo = a->lddo;
  // DPIPE_ATTN_BWD selects older variants for A/B measurements: 1 = un-pipelined v1, 2 = v2 (round 1), 3 = v2 dK/dV with
  // the LSE / D loader warps + v3 dQ (three TMEM score buffers, half-tile K / V slots), default 4 = the same dK/dV kernel +
  // v4 dQ (Q and dO resident in TMEM as the A operands), 5 = v3 dK/dV (K resident in TMEM) + v4 dQ
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("DPIPE_ATTN_BWD"); variant = (e && e[0] >= '1' && e[0] <= '5') ? e[0] - '0' : 4; }
  if (variant == 5) {
    CUtensorMap tk64, tv64;
    if ((rc = make_tmap_3d_bf16(&tk64, a->k, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, 64, 1))) return rc;
    if ((rc = make_tmap_3d_bf16(&tv64, a->v, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, 64, 1))) return rc;
    attn_bwd_dkv3_kernel<<<dim3((Lk + BT - 1) / BT, H, B), AB_THREADS, DKV3_SMEM_BYTES, s>>>(tq, tv, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
    attn_bwd_dq4_kernel<<<dim3((Lq + BT - 1) / BT, H, B), AB_THREADS, DQ4_SMEM_BYTES, s>>>(tk64, tv64, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
  } else if (variant == 4) {
    CUtensorMap tk64, tv64;
    if ((rc = make_tmap_3d_bf16(&tk64, a->k, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, 64, 1))) return rc;
    if ((rc = make_tmap_3d_bf16(&tv64, a->v, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, 64, 1))) return rc;
    attn_bwd_dkv2_kernel<true><<<dim3((Lk + BT - 1) / BT, H, B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
    attn_bwd_dq4_kernel<<<dim3((Lq + BT - 1) / BT, H, B), AB_THREADS, DQ4_SMEM_BYTES, s>>>(tk64, tv64, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
  } else if (variant == 3) {
    CUtensorMap tk64, tv64;
    if ((rc = make_tmap_3d_bf16(&tk64, a->k, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, 64, 1))) return rc;
    if ((rc = make_tmap_3d_bf16(&tv64, a->v, 128, Lk, bh, 128, (uint64_t)Lk * 128, 64, 64, 1))) return rc;
    attn_bwd_dkv2_kernel<true><<<dim3((Lk + BT - 1) / BT, H, B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
    attn_bwd_dq3_kernel<<<dim3((Lq + BT - 1) / BT, H, B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk64, tv64, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
  } else if (variant == 1) {
    attn_bwd_dkv_kernel<<<dim3((Lk + BT - 1) / BT, H, B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
    attn_bwd_dq_kernel<<<dim3((Lq + BT - 1) / BT, H, B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
  } else {
    attn_bwd_dkv2_kernel<false><<<dim3((Lk + BT - 1) / BT, H, B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
    attn_bwd_dq2_kernel<<<dim3((Lq + BT - 1) / BT, H, B), AB_THREADS, AB_SMEM_BYTES, s>>>(tq, tk, tv, tdo, p);
    DPIPE_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}
