// FlashAttention-style fused attention forward for sm_100a (head_dim 128, non-causal, bf16 in / fp32 softmax).
//
// One CTA owns two 128-row query tiles of one (batch, head) and streams K/V tiles of 128 keys:
//   warp 0      : TMA producer (Q once, K ring, V ring)
//   warp 1      : tcgen05.mma issuer.  S_i = Q_i K_j^T (SS, both K-major) into TMEM; O_i += P_i V_j with the
//                 A operand P_i read straight from TMEM (bf16, aliasing the front of S_i) and V MN-major in smem.
//   warp 2      : TMEM allocator
//   warps 4-7   : softmax for query tile 0 (one thread per row: tcgen05.ld S -> online softmax -> tcgen05.st P)
//   warps 8-11  : softmax for query tile 1
// The two query tiles ping-pong: while one warpgroup exponentiates, the tensor core works on the other tile.
// O accumulates in TMEM across the whole K loop; the running max is only refreshed (and O rescaled in TMEM)
// when it grows by more than 2^8, which keeps the rescale off the critical path and is exact after the final
// normalisation by the matching row sum.
//
// TMEM map (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_i = first 64 columns of S_i.
//
// Replaces torch SDPA / flash_attn as dispatched by the reference's Flux blocks
// (reference: diffusers attention inside models/flux.py:502,525; Wan: models/wan/attention.py:108-122).
#include <stdlib.h>

#include "host_util.h"
#include "sm100_common.cuh"

namespace dpipe {

constexpr int ATT_THREADS = 384;
constexpr int TQ = 128;   // query rows per tile
constexpr int TK = 128;   // keys per tile
constexpr int HD = 128;   // head dim
constexpr int HALF_TILE_BYTES = 128 * 64 * 2;   // one [128 rows][64 cols] swizzled sub-tile = 16 KiB
constexpr int TILE_BYTES = 2 * HALF_TILE_BYTES; // 32 KiB
constexpr int KV_STAGES = 2;
constexpr int ATT_SMEM_BYTES = 2 * TILE_BYTES + 2 * KV_STAGES * TILE_BYTES + 1024 + 256;

struct AttnFwdParams {
  __nv_bfloat16* o;
  int64_t ldo;
  float* lse;
  int batch, heads, seq_q, seq_k;
  float scale_log2;  // softmax scale * log2(e)
};

// POLY: of every 8 column pairs of a score row, this many take their exponential from the FMA-pipe polynomial instead of
// MUFU (0 = all MUFU).  The row maximum, the scaled subtraction and the row sum run on packed fp32 pairs / 3-input max.
template <int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                const __grid_constant__ CUtensorMap tma_v, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_smem = smem;                                   // 2 tiles
  uint8_t* k_smem = smem + 2 * TILE_BYTES;                  // KV_STAGES tiles
  uint8_t* v_smem = k_smem + KV_STAGES * TILE_BYTES;        // KV_STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + KV_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* k_full = bars + 1;        // [KV_STAGES]
  uint64_t* k_empty = k_full + KV_STAGES;
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;  // [2]  MMA -> softmax : S_i ready
  uint64_t* p_full = s_full + 2;           // [2]  softmax -> MMA : P_i written (O_i rescaled)
  uint64_t* o_done = p_full + 2;           // [2]  MMA -> softmax : last PV of tile i retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const int qblk = blockIdx.x;          // pair of query tiles
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * p.heads + head;
  const int q0 = qblk * 2 * TQ;
  const int nkv = (p.seq_k + TK - 1) / TK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(smem_u32(q_full), 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1);
      mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&v_full[s]), 1);
      mbar_init(smem_u32(&v_empty[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_full[i]), 4);  // one arrival per softmax warp
      mbar_init(smem_u32(&o_done[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      const uint32_t qb = smem_u32(q_full);
      mbar_expect_tx(qb, 2 * TILE_BYTES);
      for (int i = 0; i < 2; ++i)
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_q, qb, smem_u32(q_smem + i * TILE_BYTES + h * HALF_TILE_BYTES), h * 64, q0 + i * TQ, bh,
                      kEvictFirst);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        const int kv0 = j * TK;
        mbar_wait(smem_u32(&k_empty[stage]), phase ^ 1);
        const uint32_t kb = smem_u32(&k_full[stage]);
        mbar_expect_tx(kb, TILE_BYTES);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_k, kb, smem_u32(k_smem + stage * TILE_BYTES + h * HALF_TILE_BYTES), h * 64, kv0, bh,
                      kEvictLast);
        mbar_wait(smem_u32(&v_empty[stage]), phase ^ 1);
        const uint32_t vb = smem_u32(&v_full[stage]);
        mbar_expect_tx(vb, TILE_BYTES);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(&tma_v, vb, smem_u32(v_smem + stage * TILE_BYTES + h * HALF_TILE_BYTES), h * 64, kv0, bh,
                      kEvictLast);
        if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(TQ, TK, false, false);  // S = Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(TQ, HD, false, true);   // O += P V : A in TMEM, V MN-major
      auto issue_s = [&](int i, int kstage) {
        const uint32_t a_base = smem_u32(q_smem + i * TILE_BYTES);
        const uint32_t b_base = smem_u32(k_smem + kstage * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k >> 2) * HALF_TILE_BYTES + (k & 3) * 32;
          umma_ss<1>(tmem_base + i * 128, make_smem_desc(a_base + off, 16, 1024), make_smem_desc(b_base + off, 16, 1024),
                     idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit<1>(smem_u32(&s_full[i]));
      };
      auto issue_o = [&](int i, int vstage, bool accumulate) {
        const uint32_t b_base = smem_u32(v_smem + vstage * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < TK / 16; ++k) {
          // V sub-tile h holds d in [64h, 64h+64): LBO = distance between the two halves, SBO = 8 keys, 16 keys per step
          umma_ts(tmem_base + 256 + i * 128, tmem_base + i * 128 + k * 8,
                  make_smem_desc(b_base + k * 2048, HALF_TILE_BYTES, 1024), idesc_o, (accumulate || k != 0) ? 1u : 0u);
        }
      };
      mbar_wait(smem_u32(q_full), 0);
      // prologue: S_0(0), S_1(0)
      mbar_wait(smem_u32(&k_full[0]), 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit<1>(smem_u32(&k_empty[0]));
      int kstage = 1 % KV_STAGES;
      uint32_t kphase = (KV_STAGES == 1) ? 1 : 0;
      int vstage = 0;
      uint32_t vphase = 0;
      for (int j = 0; j < nkv; ++j) {
        const uint32_t pph = j & 1;
        mbar_wait(smem_u32(&v_full[vstage]), vphase);
        const bool more = (j + 1 < nkv);
        if (more) mbar_wait(smem_u32(&k_full[kstage]), kphase);
        for (int i = 0; i < 2; ++i) {
          mbar_wait(smem_u32(&p_full[i]), pph);
          tc_fence_after();
          issue_o(i, vstage, j > 0);
          if (!more) umma_commit<1>(smem_u32(&o_done[i]));
          if (more) issue_s(i, kstage);
        }
        umma_commit<1>(smem_u32(&v_empty[vstage]));
        if (more) umma_commit<1>(smem_u32(&k_empty[kstage]));
        if (++vstage == KV_STAGES) { vstage = 0; vphase ^= 1; }
        if (more) { if (++kstage == KV_STAGES) { kstage = 0; kphase ^= 1; } }
      }
    }
  } else if (warp >= 4) {
    // ================= softmax warpgroups =================
    const int i = (warp >= 8) ? 1 : 0;       // query tile
    const uint32_t quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int qrow = q0 + i * TQ + row_in_tile;
    const uint32_t lane_base = (quad * 32u) << 16;
    const uint32_t t_s = tmem_base + lane_base + i * 128;        // S_i (fp32, 128 cols)
    const uint32_t t_p = t_s;                                     // P_i (bf16 pairs, 64 cols) aliases S_i
    const uint32_t t_o = tmem_base + lane_base + 256 + i * 128;  // O_i
    const float c = p.scale_log2;
    float m = -INFINITY;   // running (possibly stale) max, scaled-log2 units
    float l = 0.f;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(smem_u32(&s_full[i]), j & 1);
      tc_fence_after();
      const int valid = min(TK, p.seq_k - j * TK);  // keys of this tile that exist
      // ---- one TMEM read of the whole 128-column row; everything below stays in registers ----
      uint32_t r[128];
      tmem_ld_x32(t_s, r);
      tmem_ld_x32(t_s + 32, r + 32);
      tmem_ld_x32(t_s + 64, r + 64);
      tmem_ld_x32(t_s + 96, r + 96);
      tmem_ld_wait();
      if (valid < TK) {   // last key tile: keys beyond seq_k get -inf (warp-uniform branch)
#pragma unroll
        for (int x = 0; x < 128; ++x)
          if (x >= valid) r[x] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int x = 0; x < 128; x += 8) {   // FMNMX3: two new columns per instruction
        mx0 = fmax3(mx0, __uint_as_float(r[x]), __uint_as_float(r[x + 1]));
        mx1 = fmax3(mx1, __uint_as_float(r[x + 2]), __uint_as_float(r[x + 3]));
        mx2 = fmax3(mx2, __uint_as_float(r[x + 4]), __uint_as_float(r[x + 5]));
        mx3 = fmax3(mx3, __uint_as_float(r[x + 6]), __uint_as_float(r[x + 7]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_new = fmaxf(m, mx * c);
      // ---- lazy rescale: only when some row of this warp grew by more than 2^8 ----
      if (j == 0) {
        m = m_new;
      } else {
        const bool need = (m_new - m) > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = exp2f(m - m_new);
          m = m_new;
          l *= alpha;
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t ro[32];
            tmem_ld_x32(t_o + cc * 32, ro);
            tmem_ld_wait();
#pragma unroll
            for (int x = 0; x < 32; ++x) ro[x] = __float_as_uint(__uint_as_float(ro[x]) * alpha);
            tmem_st_x32(t_o + cc * 32, ro);
          }
          tmem_st_wait();
        }
      }
      // ---- P = exp2(S*c - m) in place (bf16 pairs packed into the low half of r), fp32 row sum ----
      f32x2 sum_a = f2_pack(0.f, 0.f), sum_b = sum_a;
      const f32x2 c2 = f2_pack(c, c), neg_m2 = f2_pack(-m, -m);
#pragma unroll
      for (int x = 0; x < 128; x += 4) {
        const f32x2 a0 = f2_fma(f2_pack_bits(r[x], r[x + 1]), c2, neg_m2);
        const f32x2 a1 = f2_fma(f2_pack_bits(r[x + 2], r[x + 3]), c2, neg_m2);
        // pairs are numbered x / 2 = 0..63; within every group of 8 the first POLY go to the polynomial
        const f32x2 e0 = (((x >> 1) & 7) < POLY) ? f2_exp2_poly(a0) : f2_exp2_mufu(a0);
        const f32x2 e1 = ((((x >> 1) + 1) & 7) < POLY) ? f2_exp2_poly(a1) : f2_exp2_mufu(a1);
        sum_a = f2_add(sum_a, e0);
        sum_b = f2_add(sum_b, e1);
        float p0, p1, p2, p3;
        f2_unpack(e0, p0, p1);
        f2_unpack(e1, p2, p3);
        r[x >> 1] = pack_bf16(p0, p1);
        r[(x >> 1) + 1] = pack_bf16(p2, p3);
      }
      {
        float s0, s1, s2, s3;
        f2_unpack(sum_a, s0, s1);
        f2_unpack(sum_b, s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      tmem_st_x32(t_p, r);
      tmem_st_x32(t_p + 32, r + 32);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[i]));
    }

    // ---- epilogue: O / l -> bf16 -> global ; LSE ----
    mbar_wait(smem_u32(&o_done[i]), 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool row_ok = qrow < p.seq_q;
    __nv_bfloat16* orow = p.o + ((int64_t)b * p.seq_q + qrow) * p.ldo + head * HD;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t r[32];
      tmem_ld_x32(t_o + cc * 32, r);
      tmem_ld_wait();
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(orow + cc * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 qv;
          qv.x = pack_bf16(__uint_as_float(r[x * 8 + 0]) * inv_l, __uint_as_float(r[x * 8 + 1]) * inv_l);
          qv.y = pack_bf16(__uint_as_float(r[x * 8 + 2]) * inv_l, __uint_as_float(r[x * 8 + 3]) * inv_l);
          qv.z = pack_bf16(__uint_as_float(r[x * 8 + 4]) * inv_l, __uint_as_float(r[x * 8 + 5]) * inv_l);
          qv.w = pack_bf16(__uint_as_float(r[x * 8 + 6]) * inv_l, __uint_as_float(r[x * 8 + 7]) * inv_l);
          dst[x] = qv;
        }
      }
    }
    if (row_ok && p.lse) p.lse[(int64_t)bh * p.seq_q + qrow] = m + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace dpipe

extern "C" int dpipe_attn_fwd(const dpipe_attn_args* a, void* stream) {
  using namespace dpipe;
  if (!a || !a->q || !a->k || !a->v || !a->o) return fail(DPIPE_EINVAL, "dpipe_attn_fwd: null argument");
  if (a->batch <= 0 || a->heads <= 0 || a->seq_q <= 0 || a->seq_k <= 0)
    return fail(DPIPE_EINVAL, "dpipe_attn_fwd: empty problem");
  if (a->ldo % 8 != 0 || a->ldo < (int64_t)a->heads * HD) return fail(DPIPE_EINVAL, "dpipe_attn_fwd: bad ldo");
  CUtensorMap tq, tk, tv;
  const uint64_t bh = (uint64_t)a->batch * a->heads;
  int rc;
  if ((rc = make_tmap_3d_bf16(&tq, a->q, HD, a->seq_q, bh, HD, (uint64_t)a->seq_q * HD, 64, TQ, 1))) return rc;
  if ((rc = make_tmap_3d_bf16(&tk, a->k, HD, a->seq_k, bh, HD, (uint64_t)a->seq_k * HD, 64, TK, 1))) return rc;
  if ((rc = make_tmap_3d_bf16(&tv, a->v, HD, a->seq_k, bh, HD, (uint64_t)a->seq_k * HD, 64, TK, 1))) return rc;
  static bool configured = false;
  static int poly = 2;   // DPIPE_ATTN_FWD_POLY = 0 | 2 | 4: column pairs out of 8 exponentiated on the FMA pipe (A/B knob)
  if (!configured) {
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    const char* e = getenv("DPIPE_ATTN_FWD_POLY");
    if (e && (e[0] == '0' || e[0] == '2' || e[0] == '4') && e[1] == 0) poly = e[0] - '0';
    configured = true;
  }
  AttnFwdParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(a->o);
  p.ldo = a->ldo;
  p.lse = a->lse;
  p.batch = a->batch; p.heads = a->heads; p.seq_q = a->seq_q; p.seq_k = a->seq_k;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  dim3 grid((a->seq_q + 2 * TQ - 1) / (2 * TQ), a->heads, a->batch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (poly == 0) attn_fwd_kernel<0><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tq, tk, tv, p);
  else if (poly == 4) attn_fwd_kernel<4><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tq, tk, tv, p);
  else attn_fwd_kernel<2><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tq, tk, tv, p);
  DPIPE_CUDA_CHECK(cudaGetLastError());
  return 0;
}
