// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (128B-swizzled tiles) -> shared memory ring -> tcgen05.mma (one elected thread) -> TMEM
//   (double-buffered 128x256 fp32 accumulators) -> 4 epilogue warps (tcgen05.ld) -> fused epilogue.
//
//   D[M,N] = Aop[M,K] * Bop[N,K]^T.  Each operand may be K-major (reduction dim contiguous) or
//   MN-major (non-reduced dim contiguous), so forward (X*W^T), dgrad (dY*W) and wgrad (dY^T*X) all run
//   on the same kernel without a transpose pass.
//
//   cta_group 1: one CTA per 128x256 output tile.   cta_group 2: a CTA pair (2 SMs, cluster of 2)
//   owns a 256x256 tile; each CTA stages its 128 rows of A and half of B (tcgen05.mma.cta_group::2).
//
// Replaces the cuBLASLt GEMMs behind nn.Linear in the reference's Flux blocks
// (reference: models/flux.py:502,525 and SURVEY.md section 2.3 shapes).
#include <stdlib.h>

#include "host_util.h"
#include "sm100_common.cuh"

namespace dpipe {

constexpr int BM = 128;  // accumulator rows per CTA (TMEM lanes)
constexpr int BN = 256;  // accumulator columns per tile (UMMA N)
constexpr int BK = 64;   // K per pipeline stage = one 128-byte swizzle row of bf16
constexpr int UK = 16;   // K per tcgen05.mma (kind::f16)
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int SUB_BYTES = 64 * 64 * 2;     // one 64(mn) x 64(k) MN-major sub-tile = 8 KiB
constexpr int GEMM_THREADS = 384;  // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps (2 per TMEM lane quarter)

template <int CG>
struct GemmCfg {
  static constexpr int B_ROWS = BN / CG;                   // rows of B staged by one CTA
  static constexpr int B_TILE_BYTES = B_ROWS * BK * 2;     // 32 KiB (CG=1) / 16 KiB (CG=2)
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = (CG == 1) ? 4 : 6;         // 192 KiB of operand staging either way
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int num_m_tiles;  // tiles of BM*CG rows
  int num_tiles;
  __nv_bfloat16* out;
  int64_t ldo;
  __nv_bfloat16* out2;
  int64_t ldo2;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* aux;
  int64_t ldaux;
  const __nv_bfloat16* gate;
  int64_t gate_stride;
  int rows_per_batch;
  int accumulate;
  // MN-major descriptor geometry (bytes): k-step per MMA, leading (64-column chunk) and stride (8-k group)
  // offsets.  Defaults 2048 / 8192 / 1024; DPIPE_DEBUG_MN_DESC="kstep,lbo,sbo" overrides for bring-up.
  uint32_t mn_kstep, mn_lbo, mn_sbo;
  dpipe_qkv_epilogue qkv;
};

// ---------------------------------------------------------------------------------------------
// epilogue helpers: every thread owns one accumulator row and walks it 32 columns at a time
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, bool ok, float* f) {
  if (ok) {
    const uint4* v = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 q = v[j];
      f[j * 8 + 0] = bf16_lo(q.x); f[j * 8 + 1] = bf16_hi(q.x);
      f[j * 8 + 2] = bf16_lo(q.y); f[j * 8 + 3] = bf16_hi(q.y);
      f[j * 8 + 4] = bf16_lo(q.z); f[j * 8 + 5] = bf16_hi(q.z);
      f[j * 8 + 6] = bf16_lo(q.w); f[j * 8 + 7] = bf16_hi(q.w);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = 0.f;
  }
}
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* p, const float* f) {
  uint4* v = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_bf16(f[j * 8 + 0], f[j * 8 + 1]);
    q.y = pack_bf16(f[j * 8 + 2], f[j * 8 + 3]);
    q.z = pack_bf16(f[j * 8 + 4], f[j * 8 + 5]);
    q.w = pack_bf16(f[j * 8 + 6], f[j * 8 + 7]);
    v[j] = q;
  }
}
__device__ __forceinline__ void load_f32x32(const float* p, float* f) {
  const float4* v = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 q = v[j];
    f[j * 4 + 0] = q.x; f[j * 4 + 1] = q.y; f[j * 4 + 2] = q.z; f[j * 4 + 3] = q.w;
  }
}

// bias + GELU(tanh) for one 32-column chunk; `col` is the global column of f[0]
__device__ __forceinline__ void epi_bias_gelu_chunk(const GemmParams& p, bool ok, int row, int col,
                                                    int out_col, float* acc) {
  float b[32];
  load_bf16x32(p.bias ? p.bias + col : nullptr, p.bias != nullptr, b);
  float u[32], h[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    u[j] = bf16_round(acc[j] + b[j]);
    h[j] = gelu_tanh(u[j]);
  }
  if (ok) {
    if (p.out2) store_bf16x32(p.out2 + (int64_t)row * p.ldo2 + out_col, u);
    store_bf16x32(p.out + (int64_t)row * p.ldo + out_col, h);
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, int row, int n0, uint32_t taddr, int half) {
  // `half` selects which 128 of the tile's 256 accumulator columns this warp drains
  const bool row_ok = row < p.M;
  const int batch = row_ok ? row / p.rows_per_batch : 0;

  if constexpr (EPI == DPIPE_EPI_QKV_ROPE) {
    if (n0 < p.qkv.n_qkv) {
      // ---- q / k / v columns: this 256-wide tile holds two heads of one of q, k, v ----
      const dpipe_qkv_epilogue& e = p.qkv;
      const int hd = e.heads * 128;
      const int which = n0 / hd;  // 0 q, 1 k, 2 v (tile never straddles: hd % 256 == 0)
      const int tok = row_ok ? row - batch * p.rows_per_batch : 0;
      const int pos = e.seq_offset + tok;
#pragma unroll 1
      for (int hh = half; hh <= half; ++hh) {
        const int col_h = n0 + hh * 128;  // global column of this head's first element
        if (col_h >= p.N) break;
        const int head = (col_h - which * hd) >> 7;
        const int64_t o = (((int64_t)batch * e.heads + head) * e.seq_total + pos) * 128;
        if (which == 2) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld_x32(taddr + hh * 128 + c * 32, r);
            tmem_ld_wait();
            float b[32], f[32];
            load_bf16x32(p.bias ? p.bias + col_h + c * 32 : nullptr, p.bias != nullptr, b);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]) + b[j];
            if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(e.v) + o + c * 32, f);
          }
        } else {
          // pass 1: sum of squares of the bf16-rounded projection (reference: RMSNorm on the bf16 Linear output)
          float ss = 0.f;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld_x32(taddr + hh * 128 + c * 32, r);
            tmem_ld_wait();
            float b[32];
            load_bf16x32(p.bias ? p.bias + col_h + c * 32 : nullptr, p.bias != nullptr, b);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = bf16_round(__uint_as_float(r[j]) + b[j]);
              ss += v * v;
            }
          }
          const float rstd = rsqrtf(ss * (1.0f / 128.0f) + e.eps);
          const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(which == 0 ? e.q_norm_w : e.k_norm_w);
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(which == 0 ? e.q : e.k);
          __nv_bfloat16* dst_hat = reinterpret_cast<__nv_bfloat16*>(which == 0 ? e.qhat : e.khat);
          float* dst_rstd = which == 0 ? e.q_rstd : e.k_rstd;
          if (row_ok && dst_rstd) dst_rstd[((int64_t)batch * e.heads + head) * e.seq_total + pos] = rstd;
          // pass 2: normalise, scale, rotate, store
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld_x32(taddr + hh * 128 + c * 32, r);
            tmem_ld_wait();
            float b[32], wv[32], xh[32], y[32];
            load_bf16x32(p.bias ? p.bias + col_h + c * 32 : nullptr, p.bias != nullptr, b);
            load_bf16x32(w + c * 32, true, wv);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = bf16_round(__uint_as_float(r[j]) + b[j]);
              xh[j] = bf16_round(v * rstd);
              y[j] = bf16_round(xh[j] * wv[j]);
            }
            if (row_ok) {
              if (dst_hat) store_bf16x32(dst_hat + o + c * 32, xh);
              float cs[32], sn[32], ro[32];
              load_f32x32(e.rope_cos + (int64_t)pos * 128 + c * 32, cs);
              load_f32x32(e.rope_sin + (int64_t)pos * 128 + c * 32, sn);
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                ro[j] = y[j] * cs[j] - y[j + 1] * sn[j];
                ro[j + 1] = y[j + 1] * cs[j + 1] + y[j] * sn[j + 1];
              }
              store_bf16x32(dst + o + c * 32, ro);
            }
          }
        }
      }
      return;
    }
  }

#pragma unroll 1
  for (int c = half * 4; c < half * 4 + 4; ++c) {
    const int col = n0 + c * 32;
    if (col >= p.N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_x32(taddr + c * 32, r);
    tmem_ld_wait();
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
    const bool ok = row_ok && (col + 32 <= p.N);
    const bool tail = row_ok && !ok;  // partial chunk at the right edge (N % 32 != 0)

    if constexpr (EPI == DPIPE_EPI_STORE) {
      float b[32];
      load_bf16x32(p.bias ? p.bias + col : nullptr, p.bias != nullptr && col + 32 <= p.N, b);
      if (p.bias && col + 32 > p.N) {
        for (int j = 0; j < 32 && col + j < p.N; ++j) b[j] = __bfloat162float(p.bias[col + j]);
      }
      __nv_bfloat16* dst = p.out + (int64_t)row * p.ldo + col;
      if (ok) {
        if (p.accumulate) {
          float old[32];
          load_bf16x32(dst, true, old);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += old[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += b[j];
        store_bf16x32(dst, acc);
      } else if (tail) {
        for (int j = 0; j < 32 && col + j < p.N; ++j) {
          float v = acc[j] + b[j];
          if (p.accumulate) v += __bfloat162float(dst[j]);
          dst[j] = __float2bfloat16_rn(v);
        }
      }
    } else if constexpr (EPI == DPIPE_EPI_BIAS_GELU) {
      epi_bias_gelu_chunk(p, ok, row, col, col, acc);
    } else if constexpr (EPI == DPIPE_EPI_QKV_ROPE) {
      // columns beyond n_qkv: the MLP half of the fused linear1 of a single-stream block
      epi_bias_gelu_chunk(p, ok, row, col, col - p.qkv.n_qkv, acc);
    } else if constexpr (EPI == DPIPE_EPI_GATE_RES) {
      float b[32], g[32], res[32], y[32], o[32];
      load_bf16x32(p.bias ? p.bias + col : nullptr, p.bias != nullptr, b);
      load_bf16x32(p.gate + (int64_t)batch * p.gate_stride + col, ok, g);
      load_bf16x32(p.aux + (int64_t)row * p.ldaux + col, ok, res);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        y[j] = bf16_round(acc[j] + b[j]);
        o[j] = res[j] + bf16_round(g[j] * y[j]);
      }
      if (ok) {
        if (p.out2) store_bf16x32(p.out2 + (int64_t)row * p.ldo2 + col, y);
        store_bf16x32(p.out + (int64_t)row * p.ldo + col, o);
      }
    } else if constexpr (EPI == DPIPE_EPI_MUL_GELU_GRAD) {
      float u[32], o[32];
      load_bf16x32(p.aux + (int64_t)row * p.ldaux + col, ok, u);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = acc[j] * gelu_tanh_grad(u[j]);
      if (ok) store_bf16x32(p.out + (int64_t)row * p.ldo + col, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int CG, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + Cfg::STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = warp_id_uniform();
  const uint32_t lane = lane_id();
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 8 * CG);  // one arrival per epilogue warp of every CTA in the group
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int num_kb = (p.K + BK - 1) / BK;
  const int group = blockIdx.x / CG;
  const int num_groups = gridDim.x / CG;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = group; t < p.num_tiles; t += num_groups) {
        const int mt = t % p.num_m_tiles, nt = t / p.num_m_tiles;
        const int m0 = (mt * CG + (int)cta_rank) * BM;
        const int nb0 = nt * BN + (int)cta_rank * Cfg::B_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t a_dst = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_dst = a_dst + A_TILE_BYTES;
          const int k0 = kb * BK;
          uint32_t bar = smem_u32(&full_bar[stage]);
          if constexpr (CG == 1) {
            mbar_expect_tx(bar, Cfg::STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d(&tma_a, bar, a_dst, k0, m0);
            } else {
              tma_load_2d(&tma_a, bar, a_dst, m0, k0);
              tma_load_2d(&tma_a, bar, a_dst + SUB_BYTES, m0 + 64, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d(&tma_b, bar, b_dst, k0, nb0, kEvictLast);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::B_ROWS / 64; ++j)
                tma_load_2d(&tma_b, bar, b_dst + j * SUB_BYTES, nb0 + 64 * j, k0, kEvictLast);
            }
          } else {
            // the pair's transaction bytes are all counted on the leader CTA's barrier
            if (cta_rank == 0) mbar_expect_tx(bar, 2 * Cfg::STAGE_BYTES);
            const uint32_t lbar = mapa_shared(bar, 0);
            if constexpr (!A_MN) {
              tma_load_2d_cg2(&tma_a, lbar, a_dst, k0, m0);
            } else {
              tma_load_2d_cg2(&tma_a, lbar, a_dst, m0, k0);
              tma_load_2d_cg2(&tma_a, lbar, a_dst + SUB_BYTES, m0 + 64, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d_cg2(&tma_b, lbar, b_dst, k0, nb0, kEvictLast);
            } else {
#pragma unroll
              for (int j = 0; j < Cfg::B_ROWS / 64; ++j)
                tma_load_2d_cg2(&tma_b, lbar, b_dst + j * SUB_BYTES, nb0 + 64 * j, k0, kEvictLast);
            }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA of the group only) =================
    if (cta_rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = group; t < p.num_tiles; t += num_groups, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // K-major: step 16 elements (32 B) inside the 128 B swizzle row.
            // MN-major: step 16 k-rows = two 1024 B swizzle atoms.
            const uint64_t adesc = A_MN ? make_smem_desc(a_base + k * p.mn_kstep, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc(b_base + k * p.mn_kstep, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc(b_base + k * 32, 16, 1024);
            umma_ss<CG>(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<CG>(smem_u32(&empty_bar[stage]));  // slot is free once these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(smem_u32(&tfull_bar[acc]));  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue warps: TMEM -> registers -> global =================
    const uint32_t quad = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int t = group; t < p.num_tiles; t += num_groups, ++it) {
      const int mt = t % p.num_m_tiles, nt = t / p.num_m_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      const int row = (mt * CG + (int)cta_rank) * BM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + acc * BN;
      epilogue_tile<EPI>(p, row, nt * BN, taddr, (int)((warp - 4) >> 2));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(smem_u32(&tempty_bar[acc]));
        else mbar_arrive_cluster(smem_u32(&tempty_bar[acc]), 0);
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
template <int CG, bool A_MN, bool B_MN, int EPI>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<CG>;
  auto kern = gemm_bf16_kernel<CG, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    DPIPE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  int groups = num_sms() / CG;
  if (groups > p.num_tiles) groups = p.num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CG);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DPIPE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return 0;
}

template <int CG>
static int dispatch(const dpipe_gemm_args* a, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                    cudaStream_t s) {
  const int layout = (a->a_mn ? 2 : 0) | (a->b_mn ? 1 : 0);
  switch (a->epilogue) {
    case DPIPE_EPI_STORE:
      if (layout == 0) return launch<CG, false, false, DPIPE_EPI_STORE>(ta, tb, p, s);
      if (layout == 1) return launch<CG, false, true, DPIPE_EPI_STORE>(ta, tb, p, s);
      if (layout == 3) return launch<CG, true, true, DPIPE_EPI_STORE>(ta, tb, p, s);
      break;
    case DPIPE_EPI_BIAS_GELU:
      if (layout == 0) return launch<CG, false, false, DPIPE_EPI_BIAS_GELU>(ta, tb, p, s);
      break;
    case DPIPE_EPI_GATE_RES:
      if (layout == 0) return launch<CG, false, false, DPIPE_EPI_GATE_RES>(ta, tb, p, s);
      break;
    case DPIPE_EPI_QKV_ROPE:
      if (layout == 0) return launch<CG, false, false, DPIPE_EPI_QKV_ROPE>(ta, tb, p, s);
      break;
    case DPIPE_EPI_MUL_GELU_GRAD:
      if (layout == 1) return launch<CG, false, true, DPIPE_EPI_MUL_GELU_GRAD>(ta, tb, p, s);
      break;
    default:
      break;
  }
  return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: epilogue %d is not built for operand layout a_mn=%d b_mn=%d",
              a->epilogue, a->a_mn, a->b_mn);
}

}  // namespace dpipe

extern "C" int dpipe_gemm_bf16(const dpipe_gemm_args* a, void* stream) {
  using namespace dpipe;
  if (!a) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: null args");
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: empty problem %dx%dx%d", a->M, a->N, a->K);
  if (a->cta_group != 1 && a->cta_group != 2) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: cta_group must be 1 or 2");
  if (!a->A || !a->B || !a->out) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: null operand");
  if (a->N % 8 != 0 || a->ldo % 8 != 0) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: N and ldo must be multiples of 8");
  if (a->rows_per_batch <= 0) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: rows_per_batch must be > 0");
  if (a->epilogue != DPIPE_EPI_STORE && a->N % 32 != 0)
    return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: fused epilogues need N %% 32 == 0");
  if ((a->epilogue == DPIPE_EPI_GATE_RES && (!a->gate || !a->aux)) ||
      (a->epilogue == DPIPE_EPI_MUL_GELU_GRAD && !a->aux))
    return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: epilogue %d needs aux/gate inputs", a->epilogue);

  GemmParams p = {};
  p.M = a->M; p.N = a->N; p.K = a->K;
  const int cg = a->cta_group;
  p.num_m_tiles = (a->M + BM * cg - 1) / (BM * cg);
  p.num_tiles = p.num_m_tiles * ((a->N + BN - 1) / BN);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2); p.ldo2 = a->ldo2;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a->bias);
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); p.ldaux = a->ldaux;
  p.gate = reinterpret_cast<const __nv_bfloat16*>(a->gate); p.gate_stride = a->gate_stride;
  p.rows_per_batch = a->rows_per_batch;
  p.accumulate = a->accumulate;
  p.mn_kstep = 2048; p.mn_lbo = SUB_BYTES; p.mn_sbo = 1024;
  if (const char* dbg = getenv("DPIPE_DEBUG_MN_DESC")) {
    unsigned ks, lbo, sbo;
    if (sscanf(dbg, "%u,%u,%u", &ks, &lbo, &sbo) == 3) { p.mn_kstep = ks; p.mn_lbo = lbo; p.mn_sbo = sbo; }
  }
  if (a->epilogue == DPIPE_EPI_QKV_ROPE) {
    if (!a->qkv) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: QKV_ROPE epilogue needs args->qkv");
    p.qkv = *a->qkv;
    const dpipe_qkv_epilogue& e = p.qkv;
    if (e.heads <= 0 || (e.heads * 128) % BN != 0 || e.n_qkv != 3 * e.heads * 128 || e.n_qkv > a->N)
      return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: bad qkv epilogue geometry (heads=%d n_qkv=%d N=%d)", e.heads, e.n_qkv, a->N);
    if (!e.q || !e.k || !e.v || !e.q_norm_w || !e.k_norm_w || !e.rope_cos || !e.rope_sin)
      return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: qkv epilogue has null pointers");
    if (e.seq_offset + a->rows_per_batch > e.seq_total)
      return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: qkv stream does not fit the joint sequence");
    if (e.n_qkv < a->N && !a->out) return fail(DPIPE_EINVAL, "dpipe_gemm_bf16: fused linear1 needs an MLP output");
  }

  CUtensorMap ta, tb;
  int rc;
  // K-major operand X[R,K]: dims {K, R}, box {64, rows}.  MN-major operand stored [K,R]: dims {R, K}, box {64, 64}.
  if (!a->a_mn) rc = make_tmap_2d_bf16(&ta, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, BM);
  else          rc = make_tmap_2d_bf16(&ta, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BK);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap_2d_bf16(&tb, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BK, BN / cg);
  else          rc = make_tmap_2d_bf16(&tb, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BK);
  if (rc) return rc;

  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return cg == 1 ? dispatch<1>(a, ta, tb, p, s) : dispatch<2>(a, ta, tb, p, s);
}
