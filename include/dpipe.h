/*
 * dpipe.h — C ABI of libdpipe_b200.so: the sm_100a kernels and the 1F1B stage executor behind the
 * reference's Flux training hot path.
 *
 * The reference (tdrussell/diffusion-pipe) has no native code and no FFI of its own: every entry
 * point here replaces a *library call site* of the reference, cited per function as
 * "replaces: <reference file:line>" (paths relative to the reference repo root).
 *
 * Conventions
 *   - every function returns 0 on success or a negative DPIPE_E* code; dpipe_last_error() gives a
 *     thread-local human readable message.  Nothing throws across this boundary.
 *   - the caller owns all memory; pointers are raw CUDA device pointers unless a parameter says
 *     "host".  bf16 tensors are row-major with explicit leading dimensions in ELEMENTS.
 *   - all kernels are enqueued asynchronously on `stream` (a cudaStream_t passed as void*).
 *   - no hidden allocation: scratch space is passed in by the caller where needed.
 */
#ifndef DPIPE_H_
#define DPIPE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPIPE_OK 0
#define DPIPE_EINVAL (-1)   /* bad argument (shape/alignment/enum) */
#define DPIPE_ECUDA (-2)    /* a CUDA runtime/driver call failed */
#define DPIPE_ENOTSUP (-3)  /* device is not sm_100 / feature not compiled */
#define DPIPE_ESTATE (-4)   /* executor used out of order */

const char* dpipe_last_error(void);
/* returns the ABI version (bumped on any signature change) */
int dpipe_abi_version(void);
/* 0 if device `dev` is an sm_100 part this library can run on */
int dpipe_check_device(int dev);

/* ------------------------------------------------------------------------------------------ */
/* Dense bf16 GEMM on tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).    */
/*   D[M,N] (fp32 accumulators) = Aop[M,K] * Bop[N,K]^T, then a fused epilogue.                  */
/* replaces: cuBLASLt via nn.Linear inside the diffusers blocks invoked from                     */
/*           models/flux.py:502,525 (forward) and their autograd backward (dgrad/wgrad).         */
/* ------------------------------------------------------------------------------------------ */
enum {
  DPIPE_EPI_STORE = 0,     /* out = acc (+bias) (+out if accumulate)                                   */
  DPIPE_EPI_BIAS_GELU = 1, /* u = bf16(acc+bias); out2 = u (optional); out = gelu_tanh(u)               */
  DPIPE_EPI_GATE_RES = 2,  /* y = bf16(acc+bias); out2 = y (optional); out = aux + gate[b,:]*y          */
  DPIPE_EPI_QKV_ROPE = 3,  /* cols < n_qkv: bias -> per-head RMSNorm(q,k) -> RoPE -> head-major scatter; */
                           /* cols >= n_qkv: BIAS_GELU into (out,out2)                                  */
  DPIPE_EPI_MUL_GELU_GRAD = 4 /* out = acc * gelu_tanh'(aux)   (dgrad through GELU)                      */
};

typedef struct dpipe_qkv_epilogue {
  void* q;            /* bf16 [batch, heads, seq_total, 128]  (post norm + rope) */
  void* k;            /* bf16 same layout */
  void* v;            /* bf16 same layout (bias only) */
  void* qhat;         /* bf16 same layout: x * rstd before the norm weight (saved for backward), may be NULL */
  void* khat;         /* bf16 same layout, may be NULL */
  float* q_rstd;      /* fp32 [batch, heads, seq_total], may be NULL */
  float* k_rstd;      /* fp32 [batch, heads, seq_total], may be NULL */
  const void* q_norm_w; /* bf16 [128] */
  const void* k_norm_w; /* bf16 [128] */
  const float* rope_cos; /* fp32 [seq_total, 128] (interleaved-repeat layout of diffusers FluxPosEmbed) */
  const float* rope_sin; /* fp32 [seq_total, 128] */
  int heads;          /* H; the q, k, v column blocks are H*128 wide each */
  int seq_total;      /* joint sequence length (text + image) */
  int seq_offset;     /* position of this stream's first token in the joint sequence */
  int n_qkv;          /* 3*H*128; columns beyond it take the BIAS_GELU path */
  float eps;          /* RMSNorm epsilon (1e-6 in Flux) */
} dpipe_qkv_epilogue;

typedef struct dpipe_gemm_args {
  const void* A; int64_t lda; int a_mn; /* a_mn=0: A stored [M,K] (K contiguous); 1: stored [K,M] */
  const void* B; int64_t ldb; int b_mn; /* b_mn=0: B stored [N,K] (K contiguous); 1: stored [K,N] */
  int M, N, K;
  int epilogue;                 /* DPIPE_EPI_* */
  void* out; int64_t ldo;       /* bf16 [M, >=N] */
  void* out2; int64_t ldo2;     /* bf16, optional second output (see epilogue) */
  const void* bias;             /* bf16 [N] or NULL */
  const void* aux; int64_t ldaux;   /* bf16 [M,N] epilogue input (residual / pre-activation) */
  const void* gate; int64_t gate_stride; /* bf16 [batch, gate_stride], column n of batch b at gate[b*gate_stride+n] */
  int rows_per_batch;           /* rows of A that belong to one sample (for gate / qkv batch index) */
  int accumulate;               /* STORE only: out = bf16(float(out) + acc) */
  int cta_group;                /* 1: one CTA per 128x256 tile; 2: CTA pair, 256x256 tile (cta_group::2) */
  const dpipe_qkv_epilogue* qkv; /* required for DPIPE_EPI_QKV_ROPE */
} dpipe_gemm_args;

int dpipe_gemm_bf16(const dpipe_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Fused attention (FlashAttention-style, head_dim 128, non-causal) on tcgen05 + TMEM.           */
/* replaces: torch SDPA dispatched by the diffusers Flux attention processor invoked from        */
/*           models/flux.py:502,525; flash_attn_varlen_func at models/wan/attention.py:108-122.  */
/* ------------------------------------------------------------------------------------------ */
typedef struct dpipe_attn_args {
  const void* q;   /* bf16 [batch, heads, seq_q, 128] contiguous */
  const void* k;   /* bf16 [batch, heads, seq_k, 128] */
  const void* v;   /* bf16 [batch, heads, seq_k, 128] */
  void* o;         /* bf16 token-major [batch*seq_q, ldo]; head h occupies columns [128h, 128h+128) */
  int64_t ldo;
  float* lse;      /* fp32 [batch, heads, seq_q]: log2-domain logsumexp = max*scale*log2(e) + log2(sum); may be NULL (fwd) */
  int batch, heads, seq_q, seq_k;
  float scale;     /* softmax scale, 1/sqrt(128) for Flux */
} dpipe_attn_args;

int dpipe_attn_fwd(const dpipe_attn_args* args, void* stream);

typedef struct dpipe_attn_bwd_args {
  const void* q;   /* bf16 [batch, heads, seq_q, 128] (as given to the forward) */
  const void* k;   /* bf16 [batch, heads, seq_k, 128] */
  const void* v;   /* bf16 [batch, heads, seq_k, 128] */
  const void* o;   /* bf16 token-major forward output [batch*seq_q, ldo] */
  int64_t ldo;
  const void* d_o; /* bf16 token-major gradient of o [batch*seq_q, lddo] */
  int64_t lddo;
  const float* lse; /* fp32 [batch, heads, seq_q] from dpipe_attn_fwd (log2 domain) */
  float* delta;     /* fp32 [batch, heads, seq_q] scratch: rowsum(o * d_o), written by this call */
  void* dq;        /* bf16 [batch, heads, seq_q, 128] */
  void* dk;        /* bf16 [batch, heads, seq_k, 128] */
  void* dv;        /* bf16 [batch, heads, seq_k, 128] */
  int batch, heads, seq_q, seq_k;
  float scale;
} dpipe_attn_bwd_args;

/* backward of dpipe_attn_fwd; replaces flash-attn / SDPA backward reached via autograd from models/flux.py:502,525 */
int dpipe_attn_bwd(const dpipe_attn_bwd_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPIPE_H_ */
