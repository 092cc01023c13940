"""Tensor-level wrappers over the C ABI (include/dpipe.h).  torch is used only for device memory and
streams; every compute call goes to libdpipe_b200.so."""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, QkvEpilogue, check, lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req_bf16(t, name):
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise TypeError(f'{name} must be a CUDA bf16 tensor, got {t.dtype} on {t.device}')
    if t.stride(-1) != 1:
        raise ValueError(f'{name} must be contiguous in its last dimension')


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, epilogue=_lib.EPI_STORE, bias=None, out2=None,
         aux=None, gate=None, rows_per_batch=None, accumulate=False, cta_group=2, qkv=None,
         M=None, N=None, K=None):
    """D[M,N] = Aop[M,K] @ Bop[N,K]^T with a fused epilogue (see include/dpipe.h).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True).
    """
    _req_bf16(a, 'a')
    _req_bf16(b, 'b')
    assert a.dim() == 2 and b.dim() == 2
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    kb = b.shape[0] if b_mn else b.shape[1]
    if kb != K:
        raise ValueError(f'reduction dims differ: {K} vs {kb}')
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    _req_bf16(out, 'out')
    args = GemmArgs()
    args.A, args.lda, args.a_mn = _ptr(a), a.stride(0), int(a_mn)
    args.B, args.ldb, args.b_mn = _ptr(b), b.stride(0), int(b_mn)
    args.M, args.N, args.K = M, N, K
    args.epilogue = epilogue
    args.out, args.ldo = _ptr(out), out.stride(0)
    if out2 is not None:
        _req_bf16(out2, 'out2')
        args.out2, args.ldo2 = _ptr(out2), out2.stride(0)
    if bias is not None:
        _req_bf16(bias, 'bias')
        args.bias = _ptr(bias)
    if aux is not None:
        _req_bf16(aux, 'aux')
        args.aux, args.ldaux = _ptr(aux), aux.stride(0)
    if gate is not None:
        _req_bf16(gate, 'gate')
        args.gate, args.gate_stride = _ptr(gate), gate.stride(0)
    args.rows_per_batch = int(rows_per_batch) if rows_per_batch else M
    args.accumulate = int(bool(accumulate))
    args.cta_group = cta_group
    keep = None
    if qkv is not None:
        keep = qkv
        args.qkv = ctypes.pointer(qkv)
    check(lib().dpipe_gemm_bf16(ctypes.byref(args), _stream()), 'dpipe_gemm_bf16')
    del keep
    return out


def make_qkv_epilogue(q, k, v, q_norm_w, k_norm_w, rope_cos, rope_sin, heads, seq_total, seq_offset,
                      qhat=None, khat=None, q_rstd=None, k_rstd=None, eps=1e-6):
    e = QkvEpilogue()
    e.q, e.k, e.v = _ptr(q), _ptr(k), _ptr(v)
    e.qhat, e.khat = _ptr(qhat), _ptr(khat)
    e.q_rstd, e.k_rstd = _ptr(q_rstd), _ptr(k_rstd)
    e.q_norm_w, e.k_norm_w = _ptr(q_norm_w), _ptr(k_norm_w)
    assert rope_cos.dtype == torch.float32 and rope_sin.dtype == torch.float32
    assert rope_cos.is_contiguous() and rope_sin.is_contiguous()
    e.rope_cos, e.rope_sin = _ptr(rope_cos), _ptr(rope_sin)
    e.heads, e.seq_total, e.seq_offset = heads, seq_total, seq_offset
    e.n_qkv = 3 * heads * 128
    e.eps = eps
    return e


def attn_fwd(q, k, v, out=None, lse=None, scale=None):
    """softmax(q k^T * scale) v.  q: [B,H,Lq,128], k/v: [B,H,Lk,128] (bf16, contiguous).
    Returns (o [B*Lq, >=H*128] token-major, lse [B,H,Lq] fp32 in the log2 domain)."""
    from ._abi import AttnArgs
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v')):
        _req_bf16(t, n)
        if not t.is_contiguous() or t.shape[-1] != 128:
            raise ValueError(f'{n} must be contiguous [B,H,L,128]')
    B, H, Lq, _ = q.shape
    Lk = k.shape[2]
    if out is None:
        out = torch.empty((B * Lq, H * 128), dtype=torch.bfloat16, device=q.device)
    if lse is None:
        lse = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device)
    a = AttnArgs()
    a.q, a.k, a.v = _ptr(q), _ptr(k), _ptr(v)
    a.o, a.ldo = _ptr(out), out.stride(0)
    a.lse = _ptr(lse)
    a.batch, a.heads, a.seq_q, a.seq_k = B, H, Lq, Lk
    a.scale = float(scale if scale is not None else 128 ** -0.5)
    check(lib().dpipe_attn_fwd(ctypes.byref(a), _stream()), 'dpipe_attn_fwd')
    return out, lse


def attn_bwd(q, k, v, o, d_o, lse, scale=None, dq=None, dk=None, dv=None, delta=None):
    """Backward of attn_fwd.  o / d_o are token-major [B*Lq, >=H*128]; returns head-major (dq, dk, dv)."""
    from ._abi import AttnBwdArgs
    B, H, Lq, _ = q.shape
    Lk = k.shape[2]
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v')):
        _req_bf16(t, n)
        if not t.is_contiguous():
            raise ValueError(f'{n} must be contiguous')
    _req_bf16(o, 'o')
    _req_bf16(d_o, 'd_o')
    dq = torch.empty_like(q) if dq is None else dq
    dk = torch.empty_like(k) if dk is None else dk
    dv = torch.empty_like(v) if dv is None else dv
    delta = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device) if delta is None else delta
    a = AttnBwdArgs()
    a.q, a.k, a.v = _ptr(q), _ptr(k), _ptr(v)
    a.o, a.ldo = _ptr(o), o.stride(0)
    a.d_o, a.lddo = _ptr(d_o), d_o.stride(0)
    a.lse, a.delta = _ptr(lse), _ptr(delta)
    a.dq, a.dk, a.dv = _ptr(dq), _ptr(dk), _ptr(dv)
    a.batch, a.heads, a.seq_q, a.seq_k = B, H, Lq, Lk
    a.scale = float(scale if scale is not None else 128 ** -0.5)
    check(lib().dpipe_attn_bwd(ctypes.byref(a), _stream()), 'dpipe_attn_bwd')
    return dq, dk, dv
