"""Tensor-level wrappers over the C ABI (include/dpipe.h).  torch is used only for device memory and streams;
every compute call goes to libdpipe_b200.so, and a missing library raises (no fallback)."""
import ctypes

import torch

from . import _lib
from ._abi import AttnArgs, AttnBwdArgs, QkBwdArgs, WanNormBwdArgs, WanNormFwdArgs
from ._lib import GemmArgs, QkvEpilogue, check, lib

EPI_STORE = _lib.EPI_STORE
EPI_BIAS_GELU = _lib.EPI_BIAS_GELU
EPI_GATE_RES = _lib.EPI_GATE_RES
EPI_QKV_ROPE = _lib.EPI_QKV_ROPE
EPI_MUL_GELU_GRAD = _lib.EPI_MUL_GELU_GRAD

# kernel-launch counter (bench.py reports it as gpu_launches)
LAUNCHES = 0
# when a list, tensor-core kernels append (start_event, end_event, algorithmic_flops, kind) — bench.py's roofline
PROFILE = None


# when a list, weight-gradient work is queued instead of launched (split-backward / zero-bubble schedule): the engine
# sets it around a BackwardInput pass and runs the queued closures in the matching BackwardWeight pass
WGRAD_DEFER = None


def deferring():
    return WGRAD_DEFER is not None


def defer(fn):
    """run `fn` now, or queue it when a split-backward pass is active"""
    if WGRAD_DEFER is not None:
        WGRAD_DEFER.append(fn)
    else:
        fn()


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(e0, flops, kind, tag=None):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.append((e0, e1, flops, kind, tag))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req_bf16(t, name):
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise TypeError(f'{name} must be a CUDA bf16 tensor, got {t.dtype} on {t.device}')
    if t.stride(-1) != 1:
        raise ValueError(f'{name} must be contiguous in its last dimension')


def _req_f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise TypeError(f'{name} must be a contiguous CUDA fp32 tensor')


# ---------------------------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------------------------
def gemm(a, b, *, a_mn=False, b_mn=False, out=None, epilogue=EPI_STORE, bias=None, out2=None, aux=None, gate=None,
         rows_per_batch=None, accumulate=False, cta_group=2, qkv=None, M=None, N=None, K=None):
    """D[M,N] = Aop[M,K] @ Bop[N,K]^T with a fused epilogue (see include/dpipe.h).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True).
    """
    global LAUNCHES
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
    if qkv is not None:
        args.qkv = ctypes.pointer(qkv)
    _e = _prof_begin()
    check(lib().dpipe_gemm_bf16(ctypes.byref(args), _stream()), 'dpipe_gemm_bf16')
    _prof_end(_e, 2.0 * M * N * K, 'gemm', (M, N, K, int(a_mn), int(b_mn), epilogue, int(bool(accumulate))))
    LAUNCHES += 1
    return out


def make_qkv_epilogue(q, k, v, q_norm_w, k_norm_w, rope_cos, rope_sin, heads, seq_total, seq_offset, qhat=None,
                      khat=None, q_rstd=None, k_rstd=None, eps=1e-6):
    e = QkvEpilogue()
    e.q, e.k, e.v = _ptr(q), _ptr(k), _ptr(v)
    e.qhat, e.khat = _ptr(qhat), _ptr(khat)
    e.q_rstd, e.k_rstd = _ptr(q_rstd), _ptr(k_rstd)
    _req_bf16(q_norm_w, 'q_norm_w')
    _req_bf16(k_norm_w, 'k_norm_w')
    e.q_norm_w, e.k_norm_w = _ptr(q_norm_w), _ptr(k_norm_w)
    _req_f32(rope_cos, 'rope_cos')
    _req_f32(rope_sin, 'rope_sin')
    e.rope_cos, e.rope_sin = _ptr(rope_cos), _ptr(rope_sin)
    e.heads, e.seq_total, e.seq_offset = heads, seq_total, seq_offset
    e.n_qkv = 3 * heads * 128
    e.eps = eps
    return e


# ---------------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------------
def attn_fwd(q, k, v, out=None, lse=None, scale=None):
    """softmax(q k^T * scale) v.  q: [B,H,Lq,128], k/v: [B,H,Lk,128] (bf16, contiguous).
    Returns (o [B*Lq, >=H*128] token-major, lse [B,H,Lq] fp32 in the log2 domain)."""
    global LAUNCHES
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
    _e = _prof_begin()
    check(lib().dpipe_attn_fwd(ctypes.byref(a), _stream()), 'dpipe_attn_fwd')
    _prof_end(_e, 4.0 * B * H * Lq * Lk * 128, 'attn_fwd')
    LAUNCHES += 1
    return out, lse


def attn_bwd(q, k, v, o, d_o, lse, scale=None, dq=None, dk=None, dv=None, delta=None):
    """Backward of attn_fwd.  o / d_o are token-major [B*Lq, >=H*128]; returns head-major (dq, dk, dv)."""
    global LAUNCHES
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
    _e = _prof_begin()
    check(lib().dpipe_attn_bwd(ctypes.byref(a), _stream()), 'dpipe_attn_bwd')
    _prof_end(_e, 10.0 * B * H * Lq * Lk * 128, 'attn_bwd')
    LAUNCHES += 3
    return dq, dk, dv


# ---------------------------------------------------------------------------------------------------------------
# LayerNorm + modulation, gated residual, reductions
# ---------------------------------------------------------------------------------------------------------------
_ROW_CHUNK = None


def row_chunk():
    global _ROW_CHUNK
    if _ROW_CHUNK is None:
        _ROW_CHUNK = lib().dpipe_row_chunk()
    return _ROW_CHUNK


def nchunks(rows_per_batch):
    rc = row_chunk()
    return (rows_per_batch + rc - 1) // rc


LN_MULT_DIRECT = 1   # include/dpipe.h DPIPE_LN_MULT_DIRECT: `scale` is the multiplier itself (affine LayerNorm)
LN_ROUND_STEPS = 2   # DPIPE_LN_ROUND_STEPS: bf16 rounding after every elementwise op (Wan modulation on bf16 tensors)


def ln_modulate_fwd(x, scale, shift, batch, rows_per_batch, eps=1e-6, out=None, save_stats=True, flags=0):
    """out = LN(x) * bf16(1+scale[b]) + shift[b].  x: [batch*rows, D]; scale/shift: [batch, D] views (bf16)."""
    global LAUNCHES
    _req_bf16(x, 'x')
    _req_bf16(scale, 'scale')
    _req_bf16(shift, 'shift')
    D = x.shape[1]
    assert scale.stride(0) == shift.stride(0)
    if out is None:
        out = torch.empty((x.shape[0], D), dtype=torch.bfloat16, device=x.device)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        rstd = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _e = _prof_begin()
    check(lib().dpipe_ln_modulate_fwd_ex(_ptr(x), x.stride(0), _ptr(scale), _ptr(shift), scale.stride(0), _ptr(out),
                                         out.stride(0), _ptr(mean), _ptr(rstd), batch, rows_per_batch, D, eps, flags,
                                         _stream()), 'dpipe_ln_modulate_fwd')
    _prof_end(_e, 0.0, 'ln_fwd')
    LAUNCHES += 1
    return out, mean, rstd


def ln_modulate_bwd(dxn, x, scale, mean, rstd, batch, rows_per_batch, dres=None, dx=None, partials=None, flags=0):
    """Returns (dx, partials) with partials[batch][nchunk][2][D]: slot 0 = d scale, slot 1 = d shift."""
    global LAUNCHES
    _req_bf16(dxn, 'dxn')
    _req_bf16(x, 'x')
    _req_bf16(scale, 'scale')
    D = x.shape[1]
    nc = nchunks(rows_per_batch)
    if dx is None:
        dx = torch.empty((x.shape[0], D), dtype=torch.bfloat16, device=x.device)
    if partials is None:
        partials = torch.empty((batch, nc, 2, D), dtype=torch.float32, device=x.device)
    if dres is not None:
        _req_bf16(dres, 'dres')
    _e = _prof_begin()
    check(lib().dpipe_ln_modulate_bwd_ex(_ptr(dxn), dxn.stride(0), _ptr(x), x.stride(0), _ptr(scale), scale.stride(0),
                                         _ptr(mean), _ptr(rstd), _ptr(dres), dres.stride(0) if dres is not None else 0,
                                         _ptr(dx), dx.stride(0), _ptr(partials), batch, rows_per_batch, D, flags,
                                         _stream()), 'dpipe_ln_modulate_bwd')
    _prof_end(_e, 0.0, 'ln_bwd')
    LAUNCHES += 1
    return dx, partials


def gate_bwd(dx, y, gate, batch, rows_per_batch, dy=None, partials=None):
    """dy = bf16(gate*dx); partial slot 0 = sum dx*y (d gate per sample), slot 1 = sum dy (d bias)."""
    global LAUNCHES
    _req_bf16(dx, 'dx')
    _req_bf16(y, 'y')
    _req_bf16(gate, 'gate')
    D = y.shape[1]
    nc = nchunks(rows_per_batch)
    if dy is None:
        dy = torch.empty((y.shape[0], D), dtype=torch.bfloat16, device=y.device)
    if partials is None:
        partials = torch.empty((batch, nc, 2, D), dtype=torch.float32, device=y.device)
    _e = _prof_begin()
    check(lib().dpipe_gate_bwd(_ptr(dx), dx.stride(0), _ptr(y), y.stride(0), _ptr(gate), gate.stride(0), _ptr(dy),
                               dy.stride(0), _ptr(partials), batch, rows_per_batch, D, _stream()), 'dpipe_gate_bwd')
    _prof_end(_e, 0.0, 'gate_bwd')
    LAUNCHES += 1
    return dy, partials


def colreduce_finish(partials, per_sample0=None, per_sample1=None, summed0=None, summed1=None):
    """partials [batch, nchunk, nslot, D] fp32.  per_sample*: fp32 [batch, D] views (row stride = stride(0));
    summed*: fp32 [D]."""
    global LAUNCHES
    batch, nc, nslot, D = partials.shape
    ld0 = per_sample0.stride(0) if per_sample0 is not None else 0
    ld1 = per_sample1.stride(0) if per_sample1 is not None else 0
    _e = _prof_begin()
    check(lib().dpipe_colreduce_finish(_ptr(partials), batch, nc, nslot, D, _ptr(per_sample0), ld0, _ptr(per_sample1),
                                       ld1, _ptr(summed0), _ptr(summed1), _stream()), 'dpipe_colreduce_finish')
    _prof_end(_e, 0.0, 'finish')
    LAUNCHES += 1


def colsum(x, out=None):
    """fp32 column sums of a bf16 [rows, N] matrix."""
    global LAUNCHES
    _req_bf16(x, 'x')
    rows, N = x.shape
    nc = lib().dpipe_colsum_chunks(rows)
    partials = torch.empty((nc, N), dtype=torch.float32, device=x.device)
    if out is None:
        out = torch.empty(N, dtype=torch.float32, device=x.device)
    _e = _prof_begin()
    check(lib().dpipe_colsum(_ptr(x), x.stride(0), rows, N, _ptr(partials), _ptr(out), _stream()), 'dpipe_colsum')
    _prof_end(_e, 0.0, 'colsum')
    LAUNCHES += 2
    return out


FP8_FORMATS = {torch.float8_e4m3fn: 0, torch.float8_e5m2: 1}      # DPIPE_FP8_E4M3 / DPIPE_FP8_E5M2


def fp8_to_bf16(src, dst):
    """dst (bf16 [rows, cols] view, last dimension contiguous) = widen(src) for a 2-D fp8 matrix (float8_e4m3fn or
    float8_e5m2; exact).  The frozen base of a LoRA run stored in fp8 is expanded into a GEMM operand buffer with this
    before each GEMM that reads it (reference: autocast widens the float8 weight inside nn.Linear)."""
    global LAUNCHES
    if src.dtype not in FP8_FORMATS or not src.is_cuda or src.dim() != 2 or src.stride(1) != 1:
        raise TypeError(f'src must be a 2-D CUDA float8 tensor with a contiguous last dimension, got {src.dtype} {tuple(src.shape)}')
    _req_bf16(dst, 'dst')
    if dst.dim() != 2 or tuple(dst.shape) != tuple(src.shape):
        raise ValueError(f'dst {tuple(dst.shape)} must have the shape of src {tuple(src.shape)}')
    _e = _prof_begin()
    check(lib().dpipe_fp8_to_bf16(_ptr(src), src.stride(0), _ptr(dst), dst.stride(0), src.shape[0], src.shape[1],
                                  FP8_FORMATS[src.dtype], _stream()), 'dpipe_fp8_to_bf16')
    _prof_end(_e, 0.0, 'fp8_widen')
    LAUNCHES += 1
    return dst


def qknorm_rope_bwd(dq, dk, dv, qhat, khat, q_rstd, k_rstd, q_norm_w, k_norm_w, rope_cos, rope_sin, dqkv, dbias, dw,
                    batch, heads, seq_total, seq_offset, rows_per_batch):
    """Backward of the QKV_ROPE epilogue for one stream: writes token-major dqkv [batch*rows, 3*H*128] and
    accumulates dbias (fp32 [3*H*128]) and dw (fp32 [2,128]) — both must be zero-initialised by the caller."""
    global LAUNCHES
    a = QkBwdArgs()
    a.dq, a.dk, a.dv = _ptr(dq), _ptr(dk), _ptr(dv)
    a.qhat, a.khat = _ptr(qhat), _ptr(khat)
    a.q_rstd, a.k_rstd = _ptr(q_rstd), _ptr(k_rstd)
    a.q_norm_w, a.k_norm_w = _ptr(q_norm_w), _ptr(k_norm_w)
    a.rope_cos, a.rope_sin = _ptr(rope_cos), _ptr(rope_sin)
    _req_bf16(dqkv, 'dqkv')
    a.dqkv, a.ld = _ptr(dqkv), dqkv.stride(0)
    _req_f32(dbias, 'dbias')
    _req_f32(dw, 'dw')
    a.dbias, a.dw = _ptr(dbias), _ptr(dw)
    a.batch, a.heads, a.seq_total, a.seq_offset, a.rows_per_batch = batch, heads, seq_total, seq_offset, rows_per_batch
    _e = _prof_begin()
    check(lib().dpipe_qknorm_rope_bwd(ctypes.byref(a), _stream()), 'dpipe_qknorm_rope_bwd')
    _prof_end(_e, 0.0, 'qknorm_bwd')
    LAUNCHES += 1


def wan_norm_rope_fwd(projs, batch, seq, heads, cos=None, sin=None, eps=1e-6):
    """Wan q/k/v pre-processing (csrc/wan_norm.cu).  projs: up to three dicts {src [batch*seq, >=C] bf16 view, weight
    [C] bf16 or None, rope bool}.  Returns per projection (dst [batch, heads, seq, 128], xhat [batch*seq, C] or None,
    rstd [batch*seq] or None)."""
    global LAUNCHES
    a = WanNormFwdArgs()
    C = heads * 128
    outs = []
    for i, pj in enumerate(projs):
        src = pj['src']
        _req_bf16(src, 'src')
        dev = src.device
        dst = torch.empty((batch, heads, seq, 128), dtype=torch.bfloat16, device=dev)
        w = pj.get('weight')
        xhat = rstd = None
        if w is not None:
            _req_bf16(w, 'weight')
            xhat = torch.empty((batch * seq, C), dtype=torch.bfloat16, device=dev)
            rstd = torch.empty(batch * seq, dtype=torch.float32, device=dev)
        e = a.proj[i]
        e.src, e.ld, e.weight, e.dst, e.xhat, e.rstd = _ptr(src), src.stride(0), _ptr(w), _ptr(dst), _ptr(xhat), _ptr(rstd)
        e.rope = 1 if pj.get('rope') else 0
        outs.append((dst, xhat, rstd))
    a.nproj = len(projs)
    if cos is not None:
        _req_f32(cos, 'cos')
        _req_f32(sin, 'sin')
    a.cos, a.sin = _ptr(cos), _ptr(sin)
    a.batch, a.seq, a.heads, a.eps = batch, seq, heads, eps
    _e = _prof_begin()
    check(lib().dpipe_wan_norm_rope_fwd(ctypes.byref(a), _stream()), 'dpipe_wan_norm_rope_fwd')
    _prof_end(_e, 0.0, 'wan_norm_fwd')
    LAUNCHES += 1
    return outs


def wan_norm_rope_bwd(projs, batch, seq, heads, cos=None, sin=None):
    """Backward of wan_norm_rope_fwd.  projs: dicts {dy [batch, heads, seq, 128], dx [batch*seq, >=C] bf16 view (written),
    weight, xhat, rstd (None for v), rope}.  Returns, per projection, the fp32 d weight [C] (or None)."""
    global LAUNCHES
    a = WanNormBwdArgs()
    C = heads * 128
    rows = batch * seq
    nchunk = (rows + lib().dpipe_wan_norm_rows() - 1) // lib().dpipe_wan_norm_rows()
    parts = []
    for i, pj in enumerate(projs):
        dy, dx = pj['dy'], pj['dx']
        _req_bf16(dy, 'dy')
        _req_bf16(dx, 'dx')
        if not dy.is_contiguous():
            raise ValueError('dy must be a contiguous [batch, heads, seq, 128] tensor')
        w = pj.get('weight')
        part = torch.empty((1, nchunk, 1, C), dtype=torch.float32, device=dy.device) if w is not None else None
        e = a.proj[i]
        e.dy, e.xhat, e.rstd, e.weight = _ptr(dy), _ptr(pj.get('xhat')), _ptr(pj.get('rstd')), _ptr(w)
        e.dx, e.ld, e.dw_partials = _ptr(dx), dx.stride(0), _ptr(part)
        e.rope = 1 if pj.get('rope') else 0
        parts.append(part)
    a.nproj = len(projs)
    a.cos, a.sin = _ptr(cos), _ptr(sin)
    a.batch, a.seq, a.heads = batch, seq, heads
    _e = _prof_begin()
    check(lib().dpipe_wan_norm_rope_bwd(ctypes.byref(a), _stream()), 'dpipe_wan_norm_rope_bwd')
    _prof_end(_e, 0.0, 'wan_norm_bwd')
    LAUNCHES += 1
    dws = []
    for part in parts:
        if part is None:
            dws.append(None)
            continue
        dw = torch.empty(C, dtype=torch.float32, device=part.device)
        colreduce_finish(part, summed0=dw)
        dws.append(dw)
    return dws


def mod_fwd(temb, weight, bias):
    """mod[B, N] = W @ bf16(silu(temb)) + bias  (rank-B, HBM-bound kernel; temb [B, K] bf16)."""
    global LAUNCHES
    _req_bf16(temb, 'temb')
    _req_bf16(weight, 'weight')
    assert temb.is_contiguous() and weight.is_contiguous()
    B, K = temb.shape
    N = weight.shape[0]
    out = torch.empty((B, N), dtype=torch.bfloat16, device=temb.device)
    _e = _prof_begin()
    check(lib().dpipe_mod_fwd(_ptr(temb), _ptr(weight), _ptr(bias), _ptr(out), B, N, K, _stream()), 'dpipe_mod_fwd')
    _prof_end(_e, 0.0, 'mod_fwd')
    LAUNCHES += 1
    return out


def mod_bwd(dmod32, temb, weight, wgrad, accumulate, dtemb32):
    """Backward of mod_fwd.  dmod32 fp32 [B, N]; wgrad bf16 [N, K] or None; dtemb32 fp32 [B, K] is accumulated into.
    Returns dbias fp32 [N]."""
    global LAUNCHES
    _req_f32(dmod32, 'dmod32')
    _req_f32(dtemb32, 'dtemb32')
    B, K = temb.shape
    N = weight.shape[0]
    nc = lib().dpipe_mod_bwd_chunks(N)
    partials = torch.empty((nc, B, K), dtype=torch.float32, device=temb.device)
    dbias = torch.empty(N, dtype=torch.float32, device=temb.device)
    _e = _prof_begin()
    check(lib().dpipe_mod_bwd(_ptr(dmod32), dmod32.stride(0), _ptr(temb), _ptr(weight), _ptr(wgrad), int(bool(accumulate)),
                              _ptr(dbias), _ptr(partials), _ptr(dtemb32), B, N, K, _stream()), 'dpipe_mod_bwd')
    _prof_end(_e, 0.0, 'mod_bwd')
    LAUNCHES += 2
    return dbias


def mse_loss(out, target, mask=None, want_grad=True):
    """(loss fp32 scalar tensor, dout bf16 or None) for loss = mean((out-target)^2 * mask)."""
    global LAUNCHES
    _req_bf16(out, 'out')
    assert out.is_contiguous()
    target = target.to(torch.float32).contiguous()
    if mask is not None:
        mask = mask.to(torch.float32).expand_as(target).contiguous()
    ws = torch.empty(1024, dtype=torch.float32, device=out.device)
    loss = torch.empty((), dtype=torch.float32, device=out.device)
    dout = torch.empty_like(out) if want_grad else None
    _e = _prof_begin()
    check(lib().dpipe_mse_loss(_ptr(out), _ptr(target), _ptr(mask), out.numel(), _ptr(ws), _ptr(loss), _ptr(dout),
                               _stream()), 'dpipe_mse_loss')
    _prof_end(_e, 0.0, 'mse')
    LAUNCHES += 2
    return loss, dout


# ---------------------------------------------------------------------------------------------------------------
# optimizer-step tail (gradient norm / clipping) and micro-batch preparation
# ---------------------------------------------------------------------------------------------------------------
_GRAD_DTYPES = {torch.bfloat16: 0, torch.float32: 1}


def _tensor_lists(tensors):
    n = len(tensors)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
    numels = (ctypes.c_int64 * n)(*[t.numel() for t in tensors])
    dtypes = (ctypes.c_int * n)(*[_GRAD_DTYPES[t.dtype] for t in tensors])
    return ptrs, numels, dtypes


def grads_supported(tensors):
    return all(t.is_cuda and t.dtype in _GRAD_DTYPES and t.is_contiguous() for t in tensors)


def grad_sumsq(tensors):
    """fp32 device scalar: sum over `tensors` (contiguous CUDA bf16 / fp32) of sum(g^2) — one pass, no fp32 copies"""
    global LAUNCHES
    dev = tensors[0].device
    nl = max(1, (len(tensors) + 63) // 64)
    blocks = lib().dpipe_grad_sumsq_blocks()
    partials = torch.empty(blocks * nl, dtype=torch.float32, device=dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    ptrs, numels, dtypes = _tensor_lists(tensors)
    _e = _prof_begin()
    check(lib().dpipe_grad_sumsq(ptrs, numels, dtypes, len(tensors), _ptr(partials), partials.numel(), _ptr(out), _stream()),
          'dpipe_grad_sumsq')
    _prof_end(_e, 0.0, 'grad_sumsq')
    LAUNCHES += nl + 1
    return out


def grad_scale(tensors, coef):
    """g *= coef (fp32 device scalar) in place; the kernels return immediately when coef >= 1"""
    global LAUNCHES
    _req_f32(coef, 'coef')
    ptrs, numels, dtypes = _tensor_lists(tensors)
    _e = _prof_begin()
    check(lib().dpipe_grad_scale(ptrs, numels, dtypes, len(tensors), _ptr(coef), _stream()), 'dpipe_grad_scale')
    _prof_end(_e, 0.0, 'grad_scale')
    LAUNCHES += max(1, (len(tensors) + 63) // 64)


def noise_pack(x1, x0, t, pack):
    """(x_t, target) = ((1-t) x1 + t x0, x0 - x1) on the device, fp32, bit-identical to the separate host ops; x1/x0 any
    [bs, ...] shape; pack=True ([bs, c, h, w]) writes diffusers' 2x2-packed [bs, (h/2)(w/2), 4c] layout."""
    global LAUNCHES
    for a, n in ((x1, 'x1'), (x0, 'x0'), (t, 't')):
        _req_f32(a, n)
    assert x1.shape == x0.shape and x1.dim() >= 2 and t.numel() == x1.shape[0]
    bs = x1.shape[0]
    if pack:
        assert x1.dim() == 4, 'packing takes [bs, c, h, w] latents'
        c, frames, h, w = x1.shape[1], 1, x1.shape[2], x1.shape[3]
        shape = (bs, (h // 2) * (w // 2), 4 * c)
    else:
        c, frames, h, w = 1, 1, 1, x1.numel() // bs
        shape = tuple(x1.shape)
    xt = torch.empty(shape, dtype=torch.float32, device=x1.device)
    target = torch.empty(shape, dtype=torch.float32, device=x1.device)
    check(lib().dpipe_noise_pack(_ptr(x1), _ptr(x0), _ptr(t), _ptr(xt), _ptr(target), bs, c, frames, h, w, int(bool(pack)),
                                 _stream()), 'dpipe_noise_pack')
    LAUNCHES += 1
    return xt, target


def noise_on_device(x1, x0, t, pack, device):
    """`device_prepare_inputs`: the clean latents, the noise and the timesteps drawn on the host (same RNG stream and order
    as the reference's prepare_inputs) go to the device through pinned memory, asynchronously; the flow-matching mix, the
    target and the packing run there in one kernel.  Returns device tensors with the host path's bits."""
    def up(a):
        a = a.float().contiguous()
        if a.device.type == 'cpu':
            a = a.pin_memory() if not a.is_pinned() else a
        return a.to(device, non_blocking=True)
    return noise_pack(up(x1), up(x0), up(t), pack)
