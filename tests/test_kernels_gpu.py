"""GPU: every C-ABI kernel against a plain PyTorch fp32 computation on the same bf16 inputs.
Tolerances: bf16 outputs -> 2e-2 of the reference's max magnitude per element (one bf16 ulp at the top of the range is
0.8%; accumulation order differs), fp32 statistics -> 1e-4."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(out, ref, tol=2e-2):
    out, ref = out.float(), ref.float()
    assert not torch.isnan(out).any()
    scale = ref.abs().max().item() + 1e-12
    err = (out - ref).abs().max().item()
    assert err <= tol * scale, f'max err {err} vs scale {scale}'


@pytest.fixture(scope='module')
def ops():
    from diffusion_pipe_b200 import ops as o
    return o


@pytest.mark.parametrize('cg', [1, 2])
@pytest.mark.parametrize('shape', [(512, 768, 512), (304, 264, 200), (64, 64, 3072), (1, 512, 256)])
def test_gemm_layouts(ops, cg, shape):
    M, N, K = shape
    torch.manual_seed(0)
    A = torch.randn(M, K, device='cuda').bfloat16()
    B = torch.randn(N, K, device='cuda').bfloat16()
    ref = A.float() @ B.float().t()
    _check(ops.gemm(A, B, cta_group=cg), ref)
    _check(ops.gemm(A, B.t().contiguous(), b_mn=True, cta_group=cg), ref)
    if M % 8 == 0:
        _check(ops.gemm(A.t().contiguous(), B.t().contiguous(), a_mn=True, b_mn=True, cta_group=cg), ref)


def test_gemm_strided_operands_and_accumulate(ops):
    torch.manual_seed(1)
    M, N, K = 256, 512, 384
    big_a = torch.randn(M, K + 128, device='cuda').bfloat16()
    big_w = torch.randn(N, K + 64, device='cuda').bfloat16()
    A, W = big_a[:, 64:64 + K], big_w[:, :K]
    bias = torch.randn(N, device='cuda').bfloat16()
    out = torch.randn(M, N + 64, device='cuda').bfloat16()
    old = out.clone()
    ops.gemm(A, W, out=out[:, :N], bias=bias, accumulate=True)
    _check(out[:, :N], A.float() @ W.float().t() + bias.float() + old[:, :N].float())
    assert torch.equal(out[:, N:], old[:, N:])


def test_gemm_epilogues(ops):
    torch.manual_seed(2)
    Bsz, L, N, K = 2, 192, 512, 256
    M = Bsz * L
    A = torch.randn(M, K, device='cuda').bfloat16()
    W = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
    bias = torch.randn(N, device='cuda').bfloat16()
    u = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    h = ops.gemm(A, W, bias=bias, epilogue=ops.EPI_BIAS_GELU, out2=u)
    uref = (A.float() @ W.float().t() + bias.float()).bfloat16()
    _check(u, uref)
    _check(h, torch.nn.functional.gelu(uref.float(), approximate='tanh'))
    res = torch.randn(M, N, device='cuda').bfloat16()
    gate = torch.randn(Bsz, 3 * N, device='cuda').bfloat16()
    y = torch.empty_like(u)
    o = ops.gemm(A, W, bias=bias, epilogue=ops.EPI_GATE_RES, aux=res, gate=gate[:, N:2 * N], out2=y, rows_per_batch=L)
    _check(y, uref)
    _check(o, res.float() + (gate[:, N:2 * N].float().repeat_interleave(L, 0) * uref.float()).bfloat16().float())
    Wt = (torch.randn(K, N, device='cuda') * 0.1).bfloat16()
    uu = torch.randn(M, N, device='cuda').bfloat16()
    og = ops.gemm(A, Wt, b_mn=True, epilogue=ops.EPI_MUL_GELU_GRAD, aux=uu)
    uf = uu.float().requires_grad_(True)
    torch.nn.functional.gelu(uf, approximate='tanh').sum().backward()
    _check(og, (A.float() @ Wt.float()) * uf.grad)


def _ref_attn(q, k, v, scale):
    s = torch.einsum('bhqd,bhkd->bhqk', q.float(), k.float()) * scale
    return torch.einsum('bhqk,bhkd->bhqd', torch.softmax(s, -1), v.float()), torch.logsumexp(s, -1)


@pytest.mark.parametrize('shape', [(1, 2, 256, 256), (2, 2, 96, 96), (1, 2, 300, 200), (1, 1, 128, 640)])
def test_attention_forward_backward(ops, shape):
    B, H, Lq, Lk = shape
    torch.manual_seed(3)
    scale = 128 ** -0.5
    q = torch.randn(B, H, Lq, 128, device='cuda').bfloat16()
    k = torch.randn(B, H, Lk, 128, device='cuda').bfloat16()
    v = torch.randn(B, H, Lk, 128, device='cuda').bfloat16()
    ld = H * 128 + 64
    o = torch.zeros(B * Lq, ld, device='cuda', dtype=torch.bfloat16)
    _, lse = ops.attn_fwd(q, k, v, out=o)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    oref, lseref = _ref_attn(qf, kf, vf, scale)
    _check(o[:, :H * 128].view(B, Lq, H, 128).permute(0, 2, 1, 3), oref)
    _check(lse * math.log(2.0), lseref, 1e-3)
    assert torch.all(o[:, H * 128:] == 0)
    d_o = (torch.randn(B * Lq, ld, device='cuda') * 0.5).bfloat16()
    dq, dk, dv = ops.attn_bwd(q, k, v, o, d_o, lse)
    oref.backward(d_o[:, :H * 128].float().view(B, Lq, H, 128).permute(0, 2, 1, 3))
    _check(dq, qf.grad)
    _check(dk, kf.grad)
    _check(dv, vf.grad)


def test_attention_rescale_path(ops):
    torch.manual_seed(4)
    q = (torch.randn(1, 2, 512, 128, device='cuda') * 6).bfloat16()   # large logits: the running max keeps growing
    k = torch.randn(1, 2, 1024, 128, device='cuda').bfloat16()
    v = torch.randn(1, 2, 1024, 128, device='cuda').bfloat16()
    o, lse = ops.attn_fwd(q, k, v)
    oref, lseref = _ref_attn(q, k, v, 128 ** -0.5)
    _check(o.view(1, 512, 2, 128).permute(0, 2, 1, 3), oref)
    _check(lse * math.log(2.0), lseref, 1e-3)


@pytest.mark.parametrize('D', [512, 5120])      # 5120 = Wan-14B width: 640-thread CTAs, the <2, 640, 1> backward variant (512 -> <2, 384, 2>)
def test_ln_modulate_and_gate_backward(ops, D):
    torch.manual_seed(5)
    B, L = 2, 50
    x = torch.randn(B * L, D, device='cuda').bfloat16()
    mod = (0.3 * torch.randn(B, 3 * D, device='cuda')).bfloat16()
    scale, shift, gate = mod[:, D:2 * D], mod[:, :D], mod[:, 2 * D:]
    out, mean, rstd = ops.ln_modulate_fwd(x, scale, shift, B, L)
    xf = x.float().view(B, L, D).requires_grad_(True)
    scf = scale.float().requires_grad_(True)
    shf = shift.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xf, (D,), eps=1e-6) * (1 + scf).bfloat16().float()[:, None] + shf[:, None]
    _check(out.view(B, L, D), ref)
    _check(mean, xf.mean(-1).reshape(-1), 1e-4)
    dxn = torch.randn(B * L, D, device='cuda').bfloat16()
    dres = torch.randn(B * L, D, device='cuda').bfloat16()
    dx, part = ops.ln_modulate_bwd(dxn, x, scale, mean, rstd, B, L, dres=dres)
    dmod = torch.zeros(B, 2 * D, device='cuda')
    ops.colreduce_finish(part, per_sample0=dmod[:, :D], per_sample1=dmod[:, D:])
    # reference gradient with (1+scale) treated as the bf16 constant the kernel uses
    one_plus = (1 + scale.float()).bfloat16().float()
    xf2 = x.float().view(B, L, D).requires_grad_(True)
    xhat = torch.nn.functional.layer_norm(xf2, (D,), eps=1e-6)
    (xhat * one_plus[:, None] * dxn.float().view(B, L, D)).sum().backward()
    _check(dx.view(B, L, D), xf2.grad + dres.float().view(B, L, D))
    _check(dmod[:, :D], (dxn.float().view(B, L, D) * xhat.detach()).sum(1), 1e-3)
    _check(dmod[:, D:], dxn.float().view(B, L, D).sum(1), 1e-3)
    y = torch.randn(B * L, D, device='cuda').bfloat16()
    dy, part = ops.gate_bwd(dxn, y, gate, B, L)
    dg = torch.zeros(B, D, device='cuda')
    db = torch.zeros(D, device='cuda')
    ops.colreduce_finish(part, per_sample0=dg, summed1=db)
    dyref = (gate.float().repeat_interleave(L, 0) * dxn.float()).bfloat16().float()
    _check(dy, dyref)
    _check(dg, (dxn.float() * y.float()).view(B, L, D).sum(1), 1e-3)
    _check(db, dyref.sum(0), 1e-3)
    _check(ops.colsum(y), y.float().sum(0), 1e-3)


def test_qkv_rope_epilogue_and_its_backward(ops):
    torch.manual_seed(6)
    Bsz, L, H, K = 2, 160, 2, 256
    Ltot, off = 224, 64
    N = 3 * H * 128
    M = Bsz * L
    A = torch.randn(M, K, device='cuda').bfloat16()
    W = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
    bias = torch.randn(N, device='cuda').bfloat16()
    wq = (1 + 0.1 * torch.randn(128, device='cuda')).bfloat16()
    wk = (1 + 0.1 * torch.randn(128, device='cuda')).bfloat16()
    ang = torch.rand(Ltot, 64, device='cuda') * 6.28
    cos = torch.cos(ang).repeat_interleave(2, dim=1).contiguous()
    sin = torch.sin(ang).repeat_interleave(2, dim=1).contiguous()
    shp = (Bsz, H, Ltot, 128)
    q, k, v, qh, kh = (torch.zeros(shp, device='cuda', dtype=torch.bfloat16) for _ in range(5))
    qr, kr = torch.zeros(Bsz, H, Ltot, device='cuda'), torch.zeros(Bsz, H, Ltot, device='cuda')
    e = ops.make_qkv_epilogue(q, k, v, wq, wk, cos, sin, H, Ltot, off, qh, kh, qr, kr)
    ops.gemm(A, W, bias=bias, epilogue=ops.EPI_QKV_ROPE, out=A, rows_per_batch=L, qkv=e)
    pre = (A.float() @ W.float().t() + bias.float()).bfloat16().float().requires_grad_(True)
    qkv = pre.view(Bsz, L, 3, H, 128).permute(2, 0, 3, 1, 4)
    c, s = cos[off:off + L][None, None], sin[off:off + L][None, None]

    def norm_rope(x, w):
        rstd = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)
        y = x * rstd * w.float()
        yr = torch.stack([-y[..., 1::2], y[..., 0::2]], -1).flatten(-2)
        return y * c + yr * s
    qref, kref = norm_rope(qkv[0], wq), norm_rope(qkv[1], wk)
    sl = slice(off, off + L)
    _check(q[:, :, sl], qref)
    _check(k[:, :, sl], kref)
    _check(v[:, :, sl], qkv[2])
    # backward of the epilogue
    dq, dk, dv = (torch.zeros(shp, device='cuda', dtype=torch.bfloat16) for _ in range(3))
    for t in (dq, dk, dv):
        t[:, :, sl] = torch.randn(Bsz, H, L, 128, device='cuda').bfloat16()
    dqkv = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    dbias = torch.zeros(N, device='cuda')
    dw = torch.zeros(2, 128, device='cuda')
    ops.qknorm_rope_bwd(dq, dk, dv, qh, kh, qr, kr, wq, wk, cos, sin, dqkv, dbias, dw, Bsz, H, Ltot, off, L)
    wqf, wkf = wq.float().requires_grad_(True), wk.float().requires_grad_(True)
    loss = (norm_rope(qkv[0], wqf) * dq[:, :, sl].float()).sum() + (norm_rope(qkv[1], wkf) * dk[:, :, sl].float()).sum() \
        + (qkv[2] * dv[:, :, sl].float()).sum()
    loss.backward()
    _check(dqkv, pre.grad, 3e-2)
    _check(dbias, pre.grad.sum(0), 2e-2)
    _check(dw[0], wqf.grad, 2e-2)
    _check(dw[1], wkf.grad, 2e-2)


def test_mse_loss(ops):
    torch.manual_seed(7)
    out = torch.randn(2, 64, 64, device='cuda').bfloat16()
    tgt = torch.randn(2, 64, 64, device='cuda')
    mask = (torch.rand(2, 64, 64, device='cuda') > 0.3).float()
    for m in (None, mask):
        loss, dout = ops.mse_loss(out, tgt, m)
        of = out.float().requires_grad_(True)
        l = torch.nn.functional.mse_loss(of, tgt, reduction='none')
        if m is not None:
            l = l * m
        l = l.mean()
        l.backward()
        assert abs(loss.item() - l.item()) <= 1e-5 * max(1, abs(l.item()))
        _check(dout, of.grad)


def test_grad_sumsq_and_scale_match_torch():
    """csrc/step_tail.cu: the multi-tensor squared norm / conditional scale behind engine._clip_grad_norm
    (utils/patches.py:175-246) — > 64 tensors (several launches), unaligned bf16 views, an fp32 tensor, tiny tensors"""
    from diffusion_pipe_b200 import ops
    torch.manual_seed(3)
    big = torch.randn(3_000_017, device='cuda').bfloat16()
    ts = [big[1:], torch.randn(1 << 20, device='cuda').bfloat16(), torch.randn(7, device='cuda').bfloat16(),
          torch.randn(12345, device='cuda'), torch.randn(1, device='cuda').bfloat16()]
    ts += [torch.randn(1000 + 13 * i, device='cuda').bfloat16() for i in range(140)]
    ts = [t.contiguous() if not t.is_contiguous() else t for t in ts]
    assert ops.grads_supported(ts)
    want = sum(t.double().pow(2).sum() for t in ts).item()
    got = ops.grad_sumsq(ts).item()
    assert abs(got - want) / want <= 1e-5, (got, want)
    assert ops.grad_sumsq(ts).item() == got                       # fixed reduction order: bitwise repeatable
    before = [t.clone() for t in ts]
    ops.grad_scale(ts, torch.ones(1, device='cuda'))              # coef >= 1: untouched
    assert all(torch.equal(a, b) for a, b in zip(ts, before))
    coef = torch.full((1,), 0.37, device='cuda')
    ops.grad_scale(ts, coef)
    for a, b in zip(ts, before):
        assert torch.equal(a, (b.float() * coef).to(b.dtype))


@pytest.mark.parametrize('shape,pack', [((2, 16, 32, 48), True), ((1, 16, 128, 128), True), ((2, 16, 3, 8, 12), False)])
def test_noise_pack_is_bit_identical_to_the_host_ops(shape, pack):
    """x_t = (1 - t) x_1 + t x_0, target = x_0 - x_1 and the 2x2 packing on the device give the bits of the reference's host
    ops (models/flux.py:368-378): no FMA contraction, same rounding points"""
    from diffusion_pipe_b200 import ops
    from diffusion_pipe_b200.flux import pack_latents
    g = torch.Generator().manual_seed(5)
    x1, x0 = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    t = torch.sigmoid(torch.randn(shape[0], generator=g))
    te = t.view(-1, *([1] * (len(shape) - 1)))
    xt_h, tg_h = (1 - te) * x1 + te * x0, x0 - x1
    if pack:
        xt_h, tg_h = pack_latents(xt_h), pack_latents(tg_h)
    xt_d, tg_d = ops.noise_pack(x1.cuda(), x0.cuda(), t.cuda(), pack)
    assert torch.equal(xt_d.cpu(), xt_h) and torch.equal(tg_d.cpu(), tg_h)
