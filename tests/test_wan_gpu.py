"""GPU: the Wan kernels and a small Wan-architecture model through the reference-facing plugin API (WanPipeline.to_layers
/ prepare_inputs-shaped tuples / get_loss_fn) against the oracle (oracle/wan_ref.py, pinned to the reference's own
model code by tests/test_oracle_wan_golden.py), and against the stored output of that reference code itself
(tests/golden/wan_golden.pt).

Tolerance: loss within 1e-3 relative of the oracle with the reference's bf16 rounding points emulated, 5e-3 of the
pure-fp32 oracle; parameter gradients within 5e-2 relative L2 of the emulating oracle."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))

CFG = {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 2, 'text_dim': 64, 'text_len': 16}


def test_full_width_rmsnorm_rope_kernels_match_torch():
    """csrc/wan_norm.cu forward and backward vs the fp32 formula (WanRMSNorm + rope_apply, models/wan/model.py:41-87)"""
    from diffusion_pipe_b200 import ops
    from oracle import wan_ref as W
    torch.manual_seed(0)
    B, L, H = 2, 37, 2
    C = H * 128
    qkv = torch.randn(B * L, 3 * C, device='cuda').bfloat16()
    wq = (1 + 0.1 * torch.randn(C, device='cuda')).bfloat16()
    wk = (1 + 0.1 * torch.randn(C, device='cuda')).bfloat16()
    cos, sin = (t.cuda().contiguous() for t in W.wan_rope_tables((1, 1, L)))
    (q, xhq, rq), (k, xhk, rk), (v, _, _) = ops.wan_norm_rope_fwd(
        [{'src': qkv[:, 0:C], 'weight': wq, 'rope': True}, {'src': qkv[:, C:2 * C], 'weight': wk, 'rope': False},
         {'src': qkv[:, 2 * C:]}], B, L, H, cos, sin)

    def ref(x, w, rope):
        x = x.float().clone().requires_grad_(True)
        w = w.float().clone().requires_grad_(True)
        xh = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6)
        y = (xh * w).view(B, L, H, 128)
        if rope:
            y = W.apply_rope(y, cos.cpu().cuda(), sin.cpu().cuda())
        return x, w, y.permute(0, 2, 1, 3)
    xq, wqr, yq = ref(qkv[:, 0:C], wq, True)
    xk, wkr, yk = ref(qkv[:, C:2 * C], wk, False)
    for got, want in ((q, yq), (k, yk), (v, qkv[:, 2 * C:].float().view(B, L, H, 128).permute(0, 2, 1, 3))):
        err = (got.float() - want).abs().max().item()
        assert err <= 0.03 * want.abs().max().item(), err
    gq, gk, gv = (torch.randn(B, H, L, 128, device='cuda').bfloat16() for _ in range(3))
    dqkv = torch.empty_like(qkv)
    dwq, dwk, _ = ops.wan_norm_rope_bwd(
        [{'dy': gq, 'dx': dqkv[:, 0:C], 'weight': wq, 'xhat': xhq, 'rstd': rq, 'rope': True},
         {'dy': gk, 'dx': dqkv[:, C:2 * C], 'weight': wk, 'xhat': xhk, 'rstd': rk},
         {'dy': gv, 'dx': dqkv[:, 2 * C:]}], B, L, H, cos, sin)
    (yq * gq.float()).sum().backward()
    (yk * gk.float()).sum().backward()
    for got, want in ((dqkv[:, 0:C], xq.grad), (dqkv[:, C:2 * C], xk.grad), (dwq, wqr.grad), (dwk, wkr.grad)):
        rel = (got.float() - want).norm() / want.norm()
        assert rel <= 2e-2, rel.item()
    assert torch.equal(dqkv[:, 2 * C:].view(B, L, H, 128).permute(0, 2, 1, 3), gv)


def _make():
    from synth import fill_parameters
    from diffusion_pipe_b200.wan import WanPipeline
    from oracle import wan_ref as W
    model = WanPipeline({'model': {'dtype': 'bfloat16', 'transformer_config': CFG}})
    ref = fill_parameters(W.RefWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    return model, ref


def _batch(bs, seed):
    from oracle import wan_ref as W
    g = torch.Generator().manual_seed(seed)
    latents = torch.randn(bs, 16, 3, 8, 12, generator=g)
    text = torch.randn(bs, 16, 64, generator=g).bfloat16().float()
    lens = torch.tensor([10, 16][:bs])
    t = torch.sigmoid(torch.randn(bs, generator=g))
    noise = torch.randn(bs, 16, 3, 8, 12, generator=g)
    feats, (target, mask) = W.prepare_inputs(latents, text, lens, t, noise)
    return feats, (target, torch.tensor([]))


def test_layers_and_loss_match_oracle():
    from oracle import flux_ref as R
    from oracle import wan_ref as W
    model, ref = _make()
    feats, label = _batch(2, 1)
    x = tuple(f.cuda() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    loss = model.get_loss_fn()(x, tuple(l.cuda() for l in label))
    loss.backward()
    for emu, tol in ((True, 1e-3), (False, 5e-3)):
        ref.set_emulate_bf16(emu)
        ref.zero_grad()
        y = tuple(f.clone() for f in feats)
        for layer in W.to_layers(ref):
            y = layer(y)
        rloss = R.loss_fn(y, label)
        rel = abs(loss.item() - rloss.item()) / abs(rloss.item())
        assert rel <= tol, (emu, loss.item(), rloss.item(), rel)
        if emu:
            rloss.backward()
            rg = {n: p.grad for n, p in ref.named_parameters()}
            errs = {}
            for n, p in model.transformer.named_parameters():
                assert p.grad is not None, n
                errs[n] = ((p.grad.float().cpu() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
            bad = sorted(((v, k) for k, v in errs.items() if v > 5e-2), reverse=True)
            assert not bad, bad[:8]


def test_forward_matches_the_references_own_model(golden_dir):
    """the stored output of /root/reference/models/wan/model.py (fp32) on the fixture's inputs"""
    g = torch.load(os.path.join(golden_dir, 'wan_golden.pt'), weights_only=False)
    model, _ = _make()
    none = torch.tensor([], device='cuda')
    out = (g['x'].cuda(), none, g['t'].cuda(), g['text'].cuda(), g['text_lens'].cuda(), none)
    with torch.no_grad():
        for layer in model.to_layers():
            out = layer(out)
    err = (out.float().cpu() - g['out']).norm() / g['out'].norm()
    assert err <= 2e-2, err.item()


def test_wan_lora_matches_oracle():
    """LoRA on the ten Linear layers of every Wan block (K-extended GEMM operands) with the real kernels vs the oracle
    with PEFT-style adapters; host logic of the same function is covered on CPU by tests/test_lora_wan_host_logic.py"""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import test_lora_wan_host_logic as H
    model, ref = H.make_pair(device='cuda')
    loss, rloss = H.run_both(model, ref, *H.make_batch(), dev='cuda')
    H.check(model, ref, loss, rloss)
