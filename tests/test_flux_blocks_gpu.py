"""GPU: the fused Flux blocks (diffusion-pipe_b200/flux_blocks.py, C-ABI kernels) against the oracle
(oracle/flux_ref.py, pinned to the reference's flow blocks by tests/test_oracle_golden.py) on identical inputs and
weights.  Checked: both outputs, gradients w.r.t. both streams and the timestep embedding, and every parameter
gradient.  Tolerance: relative Frobenius error <= 2e-2 against the fp32 oracle (bf16 compute, fp32 accumulation)
and <= 1.5e-2 against the oracle with bf16 rounding points emulated."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _copy_weights(src_module, dst_module):
    sd = {k: v.detach().float().cpu() for k, v in src_module.state_dict().items()}
    missing, unexpected = dst_module.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)


def _grads_by_name(module):
    return {n: p.grad.detach().float().cpu() for n, p in module.named_parameters() if p.grad is not None}


@pytest.mark.parametrize('kind', ['double', 'single'])
@pytest.mark.parametrize('batch', [1, 2])
def test_block_matches_oracle(kind, batch, golden_dir):
    from diffusion_pipe_b200 import flux_blocks as FB
    from oracle import flux_ref as R
    torch.manual_seed(11)
    g = torch.load(os.path.join(golden_dir, 'flux_blocks_golden.pt'), weights_only=False)
    D, H, Lt, Li = g['dim'], g['heads'], 40, 88
    if kind == 'double':
        blk = FB.FluxTransformerBlock(D, H, 2)
        ref = R.RefFluxTransformerBlock(D, H, 2)
    else:
        blk = FB.FluxSingleTransformerBlock(D, H, 2)
        ref = R.RefFluxSingleTransformerBlock(D, H, 2)
    # golden weights where the fixture has them (exactly bf16-representable), random elsewhere (modulation linears)
    with torch.no_grad():
        sd = blk.state_dict()
        for k, v in g[kind]['weights'].items():
            sd[k].copy_(v)
        for n, p in blk.named_parameters():
            if 'norm' in n and n.endswith('linear.weight'):
                p.normal_(0, 0.05)
            if 'norm' in n and n.endswith('linear.bias'):
                p.normal_(0, 0.2)
    _copy_weights(blk, ref)
    ids = torch.zeros(Lt + Li, 3)
    ids[Lt:, 1] = torch.arange(Li) // 8
    ids[Lt:, 2] = torch.arange(Li) % 8
    cos, sin = R.flux_rope_tables(ids)
    hid = torch.randn(batch, Li, D).bfloat16()
    enc = torch.randn(batch, Lt, D).bfloat16()
    temb = torch.randn(batch, D).bfloat16()
    g_h = torch.randn(batch, Li, D).bfloat16()
    g_e = torch.randn(batch, Lt, D).bfloat16()

    def run(block, dev, dtype):
        h = hid.to(dev, dtype).requires_grad_(True)
        e = enc.to(dev, dtype).requires_grad_(True)
        t = temb.to(dev, dtype).requires_grad_(True)
        eo, ho = block(h, e, t, (cos.to(dev), sin.to(dev)))
        ((ho.float() * g_h.to(dev).float()).sum() + (eo.float() * g_e.to(dev).float()).sum()).backward()
        return ho.detach(), eo.detach(), h.grad, e.grad, t.grad

    got = run(blk, 'cuda', torch.bfloat16)
    results = {}
    for emu, tol in ((False, 2e-2), (True, 1.5e-2)):
        ref.zero_grad()
        for m in ref.modules():
            if hasattr(m, 'emulate_bf16'):
                m.emulate_bf16 = emu
        want = run(ref, 'cpu', torch.float32)
        names = ('hidden', 'encoder', 'd_hidden', 'd_encoder', 'd_temb')
        for n, a, b in zip(names, got, want):
            results[(emu, n)] = _rel(a, b)
            assert results[(emu, n)] <= tol, (emu, n, results)
        if not emu:
            pg, rg = _grads_by_name(blk), _grads_by_name(ref)
            assert set(pg) == set(rg), set(pg) ^ set(rg)
            for n in rg:
                r = _rel(pg[n], rg[n])
                assert r <= 3e-2, (n, r)


def test_gradients_accumulate_across_micro_batches():
    from diffusion_pipe_b200 import flux_blocks as FB
    torch.manual_seed(12)
    blk = FB.FluxSingleTransformerBlock(256, 2, 2)
    cos = torch.ones(64, 128, device='cuda')
    sin = torch.zeros(64, 128, device='cuda')

    def step():
        h = torch.randn(1, 40, 256, device='cuda').bfloat16().requires_grad_(True)
        e = torch.randn(1, 24, 256, device='cuda').bfloat16().requires_grad_(True)
        t = torch.randn(1, 256, device='cuda').bfloat16().requires_grad_(True)
        eo, ho = blk(h, e, t, (cos, sin))
        (ho.float().sum() + eo.float().sum()).backward()

    torch.manual_seed(1)
    step()
    g1 = {n: p.grad.clone().float() for n, p in blk.named_parameters()}
    blk.zero_grad(set_to_none=True)
    torch.manual_seed(1)
    step()
    torch.manual_seed(1)
    step()
    for n, p in blk.named_parameters():
        assert _rel(p.grad, 2 * g1[n]) < 1e-2, n
    # the per-projection parameters alias one fused gradient buffer
    assert blk.attn.to_q.weight.grad.data_ptr() == blk.lin1.wgrad.data_ptr()
