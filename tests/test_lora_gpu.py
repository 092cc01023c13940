"""GPU: LoRA on the fused Flux / Qwen-Image blocks with the real kernels (K-extended GEMM operands: strided [W | B] /
[W ; A] views, adapter columns appended to the activation buffers) against the oracle with PEFT-style adapters
(oracle/lora_ref.py).  Tolerance: loss 1e-3 relative (bf16 rounding points emulated), factor gradients 6e-2 relative L2."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
RANK = 16


def _sync_factors(model_t, ref, seed=0):
    g = torch.Generator().manual_seed(seed)
    rp = dict(ref.named_parameters())
    with torch.no_grad():
        for name, p in model_t.named_parameters():
            if '.lora_A.' in name or '.lora_B.' in name:
                v = (0.05 * torch.randn(p.shape, generator=g)).bfloat16()
                p.copy_(v)
                rp[name].copy_(v.float())


def _check(model_t, ref, loss, rloss):
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    errs = {}
    for n, p in model_t.named_parameters():
        if not p.requires_grad or rg[n] is None:
            continue
        assert p.grad is not None, n
        errs[n] = ((p.grad.float().cpu() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
    bad = sorted(((v, k) for k, v in errs.items() if v > 6e-2), reverse=True)
    assert errs and not bad, bad[:8]


def _run(layers, loss_fn, feats, label, dev=None):
    x = tuple(f.clone().to(dev) if dev else f.clone() for f in feats)
    for layer in layers:
        x = layer(x)
    loss = loss_fn(x, tuple(l.to(dev) if dev else l for l in label))
    loss.backward()
    return loss


def test_flux_lora_matches_oracle():
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    from oracle import lora_ref
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': cfg}})
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=1, num_single=1, joint_dim=64, pooled_dim=32)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    ref.load_state_dict({k: v.detach().float().cpu() for k, v in model.transformer.state_dict().items()})
    model.configure_adapter({'type': 'lora', 'rank': RANK, 'alpha': RANK, 'dropout': 0.0})
    lora_ref.add_lora(ref, RANK)
    ref.set_emulate_bf16(True)
    _sync_factors(model.transformer, ref)
    g = torch.Generator().manual_seed(1)
    bs = 2
    latents, noise = torch.randn(bs, 16, 16, 16, generator=g), torch.randn(bs, 16, 16, 16, generator=g)
    t5 = torch.randn(bs, 32, 64, generator=g).bfloat16()
    clip = torch.randn(bs, 32, generator=g).bfloat16()
    t = torch.sigmoid(torch.randn(bs, generator=g))
    feats, (target, _) = R.prepare_inputs(latents, t5, clip, t, noise)
    label = (target, torch.tensor([]))
    layers = model.to_layers()
    loss = _run(layers, model.get_loss_fn(), feats, label, 'cuda')
    rloss = _run(R.to_layers(ref), R.loss_fn, feats, label)
    _check(model.transformer, ref, loss, rloss)
    assert all(p.grad is None for p in model.transformer.parameters() if not p.requires_grad)
    # an optimizer step on the factors is picked up by the next forward
    params = [p for p in model.transformer.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.5)
    opt.step()
    opt.zero_grad(set_to_none=True)
    rp = dict(ref.named_parameters())
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.requires_grad:
                rp[n].copy_(p.float().cpu())
    ref.zero_grad()
    loss2 = _run(layers, model.get_loss_fn(), feats, label, 'cuda')
    rloss2 = _run(R.to_layers(ref), R.loss_fn, feats, label)
    assert loss2.item() != loss.item()
    _check(model.transformer, ref, loss2, rloss2)


def test_qwen_lora_matches_oracle():
    from synth import fill_parameters
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    from oracle import flux_ref as R
    from oracle import lora_ref
    from oracle import qwen_ref as Q
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'transformer_config': cfg}})
    ref = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    model.configure_adapter({'type': 'lora', 'rank': RANK, 'alpha': RANK, 'dropout': 0.0})
    lora_ref.add_lora(ref, RANK)
    ref.set_emulate_bf16(True)
    _sync_factors(model.transformer, ref, seed=2)
    g = torch.Generator().manual_seed(2)
    latents, noise = torch.randn(2, 16, 1, 16, 24, generator=g), torch.randn(2, 16, 1, 16, 24, generator=g)
    pe = [torch.randn(12, 64, generator=g).bfloat16().float() for _ in range(2)]
    t = torch.sigmoid(torch.randn(2, generator=g))
    feats, (target, _) = Q.prepare_inputs(latents, pe, t, noise)
    label = (target, torch.tensor([]))
    loss = _run(model.to_layers(), model.get_loss_fn(), feats, label, 'cuda')
    rloss = _run(Q.to_layers(ref), R.loss_fn, feats, label)
    _check(model.transformer, ref, loss, rloss)
