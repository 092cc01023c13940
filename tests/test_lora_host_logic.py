"""CPU: LoRA on the fused Flux / Qwen-Image blocks (diffusion-pipe_b200/lora.py: K-extended GEMM operands, frozen base,
factor gradients) with the kernel wrappers replaced by the PyTorch test doubles of tests/kernel_doubles.py, against the
oracle with PEFT-style adapters (oracle/lora_ref.py).  The real kernels are exercised by tests/test_lora_gpu.py."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

RANK = 16


@pytest.fixture
def doubles(monkeypatch):
    import kernel_doubles
    from diffusion_pipe_b200 import ops
    kernel_doubles.install(monkeypatch, ops)
    return ops


def _sync_factors(model_t, ref, seed=0):
    """gives both sides the same non-trivial factors (B != 0 so that every gradient path is exercised)"""
    g = torch.Generator().manual_seed(seed)
    rp = dict(ref.named_parameters())
    n = 0
    with torch.no_grad():
        for name, p in model_t.named_parameters():
            if '.lora_A.' in name or '.lora_B.' in name:
                v = (0.05 * torch.randn(p.shape, generator=g)).bfloat16()
                p.copy_(v)
                rp[name].copy_(v.float())
                n += 1
    return n


def _flux_pair():
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    from oracle import lora_ref
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': cfg}}, device='cpu')
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=1, num_single=1, joint_dim=64, pooled_dim=32)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    ref.load_state_dict({k: v.detach().float() for k, v in model.transformer.state_dict().items()})
    model.configure_adapter({'type': 'lora', 'rank': RANK, 'alpha': RANK, 'dropout': 0.0, 'dtype': torch.bfloat16})
    lora_ref.add_lora(ref, RANK)
    ref.set_emulate_bf16(True)
    return model, ref


def _flux_batch(seed):
    from oracle import flux_ref as R
    g = torch.Generator().manual_seed(seed)
    bs = 2
    latents, noise = torch.randn(bs, 16, 8, 8, generator=g), torch.randn(bs, 16, 8, 8, generator=g)
    t5 = torch.randn(bs, 12, 64, generator=g).bfloat16()
    clip = torch.randn(bs, 32, generator=g).bfloat16()
    t = torch.sigmoid(torch.randn(bs, generator=g))
    feats, (target, _) = R.prepare_inputs(latents, t5, clip, t, noise)
    return feats, (target, torch.tensor([]))


def _run(layers, loss_fn, feats, label):
    x = tuple(f.clone() for f in feats)
    for layer in layers:
        x = layer(x)
    loss = loss_fn(x, label)
    loss.backward()
    return loss


def test_adapter_structure_matches_peft_layout(doubles):
    model, ref = _flux_pair()
    mine = {n: tuple(p.shape) for n, p in model.transformer.named_parameters()}
    theirs = {n: tuple(p.shape) for n, p in ref.named_parameters()}
    assert mine == theirs
    trainable = {n for n, p in model.transformer.named_parameters() if p.requires_grad}
    assert trainable == {n for n, p in ref.named_parameters() if p.requires_grad}
    assert trainable and all('.lora_A.' in n or '.lora_B.' in n for n in trainable)
    assert all(p.original_name == n for n, p in model.transformer.named_parameters())
    # every Linear of the blocks carries an adapter: 14 in a double block, 6 in a single block (incl. the AdaLN linears)
    assert sum(".lora_A." in n for n in mine) == 14 + 6
    # default init: B = 0, so the adapted model starts as the base model
    blk = model.transformer.transformer_blocks[0]
    assert float(blk.attn.to_q.lora_B.weight.detach().abs().max()) == 0.0 and float(blk.attn.to_q.lora_A.weight.detach().abs().max()) > 0
    # the base weight lives once, inside the site buffer
    site = blk.lora['qkv']
    assert blk.attn.to_k.weight.data_ptr() == site.buf[256:512].data_ptr()
    with pytest.raises(ValueError):
        from diffusion_pipe_b200.lora import LoraSite
        LoraSite([blk.attn.to_out[0]], 12)


def test_flux_lora_forward_backward_matches_oracle(doubles):
    from oracle import flux_ref as R
    model, ref = _flux_pair()
    assert _sync_factors(model.transformer, ref) == 40
    feats, label = _flux_batch(1)
    loss = _run(model.to_layers(), model.get_loss_fn(), feats, label)
    rloss = _run(R.to_layers(ref), R.loss_fn, feats, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    errs = {}
    for n, p in model.transformer.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, n                                  # the frozen base gets no gradient buffers
            continue
        assert p.grad is not None, n
        errs[n] = ((p.grad.float() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
    bad = sorted(((v, k) for k, v in errs.items() if v > 6e-2), reverse=True)
    assert not bad, bad[:8]


def test_factor_updates_reach_the_site_buffers_and_gradients_accumulate(doubles):
    """an optimizer step (in-place update of A / B) must be seen by the next forward; two micro-batches accumulate"""
    from oracle import flux_ref as R
    model, ref = _flux_pair()
    _sync_factors(model.transformer, ref, seed=3)
    b1, b2 = _flux_batch(5), _flux_batch(6)
    layers = model.to_layers()
    _run(layers, model.get_loss_fn(), *b1)
    _run(layers, model.get_loss_fn(), *b2)
    _run(R.to_layers(ref), R.loss_fn, *b1)
    _run(R.to_layers(ref), R.loss_fn, *b2)
    rg = {n: p.grad for n, p in ref.named_parameters()}
    for n, p in model.transformer.named_parameters():
        if p.requires_grad:
            rel = ((p.grad.float() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
            assert rel <= 6e-2, (n, rel)
    params = [p for p in model.transformer.parameters() if p.requires_grad]
    rparams = [p for p in ref.parameters() if p.requires_grad]
    opt, ropt = torch.optim.SGD(params, lr=0.5), torch.optim.SGD(rparams, lr=0.5)
    l0 = _run(layers, model.get_loss_fn(), *b1).item()
    opt.step(); opt.zero_grad(set_to_none=True)
    ref.zero_grad(); _run(R.to_layers(ref), R.loss_fn, *b1)
    # give the oracle the product's updated factors (bf16), then both must agree on the new loss
    rp = dict(ref.named_parameters())
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.requires_grad:
                rp[n].copy_(p.float())
    l1 = _run(layers, model.get_loss_fn(), *b1).item()
    r1 = _run(R.to_layers(ref), R.loss_fn, *b1).item()
    assert l1 != l0
    assert abs(l1 - r1) / abs(r1) <= 1e-3, (l1, r1)


def test_qwen_lora_forward_backward_matches_oracle(doubles):
    from synth import fill_parameters
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    from oracle import flux_ref as R
    from oracle import lora_ref
    from oracle import qwen_ref as Q
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'transformer_config': cfg}}, device='cpu')
    ref = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    model.configure_adapter({'type': 'lora', 'rank': RANK, 'alpha': RANK, 'dropout': 0.0})
    lora_ref.add_lora(ref, RANK)
    ref.set_emulate_bf16(True)
    assert _sync_factors(model.transformer, ref, seed=2) == 56
    g = torch.Generator().manual_seed(2)
    latents, noise = torch.randn(2, 16, 1, 8, 12, generator=g), torch.randn(2, 16, 1, 8, 12, generator=g)
    pe = [torch.randn(11, 64, generator=g).bfloat16().float() for _ in range(2)]
    t = torch.sigmoid(torch.randn(2, generator=g))
    feats, (target, _) = Q.prepare_inputs(latents, pe, t, noise)
    label = (target, torch.tensor([]))
    loss = _run(model.to_layers(), model.get_loss_fn(), feats, label)
    rloss = _run(Q.to_layers(ref), R.loss_fn, feats, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    for n, p in model.transformer.named_parameters():
        if p.requires_grad and rg[n] is not None:
            rel = ((p.grad.float() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
            assert rel <= 6e-2, (n, rel)


def test_qwen_lora_with_ragged_prompts_matches_oracle(doubles):
    """adapters + the key mask of a padded micro-batch (prompts of 4 and 11 tokens)"""
    from synth import fill_parameters
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    from oracle import flux_ref as R
    from oracle import lora_ref
    from oracle import qwen_ref as Q
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg}})
    ref = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    model.configure_adapter({'type': 'lora', 'rank': RANK, 'alpha': RANK, 'dropout': 0.0})
    lora_ref.add_lora(ref, RANK)
    ref.set_emulate_bf16(True)
    _sync_factors(model.transformer, ref, seed=6)
    g = torch.Generator().manual_seed(6)
    latents, noise = torch.randn(2, 16, 1, 8, 12, generator=g), torch.randn(2, 16, 1, 8, 12, generator=g)
    pe = [torch.randn(4, 64, generator=g).bfloat16().float(), torch.randn(11, 64, generator=g).bfloat16().float()]
    feats, (target, _) = Q.prepare_inputs(latents, pe, torch.sigmoid(torch.randn(2, generator=g)), noise)
    assert not bool(feats[2].all())
    label = (target, torch.tensor([]))
    loss = _run(model.to_layers(), model.get_loss_fn(), feats, label)
    rloss = _run(Q.to_layers(ref), R.loss_fn, feats, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    for n, p in model.transformer.named_parameters():
        if p.requires_grad and rg[n] is not None:
            rel = ((p.grad.float() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
            assert rel <= 6e-2, (n, rel)
