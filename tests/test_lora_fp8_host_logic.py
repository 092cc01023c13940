"""CPU: `transformer_dtype = 'float8'` under LoRA — the frozen 2-D weights of the blocks are STORED
in fp8 and widened into the shared operand scratch in front of every GEMM (lora.py: LoraSite._operand / ModLora._base,
csrc/fp8_dequant.cu).  Kernel wrappers are the PyTorch test doubles; the oracle is the PEFT-style restatement with the
same weights rounded through fp8 by the reference's own per-family selection rule (oracle/lora_ref.py: FP8_RULES, citing
models/flux.py:203-205, models/qwen_image.py:261-263, models/wan/wan.py:233-235).  The code table of the real kernel is
checked against torch in tests/test_abi.py; the kernel itself in tests/test_zz_first_hardware_run_gpu.py."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

RANK = 16
FP8 = {'float8': torch.float8_e4m3fn, 'float8_e5m2': torch.float8_e5m2}


@pytest.fixture
def doubles(monkeypatch):
    import kernel_doubles
    from diffusion_pipe_b200 import ops
    kernel_doubles.install(monkeypatch, ops)
    return ops


def _sync_factors(model_t, ref, seed=0):
    g = torch.Generator().manual_seed(seed)
    rp = dict(ref.named_parameters())
    with torch.no_grad():
        for name, p in model_t.named_parameters():
            if '.lora_A.' in name or '.lora_B.' in name:
                v = (0.05 * torch.randn(p.shape, generator=g)).bfloat16()
                p.copy_(v)
                rp[name].copy_(v.float())


def _run(layers, loss_fn, feats, label, dev=None):
    x = tuple(f.clone().to(dev) if dev else f.clone() for f in feats)
    for layer in layers:
        x = layer(x)
    loss = loss_fn(x, tuple(l.to(dev) for l in label) if dev else label)
    loss.backward()
    return loss


def _check_grads(model_t, ref, tol=6e-2):
    rg = {n: p.grad for n, p in ref.named_parameters()}
    for n, p in model_t.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, n
        elif rg[n] is not None:
            rel = ((p.grad.float().cpu() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
            assert rel <= tol, (n, rel)


def flux_pair(transformer_dtype, device='cpu'):
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    from oracle import lora_ref
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    mc = {'dtype': 'bfloat16', 'guidance': 1.0, 'device': device, 'transformer_config': cfg}
    if transformer_dtype:
        mc['transformer_dtype'] = transformer_dtype
    model = FluxPipeline({'model': mc})
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=1, num_single=1, joint_dim=64, pooled_dim=32)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
            if p.ndim == 2:
                p.mul_(2.0)            # std 0.04: fp8 normals and subnormals both occur
    ref.load_state_dict({k: v.detach().float().cpu() for k, v in model.transformer.state_dict().items()})
    model.configure_adapter({'type': 'lora', 'rank': RANK, 'alpha': RANK, 'dropout': 0.0, 'dtype': torch.bfloat16})
    lora_ref.add_lora(ref, RANK)
    ref.set_emulate_bf16(True)
    _sync_factors(model.transformer, ref, seed=11)
    return model, ref


def flux_batch(seed):
    from oracle import flux_ref as R
    g = torch.Generator().manual_seed(seed)
    bs = 2
    latents, noise = torch.randn(bs, 16, 8, 8, generator=g), torch.randn(bs, 16, 8, 8, generator=g)
    t5 = torch.randn(bs, 12, 64, generator=g).bfloat16()
    clip = torch.randn(bs, 32, generator=g).bfloat16()
    t = torch.sigmoid(torch.randn(bs, generator=g))
    feats, (target, _) = R.prepare_inputs(latents, t5, clip, t, noise)
    return feats, (target, torch.tensor([]))


@pytest.mark.parametrize('name', ['float8', 'float8_e5m2'])
def test_flux_fp8_base_matches_the_oracle_with_fp8_rounded_weights(doubles, name):
    from oracle import flux_ref as R
    from oracle import lora_ref
    model, ref = flux_pair(name)
    stored = {n for n, p in model.transformer.named_parameters() if p.dtype == FP8[name]}
    # exactly the parameters the reference's rule selects are held in fp8; nothing else changed dtype
    assert stored == lora_ref.fp8_stored_names(ref, 'flux') and len(stored) == 14 + 6
    assert all(p.dtype == torch.bfloat16 for n, p in model.transformer.named_parameters() if n not in stored)
    blk = model.transformer.transformer_blocks[0]
    site = blk.lora['qkv']
    assert site.buf is None and site.w8.dtype == FP8[name]                      # W exists once, in fp8
    assert blk.attn.to_k.weight.data_ptr() == site.w8[256:512].data_ptr()
    feats, label = flux_batch(1)
    loss = _run(model.to_layers(), model.get_loss_fn(), feats, label)
    unrounded = _run(R.to_layers(ref), R.loss_fn, feats, label).item()
    ref.zero_grad()
    assert lora_ref.round_base_through_fp8(ref, 'flux', FP8[name]) == stored
    rloss = _run(R.to_layers(ref), R.loss_fn, feats, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    assert unrounded != rloss.item()                                             # the rounding reaches the loss
    _check_grads(model.transformer, ref)
    # the GEMM operands are exactly [[widen(fp8(W)) | B], [A | .]]
    rp = dict(ref.named_parameters())
    for b in list(model.transformer.transformer_blocks) + list(model.transformer.single_transformer_blocks):
        for site in b.lora['sites']:
            op = site._operand()
            want = torch.cat([rp[l.weight.original_name] for l in site.lins])
            assert op.dtype == torch.bfloat16 and torch.equal(op[:site.N, :site.K].float(), want)
            assert torch.equal(site.w_dgrad[site.N:], site.a_all) and torch.equal(site.w_fwd[:, site.K:], site.b_blk)
            assert torch.equal(site.a_all[:site.r], site.A[0].detach()) and torch.equal(site.b_blk[:site.sizes[0], :site.r], site.B[0].detach())


def test_fp8_operand_scratch_is_shared_and_follows_factor_updates(doubles):
    """one scratch buffer per device serves every site; an optimizer step on the factors is seen by the next forward"""
    from diffusion_pipe_b200 import lora
    from oracle import flux_ref as R
    from oracle import lora_ref
    model, ref = flux_pair('float8')
    lora_ref.round_base_through_fp8(ref, 'flux')
    layers = model.to_layers()
    b1 = flux_batch(5)
    l0 = _run(layers, model.get_loss_fn(), *b1).item()
    assert len(lora._SCRATCH) == 1
    largest = max(max((s.N + s.R) * (s.K + s.R) for s in blk.lora['sites'])
                  for blk in list(model.transformer.transformer_blocks) + list(model.transformer.single_transformer_blocks))
    assert next(iter(lora._SCRATCH.values())).numel() >= largest
    params = [p for p in model.transformer.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.5)
    opt.step()
    opt.zero_grad(set_to_none=True)
    rp = dict(ref.named_parameters())
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.requires_grad:
                rp[n].copy_(p.float())
    l1 = _run(layers, model.get_loss_fn(), *b1).item()
    ref.zero_grad()
    r1 = _run(R.to_layers(ref), R.loss_fn, *b1).item()
    assert l1 != l0 and abs(l1 - r1) / abs(r1) <= 1e-3, (l0, l1, r1)
    _check_grads(model.transformer, ref)


def test_qwen_fp8_base_matches_oracle(doubles):
    from synth import fill_parameters
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    from oracle import flux_ref as R
    from oracle import lora_ref
    from oracle import qwen_ref as Q
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_dtype': torch.float8_e4m3fn,
                                         'transformer_config': cfg}})
    ref = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    model.configure_adapter({'type': 'lora', 'rank': RANK, 'alpha': RANK, 'dropout': 0.0})
    lora_ref.add_lora(ref, RANK)
    ref.set_emulate_bf16(True)
    stored = {n for n, p in model.transformer.named_parameters() if p.dtype == torch.float8_e4m3fn}
    assert stored == lora_ref.round_base_through_fp8(ref, 'qwen_image') and len(stored) == 2 * 14
    _sync_factors(model.transformer, ref, seed=2)
    g = torch.Generator().manual_seed(2)
    latents, noise = torch.randn(2, 16, 1, 8, 12, generator=g), torch.randn(2, 16, 1, 8, 12, generator=g)
    pe = [torch.randn(4, 64, generator=g).bfloat16().float(), torch.randn(11, 64, generator=g).bfloat16().float()]
    feats, (target, _) = Q.prepare_inputs(latents, pe, torch.sigmoid(torch.randn(2, generator=g)), noise)
    label = (target, torch.tensor([]))
    loss = _run(model.to_layers(), model.get_loss_fn(), feats, label)
    rloss = _run(Q.to_layers(ref), R.loss_fn, feats, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    _check_grads(model.transformer, ref)


def test_wan_fp8_base_matches_oracle(doubles):
    import test_lora_wan_host_logic as H
    from oracle import lora_ref
    model, ref = H.make_pair(transformer_dtype='float8')
    stored = {n for n, p in model.transformer.named_parameters() if p.dtype == torch.float8_e4m3fn}
    assert stored == lora_ref.round_base_through_fp8(ref, 'wan') and len(stored) == 2 * 10
    loss, rloss = H.run_both(model, ref, *H.make_batch())
    H.check(model, ref, loss, rloss)


def test_fp8_base_without_an_adapter_is_refused():
    from diffusion_pipe_b200.flux import FluxPipeline
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_dtype': 'float8', 'lazy_layers': True,
                                    'transformer_config': cfg}})
    with pytest.raises(NotImplementedError):
        model.to_layers()
    with pytest.raises(NotImplementedError):
        FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_dtype': 'float16', 'lazy_layers': True,
                                'transformer_config': cfg}}).to_layers()
