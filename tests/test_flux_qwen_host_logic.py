"""CPU: the HOST logic of the Flux and Qwen-Image models (flux.py, flux_blocks.py, qwen_image.py: buffer plumbing,
gradient routing into fused parameter buffers, what is saved for backward) with the kernel wrappers replaced by the
PyTorch test doubles of tests/kernel_doubles.py, against the oracles.  The real kernels are checked by the GPU suite."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def doubles(monkeypatch):
    import kernel_doubles
    from diffusion_pipe_b200 import ops
    kernel_doubles.install(monkeypatch, ops)
    return ops


def _compare(model_params, ref, loss, rloss, skip_none=False):
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    errs = {}
    for n, p in model_params:
        if rg[n] is None and skip_none:
            continue
        assert p.grad is not None, n
        errs[n] = ((p.grad.float() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
    bad = sorted(((v, k) for k, v in errs.items() if v > 5e-2), reverse=True)
    assert not bad, bad[:8]


def test_flux_model_forward_backward_matches_oracle(doubles):
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': cfg}}, device='cpu')
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=1, num_single=1, joint_dim=64, pooled_dim=32)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    ref.load_state_dict({k: v.detach().float() for k, v in model.transformer.state_dict().items()})
    ref.set_emulate_bf16(True)
    g = torch.Generator().manual_seed(1)
    bs = 2
    latents, noise = torch.randn(bs, 16, 8, 8, generator=g), torch.randn(bs, 16, 8, 8, generator=g)
    t5 = torch.randn(bs, 12, 64, generator=g).bfloat16()
    clip = torch.randn(bs, 32, generator=g).bfloat16()
    t = torch.sigmoid(torch.randn(bs, generator=g))
    feats, (target, _) = R.prepare_inputs(latents, t5, clip, t, noise)
    label = (target, torch.tensor([]))
    x = tuple(f.clone() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    loss = model.get_loss_fn()(x, label)
    loss.backward()
    y = tuple(f.clone() for f in feats)
    for layer in R.to_layers(ref):
        y = layer(y)
    rloss = R.loss_fn(y, label)
    rloss.backward()
    _compare(model.transformer.named_parameters(), ref, loss, rloss)


@pytest.mark.parametrize('control', [False, True])
def test_qwen_model_forward_backward_matches_oracle(doubles, control):
    from synth import fill_parameters
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    from oracle import flux_ref as R
    from oracle import qwen_ref as Q
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'transformer_config': cfg}}, device='cpu')
    ref = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    ref.set_emulate_bf16(True)
    g = torch.Generator().manual_seed(2)
    bs = 2
    latents, noise = torch.randn(bs, 16, 1, 8, 12, generator=g), torch.randn(bs, 16, 1, 8, 12, generator=g)
    pe = [torch.randn(11, 64, generator=g).bfloat16().float() for _ in range(bs)]
    t = torch.sigmoid(torch.randn(bs, generator=g))
    ctrl = torch.randn(bs, 16, 1, 8, 12, generator=g) if control else None
    feats, (target, _) = Q.prepare_inputs(latents, pe, t, noise, control_latents=ctrl)
    label = (target, torch.tensor([]))
    x = tuple(f.clone() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    loss = model.get_loss_fn()(x, label)
    loss.backward()
    y = tuple(f.clone() for f in feats)
    for layer in Q.to_layers(ref):
        y = layer(y)
    rloss = R.loss_fn(y, label)
    rloss.backward()
    _compare(model.transformer.named_parameters(), ref, loss, rloss, skip_none=True)
    # ragged prompts inside one micro-batch: the bool key mask of models/qwen_image.py:472-476 — every sample attends
    # over its own real prompt tokens + all image tokens; loss and every parameter gradient match the masked oracle
    model.transformer.zero_grad(set_to_none=True)
    ref.zero_grad()
    feats2, (target2, _) = Q.prepare_inputs(latents, [pe[0][:5], pe[1]], t, noise, control_latents=ctrl)
    assert not bool(feats2[2].all())
    label2 = (target2, torch.tensor([]))
    x = tuple(f.clone() for f in feats2)
    layers = model.to_layers()
    x = layers[0](x)
    assert x[-1].dtype == torch.int32 and x[-1].tolist() == [5, 11]      # the prompt lengths ride at the end of the tuple
    for layer in layers[1:]:
        x = layer(x)
    loss2 = model.get_loss_fn()(x, label2)
    loss2.backward()
    y = tuple(f.clone() for f in feats2)
    for layer in Q.to_layers(ref):
        y = layer(y)
    rloss2 = R.loss_fn(y, label2)
    rloss2.backward()
    _compare(model.transformer.named_parameters(), ref, loss2, rloss2, skip_none=True)


def test_activation_checkpointing_of_fused_blocks_is_equivalent(doubles):
    """train.py:588-603 (`activation_checkpointing = true`): torch.utils.checkpoint around every block layer — the fused
    autograd functions are recomputed in backward and give the same loss and gradients, also in the zero-bubble order"""
    from functools import partial
    from diffusion_pipe_b200.flux import FluxPipeline
    from diffusion_pipe_b200.pipe import ManualPipelineModule, initialize
    from oracle import flux_ref as R
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    g = torch.Generator().manual_seed(1)
    bs = 1
    mbs = []
    for _ in range(2):
        latents, noise = torch.randn(bs, 16, 8, 8, generator=g), torch.randn(bs, 16, 8, 8, generator=g)
        t5 = torch.randn(bs, 12, 64, generator=g).bfloat16()
        clip = torch.randn(bs, 32, generator=g).bfloat16()
        t = torch.sigmoid(torch.randn(bs, generator=g))
        feats, (target, _) = R.prepare_inputs(latents, t5, clip, t, noise)
        mbs.append((feats, (target, torch.tensor([]))))
    results = []
    for ckpt, schedule in ((False, '1f1b'), (True, '1f1b'), (True, 'zb')):
        torch.manual_seed(0)
        model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'device': 'cpu', 'transformer_config': cfg}})
        extra = {}
        if ckpt:
            extra = {'activation_checkpoint_interval': 1, 'checkpointable_layers': model.checkpointable_layers,
                     'activation_checkpoint_func': partial(torch.utils.checkpoint.checkpoint, use_reentrant=False)}
        pm = ManualPipelineModule(layers=model.to_layers(), num_stages=1, partition_method='uniform', manual_partition_split=None,
                                  loss_fn=model.get_loss_fn(), dynamic_shape=True, device=torch.device('cpu'), **extra)
        engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': 1, 'gradient_accumulation_steps': 2,
                                                       'gradient_clipping': 0.0, 'steps_per_print': 0, 'pipeline_schedule': schedule})
        captured = {}

        class Capture(torch.optim.SGD):
            def step(self, closure=None):
                captured.update({p.original_name: p.grad.detach().float().clone() for gr in self.param_groups for p in gr['params']})
        params = [p for p in pm.parameters() if p.requires_grad]
        engine._configure_optimizer(lambda ps: Capture(ps, lr=0.0), params)
        loss = float(engine.train_batch(iter([(tuple(f.clone() for f in fe), la) for fe, la in mbs])))
        results.append((loss, captured))
    base_loss, base = results[0]
    for loss, grads in results[1:]:
        assert loss == pytest.approx(base_loss, rel=1e-6)
        assert grads.keys() == base.keys()
        for k in base:
            torch.testing.assert_close(grads[k], base[k], rtol=1e-5, atol=1e-6, msg=k)


def test_flux_kontext_control_latents_match_oracle(doubles):
    """Flux Kontext (models/flux.py:381-392): control latents are appended on the sequence axis with id[..., 0] = 1 and the
    prediction keeps only the first img_seq_len tokens.  The reference slices inside OutputWrapper with a host sync
    (:545-546); here the loss slices by the target's length — same loss, same gradients."""
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'device': 'cpu', 'transformer_config': cfg}})
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=1, num_single=1, joint_dim=64, pooled_dim=32)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    ref.load_state_dict({k: v.detach().float() for k, v in model.transformer.state_dict().items()})
    ref.set_emulate_bf16(True)
    g = torch.Generator().manual_seed(3)
    batch = {'latents': torch.randn(2, 16, 8, 8, generator=g), 'control_latents': torch.randn(2, 16, 8, 8, generator=g),
             't5_embed': torch.randn(2, 12, 64, generator=g).bfloat16(), 'clip_embed': torch.randn(2, 32, generator=g).bfloat16(), 'mask': None}
    torch.manual_seed(5)
    feats, (target, mask) = model.prepare_inputs(batch)
    x_t, _, _, _, img_ids, _, _, img_seq_len = feats
    assert x_t.shape == (2, 32, 64) and target.shape == (2, 16, 64) and img_seq_len.tolist() == [16, 16]
    assert img_ids[:, :16, 0].abs().max() == 0 and bool((img_ids[:, 16:, 0] == 1).all())
    label = (target, torch.tensor([]))
    x = tuple(f.clone() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    assert x.shape == (2, 32, 64)                      # every token is predicted; the loss keeps the first 16
    loss = model.get_loss_fn()(x, label)
    loss.backward()
    y = tuple(f.clone() for f in feats)
    for layer in R.to_layers(ref):
        y = layer(y)
    assert y.shape == (2, 16, 64)
    rloss = R.loss_fn(y, label)
    rloss.backward()
    _compare(model.transformer.named_parameters(), ref, loss, rloss)


@pytest.mark.parametrize('family', ['flux', 'qwen_image', 'wan'])
@pytest.mark.parametrize('variant', ['mask', 'huber', 'smooth_l1'])
def test_masked_and_robust_losses_match_the_reference_loss(doubles, family, variant):
    """models/base.py:418-436: per-pixel masks (train.py masks, resized nearest-exact to the latent grid and packed like the
    latents) and the `huber_delta` / `smooth_l1_beta` config keys, on every family's prediction layout"""
    from oracle import flux_ref as R
    g = torch.Generator().manual_seed(11)
    cfg_extra = {'huber': {'huber_delta': 0.7}, 'smooth_l1': {'smooth_l1_beta': 0.3}}.get(variant, {})
    pix = torch.rand(2, 64, 64, generator=g).round()                     # one mask per sample at pixel resolution
    if family == 'flux':
        from diffusion_pipe_b200.flux import FluxPipeline
        pipe = FluxPipeline(dict({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'device': 'cpu'}}, **cfg_extra))
        batch = {'latents': torch.randn(2, 16, 8, 8, generator=g), 't5_embed': torch.randn(2, 6, 64, generator=g).bfloat16(),
                 'clip_embed': torch.randn(2, 32, generator=g).bfloat16(), 'mask': pix}
    elif family == 'qwen_image':
        from diffusion_pipe_b200.qwen_image import QwenImagePipeline
        pipe = QwenImagePipeline(dict({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'device': 'cpu'}}, **cfg_extra))
        batch = {'latents': torch.randn(2, 16, 1, 8, 8, generator=g), 'prompt_embeds': [torch.randn(5, 64, generator=g) for _ in range(2)], 'mask': pix}
    else:
        from diffusion_pipe_b200.wan import WanPipeline
        pipe = WanPipeline(dict({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'device': 'cpu'}}, **cfg_extra))
        batch = {'latents': torch.randn(2, 16, 3, 8, 8, generator=g), 'text_embeddings': torch.randn(2, 16, 64, generator=g),
                 'seq_lens': torch.tensor([9, 16]), 'mask': pix}
    _, (target, mask) = pipe.prepare_inputs(batch)
    assert mask is not None and mask.shape[0] == 2 and torch.broadcast_shapes(mask.shape, target.shape) == target.shape
    assert set(mask.unique().tolist()) <= {0.0, 1.0} and 0 < float(mask.mean()) < 1
    if variant != 'mask':
        mask = torch.tensor([])
    pred = (target + 0.3 * torch.randn(target.shape, generator=g)).bfloat16().requires_grad_(True)
    loss = pipe.get_loss_fn()(pred, (target, mask))
    loss.backward()
    ref_pred = pred.detach().float().requires_grad_(True)
    rloss = R.loss_fn(ref_pred, (target, mask), huber_delta=cfg_extra.get('huber_delta'), smooth_l1_beta=cfg_extra.get('smooth_l1_beta'))
    rloss.backward()
    assert loss.item() == pytest.approx(rloss.item(), rel=1e-5)
    torch.testing.assert_close(pred.grad.float(), ref_pred.grad, rtol=2e-2, atol=1e-6)


def test_flux_schnell_variant_without_guidance_embedder_matches_oracle(doubles):
    """models/flux.py:475-479: Flux-schnell has no guidance embedder (diffusers CombinedTimestepTextProjEmbeddings); the
    guidance entry of the tuple still travels and is ignored (`guidance_embeds = false` in the transformer config)"""
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32,
           'guidance_embeds': False}
    torch.manual_seed(0)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': cfg}}, device='cpu')
    names = {n for n, _ in model.transformer.named_parameters()}
    assert not any('guidance_embedder' in n for n in names) and any('timestep_embedder' in n for n in names)
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=1, num_single=1, joint_dim=64, pooled_dim=32, guidance_embeds=False)
    assert names == {n for n, _ in ref.named_parameters()}
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    ref.load_state_dict({k: v.detach().float() for k, v in model.transformer.state_dict().items()})
    ref.set_emulate_bf16(True)
    g = torch.Generator().manual_seed(3)
    bs = 2
    latents, noise = torch.randn(bs, 16, 8, 8, generator=g), torch.randn(bs, 16, 8, 8, generator=g)
    t5, clip = torch.randn(bs, 12, 64, generator=g).bfloat16(), torch.randn(bs, 32, generator=g).bfloat16()
    feats, (target, _) = R.prepare_inputs(latents, t5, clip, torch.sigmoid(torch.randn(bs, generator=g)), noise)
    label = (target, torch.tensor([]))
    x = tuple(f.clone() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    loss = model.get_loss_fn()(x, label)
    loss.backward()
    y = tuple(f.clone() for f in feats)
    for layer in R.to_layers(ref):
        y = layer(y)
    rloss = R.loss_fn(y, label)
    rloss.backward()
    _compare(model.transformer.named_parameters(), ref, loss, rloss)


def test_bypass_guidance_embedding_drops_the_guidance_term(doubles):
    """models/flux.py:132-150: `bypass_guidance_embedding = true` keeps the guidance embedder's parameters but conditions on
    timestep + pooled text only — the schnell arithmetic on a dev checkpoint"""
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'bypass_guidance_embedding': True, 'transformer_config': cfg}},
                         device='cpu')
    assert any('guidance_embedder' in n for n, _ in model.transformer.named_parameters())
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=1, num_single=1, joint_dim=64, pooled_dim=32, guidance_embeds=False)
    sd = {k: v.detach().float() for k, v in model.transformer.state_dict().items() if 'guidance_embedder' not in k}
    ref.load_state_dict(sd)
    ref.set_emulate_bf16(True)
    g = torch.Generator().manual_seed(4)
    latents, noise = torch.randn(2, 16, 8, 8, generator=g), torch.randn(2, 16, 8, 8, generator=g)
    t5, clip = torch.randn(2, 12, 64, generator=g).bfloat16(), torch.randn(2, 32, generator=g).bfloat16()
    feats, (target, _) = R.prepare_inputs(latents, t5, clip, torch.sigmoid(torch.randn(2, generator=g)), noise, guidance=3.5)
    label = (target, torch.tensor([]))
    x = tuple(f.clone() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    loss = model.get_loss_fn()(x, label)
    loss.backward()
    y = tuple(f.clone() for f in feats)
    for layer in R.to_layers(ref):
        y = layer(y)
    rloss = R.loss_fn(y, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3
    assert all(p.grad is None for n, p in model.transformer.named_parameters() if 'guidance_embedder' in n)
