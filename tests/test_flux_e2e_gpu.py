"""GPU: a small Flux-architecture model through the reference-facing plugin API (FluxPipeline.to_layers /
prepare_inputs-shaped tuples / get_loss_fn) and through the pipeline engine (train_batch), against the oracle
(RefFluxTransformer + RefPipelineEngine) on identical noised-latent inputs, weights and optimizer.

Tolerance (north star): loss within 1e-3 relative of the oracle with the reference's bf16 rounding points emulated,
5e-3 of the pure-fp32 oracle (bf16 tensor-core compute, fp32 accumulation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = {'num_attention_heads': 2, 'num_layers': 2, 'num_single_layers': 2, 'joint_attention_dim': 64,
       'pooled_projection_dim': 32}


def _make(seed=0):
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    torch.manual_seed(seed)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': CFG}})
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=2, num_single=2, joint_dim=64, pooled_dim=32)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    sd = {k: v.detach().float().cpu() for k, v in model.transformer.state_dict().items()}
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model, ref


def _batch(bs, seed):
    from oracle import flux_ref as R
    g = torch.Generator().manual_seed(seed)
    latents = torch.randn(bs, 16, 16, 16, generator=g)
    t5 = torch.randn(bs, 32, 64, generator=g).bfloat16()
    clip = torch.randn(bs, 32, generator=g).bfloat16()
    t = torch.sigmoid(torch.randn(bs, generator=g))
    noise = torch.randn(bs, 16, 16, 16, generator=g)
    feats, (target, mask) = R.prepare_inputs(latents, t5, clip, t, noise)
    return feats, (target, torch.tensor([]))


def test_layers_and_loss_match_oracle():
    from oracle import flux_ref as R
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
        for layer in R.to_layers(ref):
            y = layer(y)
        rloss = R.loss_fn(y, label)
        rel = abs(loss.item() - rloss.item()) / abs(rloss.item())
        assert rel <= tol, (emu, loss.item(), rloss.item(), rel)
        if emu:   # parameter gradients against the oracle with the reference rounding points
            rloss.backward()
            rg = {n: p.grad for n, p in ref.named_parameters()}
            errs = {}
            for n, p in model.transformer.named_parameters():
                assert p.grad is not None, n
                errs[n] = ((p.grad.float().cpu() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
            bad = sorted(((v, k) for k, v in errs.items() if v > 5e-2), reverse=True)
            assert not bad, bad[:8]


def test_engine_train_batch_matches_oracle_engine():
    from diffusion_pipe_b200.pipe import ManualPipelineModule, initialize
    from oracle import flux_ref as R
    from oracle.engine_ref import RefPipelineEngine
    model, ref = _make(3)
    gas, mbs = 2, 1
    pm = ManualPipelineModule(layers=model.to_layers(), num_stages=1, partition_method='parameters',
                              manual_partition_split=None, loss_fn=model.get_loss_fn(), dynamic_shape=True)
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': mbs,
                                                   'gradient_accumulation_steps': gas, 'gradient_clipping': 1.0,
                                                   'steps_per_print': 0})
    params = [p for p in pm.parameters() if p.requires_grad]
    engine._configure_optimizer(lambda ps: torch.optim.SGD(ps, lr=0.05), params)
    ref.set_emulate_bf16(False)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05)
    reng = RefPipelineEngine(R.to_layers(ref), R.loss_fn, ropt, None, gas, 1.0)
    mbatches = [_batch(mbs, 10 + i) for i in range(gas)]
    loss = engine.train_batch(iter(mbatches))
    rloss = reng.train_batch(mbatches)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 5e-3, (loss.item(), rloss.item())
    gn = float(engine._grad_norm)
    assert abs(gn - float(reng.grad_norm)) / float(reng.grad_norm) <= 3e-2, (gn, float(reng.grad_norm))
    # the optimizer stepped on both sides: updated weights still agree to bf16 resolution
    rsd = ref.state_dict()
    for n, p in model.transformer.named_parameters():
        d = (p.detach().float().cpu() - rsd[n]).abs().max().item()
        assert d <= 2e-2 * (rsd[n].abs().max().item() + 1e-3) + 1e-3, (n, d)
    assert all(p.grad is None for p in params)
    # a second step runs (gradient buffers are re-attached after zero_grad(set_to_none=True))
    loss2 = engine.train_batch(iter([_batch(mbs, 20 + i) for i in range(gas)]))
    assert torch.isfinite(loss2)
