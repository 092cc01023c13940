"""GPU (one device): the two engine paths round 1 only exercised on CPU test doubles or on >= 2 GPUs.

  * evaluation (train.py:176-242): `prepare_inputs(batch, timestep_quantile=q)` for the reference's nine quantiles ->
    `engine.eval_batch` on the forward-only schedule, on the real kernels, against the oracle layers on the same tuples;
    no gradient state may be left behind and the RNG isolation of the caller is respected (prepare_inputs draws noise).
  * activation checkpointing (train.py:588-603, `activation_checkpointing = true`): every block layer wrapped in
    torch.utils.checkpoint — the fused autograd functions are recomputed in backward by the real kernels and must give the
    loss and the gradients of the run that keeps its activations, in the reference's 1F1B order and in the split-backward
    (zero-bubble) order, reentrant and non-reentrant.

Tolerances: evaluation loss within 5e-3 relative of the fp32 oracle (1e-3 of the oracle with the reference's bf16 rounding
points); checkpointed vs kept-activation runs: same loss to 1e-6 relative (the forward is the same launches) and gradients
within 2e-3 relative Frobenius (a few reductions accumulate with fp32 atomics: order-dependent last bits)."""
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = {'num_attention_heads': 2, 'num_layers': 2, 'num_single_layers': 2, 'joint_attention_dim': 64,
       'pooled_projection_dim': 32}
DEVICE = 'cuda'            # tests/test_real_shapes_host_logic.py runs this module's checks on the CPU kernel doubles
QUANTILES = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]         # train.py:176-180


def _pair(seed=0):
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    torch.manual_seed(seed)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'device': DEVICE, 'transformer_config': CFG}})
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=2, num_single=2, joint_dim=64, pooled_dim=32)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    missing, unexpected = ref.load_state_dict({k: v.detach().float().cpu() for k, v in model.transformer.state_dict().items()},
                                              strict=False)
    assert not missing and not unexpected
    return model, ref


def _batch_dict(bs, seed):
    g = torch.Generator().manual_seed(seed)
    return {'latents': torch.randn(bs, 16, 16, 16, generator=g), 't5_embed': torch.randn(bs, 32, 64, generator=g).bfloat16(),
            'clip_embed': torch.randn(bs, 32, generator=g).bfloat16(), 'mask': None}


def test_eval_batch_with_quantile_timesteps_matches_the_oracle():
    from diffusion_pipe_b200 import data_feed
    from diffusion_pipe_b200.pipe import ManualPipelineModule, initialize
    from oracle import flux_ref as R
    model, ref = _pair(1)
    n_mb = 2
    pm = ManualPipelineModule(layers=model.to_layers(), num_stages=1, partition_method='uniform', loss_fn=model.get_loss_fn(),
                              dynamic_shape=True, device=torch.device(DEVICE))
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': 1, 'gradient_accumulation_steps': n_mb,
                                                   'gradient_clipping': 1.0, 'steps_per_print': 0})
    losses = []
    for qi, q in enumerate(QUANTILES):
        torch.manual_seed(100 + qi)                       # the noise x_0 is drawn inside prepare_inputs
        feats, label = model.prepare_inputs(_batch_dict(n_mb, 50 + qi), timestep_quantile=q)
        want_t = torch.sigmoid(torch.distributions.normal.Normal(0, 1).icdf(torch.full((n_mb,), q)))
        assert torch.equal(feats[3], want_t)              # models/flux.py:352-358: every sample sits at the quantile
        mbs = list(data_feed.split_batch((feats, (label[0], label[1])), n_mb))
        engine.reset_activation_shape()
        loss = float(engine.eval_batch(iter(mbs), num_micro_batches=n_mb))
        per = {}
        for emu in (True, False):
            ref.set_emulate_bf16(emu)
            with torch.no_grad():
                tot = 0.0
                for f, l in mbs:
                    y = tuple(t.clone() for t in f)
                    for layer in R.to_layers(ref):
                        y = layer(y)
                    tot += float(R.loss_fn(y, l))
            per[emu] = tot / n_mb
        assert abs(loss - per[True]) / per[True] <= 1e-3, (q, loss, per)
        assert abs(loss - per[False]) / per[False] <= 5e-3, (q, loss, per)
        losses.append(loss)
    assert len(set(round(x, 6) for x in losses)) > 1     # different quantiles are different problems
    assert all(p.grad is None for p in pm.parameters())   # forward-only: no gradient state
    assert not any(engine.pipe_buffers[k][i] is not None for k in ('inputs', 'outputs', 'grads')
                   for i in range(len(engine.pipe_buffers[k])))
    # and the engine still trains afterwards (evaluation happens between training steps, train.py:936-940)
    engine._configure_optimizer(lambda ps: torch.optim.SGD(ps, lr=0.01), [p for p in pm.parameters()])
    torch.manual_seed(5)
    feats, label = model.prepare_inputs(_batch_dict(n_mb, 77))
    assert torch.isfinite(engine.train_batch(iter(list(data_feed.split_batch((feats, (label[0], label[1])), n_mb)))))


@pytest.mark.parametrize('reentrant', [False, True])
def test_activation_checkpointing_on_the_real_kernels(reentrant):
    from diffusion_pipe_b200.flux import FluxPipeline
    from diffusion_pipe_b200.pipe import ManualPipelineModule, initialize
    from oracle import flux_ref as R
    g = torch.Generator().manual_seed(1)
    mbs = []
    for _ in range(3):
        feats, (target, _) = R.prepare_inputs(torch.randn(1, 16, 16, 16, generator=g), torch.randn(1, 32, 64, generator=g).bfloat16(),
                                              torch.randn(1, 32, generator=g).bfloat16(), torch.sigmoid(torch.randn(1, generator=g)),
                                              torch.randn(1, 16, 16, 16, generator=g))
        mbs.append((feats, (target, torch.tensor([]))))
    results = []
    for ckpt, schedule in ((False, '1f1b'), (True, '1f1b'), (True, 'zb')):
        torch.manual_seed(0)
        model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'device': DEVICE, 'transformer_config': CFG}})
        extra = {}
        if ckpt:
            extra = {'activation_checkpoint_interval': 1, 'checkpointable_layers': model.checkpointable_layers,
                     'activation_checkpoint_func': partial(torch.utils.checkpoint.checkpoint, use_reentrant=reentrant)}
        pm = ManualPipelineModule(layers=model.to_layers(), num_stages=1, partition_method='uniform', loss_fn=model.get_loss_fn(),
                                  dynamic_shape=True, device=torch.device(DEVICE), **extra)
        engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': 1, 'gradient_accumulation_steps': len(mbs),
                                                       'gradient_clipping': 0.0, 'steps_per_print': 0, 'pipeline_schedule': schedule})
        captured = {}

        class Capture(torch.optim.SGD):
            def step(self, closure=None):
                captured.update({p.original_name: p.grad.detach().float().clone() for gr in self.param_groups for p in gr['params']})
        engine._configure_optimizer(lambda ps: Capture(ps, lr=0.0), [p for p in pm.parameters() if p.requires_grad])
        loss = float(engine.train_batch(iter([(tuple(f.clone() for f in fe), la) for fe, la in mbs])))
        if DEVICE == 'cuda':
            torch.cuda.synchronize()
        results.append((loss, captured, None))
    base_loss, base, _ = results[0]
    assert len(base) > 20
    for loss, grads, _ in results[1:]:
        assert loss == pytest.approx(base_loss, rel=1e-6)
        assert grads.keys() == base.keys()
        for k in base:
            rel = ((grads[k] - base[k]).norm() / (base[k].norm() + 1e-12)).item()
            assert rel <= 2e-3, (k, rel)


@pytest.mark.parametrize('family', ['flux', 'qwen', 'wan'])
def test_device_side_noising_is_bit_identical_to_the_host_path(family):
    """SURVEY.md 8(f)4, `device_prepare_inputs = true`: timesteps and noise still come from the host generator in the
    reference's order (models/flux.py:343-372), the mix / target / packing run in one device kernel — every tensor of the
    micro-batch must carry the bits of the host path (which tests/test_host_golden.py pins to the reference's own code)."""
    if DEVICE != 'cuda':
        pytest.skip('needs the device kernel')
    g = torch.Generator().manual_seed(9)
    bs = 2
    if family == 'flux':
        from diffusion_pipe_b200.flux import FluxPipeline as P
        cfg = {'transformer_config': CFG, 'guidance': 1.0}
        batch = {'latents': torch.randn(bs, 16, 32, 48, generator=g), 't5_embed': torch.randn(bs, 32, 64, generator=g).bfloat16(),
                 'clip_embed': torch.randn(bs, 32, generator=g).bfloat16(), 'mask': torch.rand(bs, 64, 96, generator=g)}
    elif family == 'qwen':
        from diffusion_pipe_b200.qwen_image import QwenImagePipeline as P
        cfg = {'transformer_config': {'num_attention_heads': 2, 'num_layers': 1, 'joint_attention_dim': 64}}
        batch = {'latents': torch.randn(bs, 16, 1, 16, 24, generator=g), 'prompt_embeds': [torch.randn(n, 64, generator=g).bfloat16() for n in (9, 12)],
                 'mask': None}
        from diffusion_pipe_b200 import data_feed
        batch = data_feed.BatchedDataset.collate([{'latents': batch['latents'][i], 'prompt_embeds': batch['prompt_embeds'][i], 'mask': None}
                                                  for i in range(bs)])
    else:
        from diffusion_pipe_b200.wan import WanPipeline as P
        cfg = {'transformer_config': {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 1, 'text_dim': 64, 'text_len': 16}}
        batch = {'latents': torch.randn(bs, 16, 3, 8, 12, generator=g), 'text_embeddings': torch.randn(bs, 16, 64, generator=g).bfloat16(),
                 'seq_lens': torch.tensor([10, 16]), 'mask': None}
    outs = []
    for on_device in (False, True):
        model = P({'model': dict(cfg, dtype='bfloat16', lazy_layers=True, device_prepare_inputs=on_device)})
        torch.manual_seed(31)
        outs.append(model.prepare_inputs({k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}))
    (f0, l0), (f1, l1) = outs
    assert any(torch.is_tensor(t) and t.is_cuda for t in f1) and l1[0].is_cuda       # the device path really ran
    for a, b in zip(list(f0) + list(l0), list(f1) + list(l1)):
        if a is None or b is None:
            assert a is None and b is None
            continue
        assert a.dtype == b.dtype and a.shape == b.shape
        assert torch.equal(a.cpu(), b.cpu())
