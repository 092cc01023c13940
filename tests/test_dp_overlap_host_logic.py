"""CPU: the data-parallel overlap hooks (pipe/module.py `_GradsReady`, pipe/engine.py `_dp_reduce_layers`; the reference
reduces after the drain, utils/patches.py:152-156) on the three model definitions, kernels replaced by the PyTorch doubles of
tests/kernel_doubles.py:

  * a layer's callback runs BEFORE the backward of the layer in front of it has produced any parameter gradient — although
    every layer hands the time embedding (and more) through unchanged, whose own gradient completes only at the very end;
  * every parameter gradient of the layer exists when its callback runs;
  * the markers change no value: loss, input gradients and parameter gradients are bit-identical to the unarmed run."""
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


def _flux():
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'num_single_layers': 2, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': cfg}}, device='cpu')
    g = torch.Generator().manual_seed(1)
    latents, noise = torch.randn(2, 16, 8, 8, generator=g), torch.randn(2, 16, 8, 8, generator=g)
    t5, clip = torch.randn(2, 12, 64, generator=g).bfloat16(), torch.randn(2, 32, generator=g).bfloat16()
    feats, (target, _) = R.prepare_inputs(latents, t5, clip, torch.sigmoid(torch.randn(2, generator=g)), noise)
    return model, feats, (target, torch.tensor([]))


def _qwen():
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    from oracle import qwen_ref as Q
    torch.manual_seed(0)
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'transformer_config': {'num_attention_heads': 2, 'num_layers': 3,
                                                                                      'joint_attention_dim': 64}}}, device='cpu')
    g = torch.Generator().manual_seed(2)
    latents, noise = torch.randn(2, 16, 1, 8, 12, generator=g), torch.randn(2, 16, 1, 8, 12, generator=g)
    pe = [torch.randn(n, 64, generator=g).bfloat16().float() for n in (5, 11)]          # ragged prompts
    feats, (target, _) = Q.prepare_inputs(latents, pe, torch.sigmoid(torch.randn(2, generator=g)), noise)
    return model, feats, (target, torch.tensor([]))


def _wan():
    from diffusion_pipe_b200.wan import WanPipeline
    from oracle import wan_ref as W
    torch.manual_seed(0)
    model = WanPipeline({'model': {'dtype': 'bfloat16', 'transformer_config': {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 3,
                                                                               'text_dim': 64, 'text_len': 16}}}, device='cpu')
    g = torch.Generator().manual_seed(3)
    latents, noise = torch.randn(2, 16, 2, 8, 8, generator=g), torch.randn(2, 16, 2, 8, 8, generator=g)
    text = torch.randn(2, 16, 64, generator=g).bfloat16().float()
    feats, (target, _) = W.prepare_inputs(latents, text, torch.tensor([10, 16]), torch.sigmoid(torch.randn(2, generator=g)), noise)
    return model, feats, (target, torch.tensor([]))


def _randomise(model):
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in model.transformer.parameters():       # zero-initialised biases / gates would hide gradient paths
            if p.ndim == 1:
                p.add_((0.05 * torch.randn(p.shape, generator=g)).to(p.dtype))


@pytest.mark.parametrize('family', ['flux', 'qwen', 'wan'])
@pytest.mark.parametrize('interval', [0, 1])
def test_layer_callbacks_are_early_complete_and_change_nothing(doubles, family, interval):
    from functools import partial
    from diffusion_pipe_b200.pipe.module import PipelineModule
    runs = []
    for armed in (False, True):
        model, feats, label = {'flux': _flux, 'qwen': _qwen, 'wan': _wan}[family]()
        _randomise(model)
        extra = {}
        if interval:
            extra = {'activation_checkpoint_interval': interval, 'checkpointable_layers': model.checkpointable_layers,
                     'activation_checkpoint_func': partial(torch.utils.checkpoint.checkpoint, use_reentrant=False)}
        pm = PipelineModule(layers=model.to_layers(), num_stages=1, loss_fn=model.get_loss_fn(), device=torch.device('cpu'), **extra)
        funcs = pm.forward_funcs
        own = [[p for p in f.parameters() if p.requires_grad] if isinstance(f, torch.nn.Module) else [] for f in funcs]
        log = []

        def ready(first, last):
            have = [sum(p.grad is not None for p in own[i]) for i in range(len(funcs))]
            log.append((first, last, have))
        pm._grads_ready_cb = ready if armed else None
        x = tuple(f.clone().requires_grad_(f.is_floating_point()) for f in feats)
        loss = pm.loss_fn(pm(x), label)
        loss.backward()
        runs.append((loss.detach().clone(), [t.grad for t in x], {n: p.grad.clone() for n, p in pm.named_parameters() if p.grad is not None}))
        if not armed:
            continue
        n = len(funcs)
        assert [e[:2] for e in log] == [(i, i + 1) for i in reversed(range(n))][:len(log)] and len(log) >= n - 1, [e[:2] for e in log]
        for first, _last, have in log:
            assert have[first] == len(own[first]), (family, first, have[first], len(own[first]))       # complete
            for j in range(first):
                if own[j]:                                                                            # early: nothing in front of it has run
                    assert have[j] == 0, (family, 'callback of layer', first, 'ran after layer', j, 'produced gradients')
    (l0, g0, p0), (l1, g1, p1) = runs
    assert torch.equal(l0, l1)
    assert all((a is None and b is None) or torch.equal(a, b) for a, b in zip(g0, g1))
    assert p0.keys() == p1.keys() and p0 and all(torch.equal(p0[k], p1[k]) for k in p0)
