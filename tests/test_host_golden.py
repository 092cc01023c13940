"""CPU: the product's host-side functions of the hot path replayed against the outputs of the REFERENCE's own source text
(tests/golden/make_golden_host.py -> host_golden.pt): prepare_inputs of the three model families (same seed -> the same
timesteps, noise, packed tensors, ids, masks: bit-identical), the t-distribution helpers, and the default loss with its
mask / huber / smooth-L1 variants (value and gradient)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def g(golden_dir):
    return torch.load(os.path.join(golden_dir, 'host_golden.pt'), weights_only=False)


def _pipeline(family, cfg, extra):
    mc = dict({'dtype': 'bfloat16', 'device': 'cpu', 'lazy_layers': True}, **cfg)
    if family == 'flux':
        from diffusion_pipe_b200.flux import FluxPipeline
        return FluxPipeline({'model': mc})
    if family == 'qwen_image':
        from diffusion_pipe_b200.qwen_image import QwenImagePipeline
        return QwenImagePipeline({'model': mc})
    from diffusion_pipe_b200.wan import WanPipeline
    mc['transformer_config'] = {'model_type': extra['model_type']}        # (the reference reads it from the checkpoint's config.json)
    return WanPipeline({'model': mc})


def _same(a, b, what):
    if a is None or b is None:
        assert a is None and b is None, what
        return
    assert a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape), (what, a.dtype, b.dtype, tuple(a.shape), tuple(b.shape))
    assert torch.equal(a, b), (what, (a.float() - b.float()).abs().max().item())


def test_prepare_inputs_is_bit_identical_to_the_reference(g):
    seen = set()
    for c in g['cases']:
        model = _pipeline(c['family'], c['model_config'], c['extra'])
        torch.manual_seed(c['seed'])
        feats, (target, mask) = model.prepare_inputs(dict(c['inputs']), timestep_quantile=c['quantile'])
        what = f"{c['family']}/{c['name']}"
        assert len(feats) == len(c['features']), what
        for i, (a, b) in enumerate(zip(feats, c['features'])):
            _same(a, b, f'{what} feature {i}')
        _same(target, c['target'], what + ' target')
        _same(mask, c['mask'], what + ' mask')
        seen.add(c['family'])
    assert seen == {'flux', 'qwen_image', 'wan'} and len(g['cases']) == 16


def test_wan_t_distribution_table_matches(g):
    from diffusion_pipe_b200.wan import get_t_distribution
    for c in g['cases']:
        if c['family'] != 'wan':
            continue
        t = get_t_distribution(c['model_config'])
        assert len(t) == c['extra']['t_dist_len']
        assert torch.equal(t[:8], c['extra']['t_dist_head'])
        assert float(t.double().sum()) == c['extra']['t_dist_sum']


def test_default_loss_and_its_variants_match(g, monkeypatch):
    """models/base.py:418-436: value and d loss / d output; the product's fused masked-MSE kernel wrapper is replaced by its
    CPU test double, the huber / smooth-L1 variants run the product's own torch path"""
    import kernel_doubles
    from diffusion_pipe_b200 import ops
    from diffusion_pipe_b200.flux import FluxPipeline
    from oracle import flux_ref as R
    kernel_doubles.install(monkeypatch, ops)
    L = g['loss']
    for c in L['cases']:
        model = FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'lazy_layers': True}, **c['config']})
        m = L['mask'] if c['masked'] else torch.tensor([])
        o = L['output'].clone().requires_grad_(True)
        loss = model.get_loss_fn()(o, (L['target'], m))
        loss.backward()
        assert loss.item() == pytest.approx(c['loss'].item(), rel=2e-6), c['name']
        # the layer output is bf16, so its gradient is rounded to bf16; the fixture's is fp32
        rel = (o.grad.float() - c['dout']).norm() / c['dout'].norm()
        assert rel <= 4e-3, (c['name'], rel.item())
        if c['name'].startswith('mse'):
            # the oracle's loss is the same function
            o2 = L['output'].clone().float().requires_grad_(True)
            rl = R.loss_fn(o2, (L['target'], m))
            rl.backward()
            assert rl.item() == pytest.approx(c['loss'].item(), rel=1e-6) and torch.allclose(o2.grad, c['dout'], rtol=1e-6, atol=1e-9)
