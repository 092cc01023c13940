"""CPU: the adapter arithmetic and the exported key layout against the reference TREE's own LoRA consumer
(submodules/ComfyUI/comfy/weight_adapter/lora.py, run by tests/golden/make_golden_lora.py -> lora_golden.pt):
W' = W + (alpha / rank) * (lora_B @ lora_A), alpha = rank.  PEFT itself is absent, so this is what pins oracle/lora_ref.py
(forward rule) and the product's K-extended GEMM operands (lora.py) to code the reference ships."""
import os
import re
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def g(golden_dir):
    return torch.load(os.path.join(golden_dir, 'lora_golden.pt'), weights_only=False)


def test_oracle_lora_linear_equals_the_merged_weight_of_the_reference_tree(g):
    from oracle import lora_ref
    from oracle.flux_ref import RefLinear
    base = RefLinear(g['K'], g['N'], bias=True)
    with torch.no_grad():
        base.weight.copy_(g['W'])
        base.bias.copy_(g['bias'])
    lin = lora_ref.RefLoraLinear(base, g['r'])
    with torch.no_grad():
        lin.lora_A.weight.copy_(g['A'])
        lin.lora_B.weight.copy_(g['B'])
    y = lin(g['x'])
    assert torch.allclose(y, g['y'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(g['W'] + g['B'] @ g['A'], g['merged'], rtol=1e-6, atol=1e-7)
    # and the gradient of the factors is the gradient of the merged weight pushed through B @ A
    y.sum().backward()
    gw = g['x'].sum(0)[None, :].expand(g['N'], -1)                    # d sum(y) / d W'
    assert torch.allclose(lin.lora_B.weight.grad, gw @ g['A'].t(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(lin.lora_A.weight.grad, g['B'].t() @ gw, rtol=1e-5, atol=1e-6)
    assert lin.weight.grad is None and not lin.weight.requires_grad


def test_k_extended_operands_compute_the_merged_linear(g, monkeypatch):
    """product side: [x | x A^T] . [W | B]^T on the (test-double) GEMM == x W'^T + b of the fixture, bf16 tolerance"""
    import kernel_doubles
    from diffusion_pipe_b200 import lora, ops
    from diffusion_pipe_b200.flux_blocks import _plain
    kernel_doubles.install(monkeypatch, ops)
    lin = _plain(g['N'], g['K'], torch.bfloat16, 'cpu')
    with torch.no_grad():
        lin.weight.copy_(g['W'])
        lin.bias.copy_(g['bias'])
    site = lora.LoraSite([lin], g['r'])
    with torch.no_grad():
        site.A[0].copy_(g['A'])
        site.B[0].copy_(g['B'])
    site.refresh()
    xa = site.alloc_in(g['x'].shape[0], 'cpu')
    xa[:, :site.K].copy_(g['x'])
    site.project(xa)
    y = ops.gemm(xa, site.w_fwd, bias=site.bias)
    rel = (y.float() - g['y']).norm() / g['y'].norm()
    assert rel <= 1e-2, rel.item()
    # input gradient: [dy | dy B] . [W ; A] == dy W'
    dy = torch.ones(g['x'].shape[0], g['N'])
    dya = site.alloc_dy(dy.shape[0], 'cpu')
    dya[:, :site.N].copy_(dy)
    site.backproject(dya)
    dx = ops.gemm(dya, site.w_dgrad, b_mn=True)
    want = dy @ g['merged']
    assert (dx.float() - want).norm() / want.norm() <= 1e-2


def test_exported_adapter_keys_have_the_layout_the_reference_tree_loads(g, tmp_path):
    """the fixture records which key names ComfyUI's loader accepted: `<prefix><module>.lora_A.weight` / `.lora_B.weight`"""
    from safetensors.torch import load_file
    from diffusion_pipe_b200 import lora
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    assert g['accepted_keys'] == [g['module'] + '.lora_A.weight', g['module'] + '.lora_B.weight']
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu',
                                         'transformer_config': {'num_attention_heads': 2, 'num_layers': 1, 'joint_attention_dim': 64}}})
    model.configure_adapter({'type': 'lora', 'rank': 16, 'alpha': 16, 'dropout': 0.0})
    model.save_adapter(str(tmp_path), lora.lora_state_dict(model.transformer))
    keys = set(load_file(str(tmp_path / 'adapter_model.safetensors')))
    assert g['module'] + '.lora_A.weight' in keys and g['module'] + '.lora_B.weight' in keys
    mods = {re.sub(r'\.lora_[AB]\.weight$', '', k) for k in keys}
    assert all(re.fullmatch(r'diffusion_model\..+\.lora_[AB]\.weight', k) for k in keys)
    assert all(m + '.lora_A.weight' in keys and m + '.lora_B.weight' in keys for m in mods) and len(keys) == 2 * len(mods)
