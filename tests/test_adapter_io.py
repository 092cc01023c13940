"""CPU: `[adapter] init_from_existing` (train.py:534-535 -> models/base.py:367-388): a saved adapter is read back into the
factors — eager models, lazily built stage-local layers, the ComfyUI `diffusion_model.` prefix of the Qwen / Wan exports —
and malformed inputs fail the way the reference's loader does."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

FLUX_CFG = {'num_attention_heads': 2, 'num_layers': 1, 'num_single_layers': 1, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
ADAPTER = {'type': 'lora', 'rank': 16, 'alpha': 16, 'dropout': 0.0}


def _flux(lazy=False, **extra):
    from diffusion_pipe_b200.flux import FluxPipeline
    mc = dict({'dtype': 'bfloat16', 'guidance': 1.0, 'device': 'cpu', 'transformer_config': FLUX_CFG}, **extra)
    if lazy:
        mc['lazy_layers'] = True
    return FluxPipeline({'model': mc})


def _randomise_factors(module, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    with torch.no_grad():
        for n, p in module.named_parameters():
            if '.lora_A.' in n or '.lora_B.' in n:
                p.copy_((0.05 * torch.randn(p.shape, generator=g)).to(p.dtype))
                out[p.original_name] = p.detach().clone()
    return out


def test_saved_flux_adapter_initialises_eager_and_lazy_models(tmp_path):
    from diffusion_pipe_b200 import lora
    src = _flux()
    src.configure_adapter(ADAPTER)
    want = _randomise_factors(src.transformer, 3)
    assert len(want) == 40
    src.save_adapter(str(tmp_path / 'ad'), lora.lora_state_dict(src.transformer))
    assert os.listdir(tmp_path / 'ad') == ['pytorch_lora_weights.safetensors']        # diffusers' save_lora_weights layout
    # eager model
    dst = _flux()
    dst.configure_adapter(ADAPTER)
    dst.load_adapter_weights(str(tmp_path / 'ad'))
    got = {p.original_name: p for n, p in dst.transformer.named_parameters() if '.lora_' in n}
    assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)
    # the next forward sees them: the site operand is rebuilt from the loaded factors
    site = dst.transformer.transformer_blocks[0].lora['qkv']
    site.refresh()
    assert torch.equal(site.a_all[:16], want['transformer_blocks.0.attn.to_q.lora_A.weight'])
    assert torch.equal(site.b_blk[:256, :16], want['transformer_blocks.0.attn.to_q.lora_B.weight'])
    # lazily built, stage-local layers (what the pipeline engine materialises on each rank)
    lazy = _flux(lazy=True)
    lazy.configure_adapter(ADAPTER)
    lazy.load_adapter_weights(str(tmp_path / 'ad'))
    specs = lazy.to_layers()
    for spec in specs[1:3]:
        layer = spec.build()
        facs = {p.original_name: p for n, p in layer.named_parameters() if '.lora_' in n}
        assert facs and all(torch.equal(p, want[k]) for k, p in facs.items())


def test_comfyui_prefix_and_fp8_base(tmp_path):
    """Qwen / Wan exports carry `diffusion_model.` in front of every key (models/qwen_image.py:290-294); loading into a
    model whose frozen base is stored in fp8 works the same"""
    from diffusion_pipe_b200 import lora
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    cfg = {'num_attention_heads': 2, 'num_layers': 1, 'joint_attention_dim': 64}
    src = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg}})
    src.configure_adapter(ADAPTER)
    want = _randomise_factors(src.transformer, 5)
    src.save_adapter(str(tmp_path / 'q'), lora.lora_state_dict(src.transformer))
    from safetensors.torch import load_file
    assert all(k.startswith('diffusion_model.') for k in load_file(str(tmp_path / 'q' / 'adapter_model.safetensors')))
    import json
    pc = json.load(open(tmp_path / 'q' / 'adapter_config.json'))                 # peft_config.save_pretrained (models/qwen_image.py:291)
    assert pc['peft_type'] == 'LORA' and pc['r'] == pc['lora_alpha'] == 16 and 'transformer_blocks.0.attn.to_q' in pc['target_modules'] and len(pc['target_modules']) == 14
    dst = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_dtype': 'float8', 'transformer_config': cfg}})
    dst.configure_adapter(ADAPTER)
    dst.load_adapter_weights(str(tmp_path / 'q'))
    got = {p.original_name: p for n, p in dst.transformer.named_parameters() if '.lora_' in n}
    assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)


def test_malformed_adapter_directories_raise(tmp_path):
    from safetensors.torch import save_file
    m = _flux()
    m.configure_adapter(ADAPTER)
    (tmp_path / 'empty').mkdir()
    with pytest.raises(RuntimeError, match='No safetensors'):
        m.load_adapter_weights(str(tmp_path / 'empty'))
    (tmp_path / 'two').mkdir()
    for f in ('a.safetensors', 'b.safetensors'):
        save_file({'x': torch.zeros(1)}, str(tmp_path / 'two' / f))
    with pytest.raises(RuntimeError, match='Multiple'):
        m.load_adapter_weights(str(tmp_path / 'two'))
    (tmp_path / 'bad').mkdir()
    save_file({'transformer.transformer_blocks.0.attn.to_q.lora_Z.weight': torch.zeros(16, 256)}, str(tmp_path / 'bad' / 'a.safetensors'))
    with pytest.raises(RuntimeError, match='not in the model parameters'):
        m.load_adapter_weights(str(tmp_path / 'bad'))
    (tmp_path / 'shape').mkdir()
    save_file({'transformer_blocks.0.attn.to_q.lora_A.weight': torch.zeros(8, 256)}, str(tmp_path / 'shape' / 'a.safetensors'))
    with pytest.raises(RuntimeError, match='shape'):
        m.load_adapter_weights(str(tmp_path / 'shape'))
