"""CPU: checkpoints in the reference's parameter names reach the model — the eagerly built transformer and, with
`lazy_layers`, exactly the layers a rank materialises (read tensor by tensor from the shards, by `original_name`); ComfyUI
style keys (`model.diffusion_model.` prefix, models/wan/wan.py:43-46) are accepted; missing tensors raise."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))

FLUX_CFG = {'num_attention_heads': 2, 'num_layers': 2, 'num_single_layers': 2, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}


def _flux(**extra):
    from diffusion_pipe_b200.flux import FluxPipeline
    return FluxPipeline({'model': dict({'dtype': 'bfloat16', 'guidance': 1.0, 'device': 'cpu', 'transformer_config': FLUX_CFG}, **extra)})


def _save_sharded(sd, d, prefix='', shards=2):
    from safetensors.torch import save_file
    os.makedirs(d, exist_ok=True)
    keys = sorted(sd)
    for i in range(shards):
        save_file({prefix + k: sd[k].contiguous() for k in keys[i::shards]}, os.path.join(d, f'model-{i + 1:05d}-of-{shards:05d}.safetensors'))


def test_eager_and_lazy_flux_load_the_same_checkpoint(tmp_path):
    torch.manual_seed(0)
    src = _flux()
    with torch.no_grad():
        for p in src.transformer.parameters():
            p.normal_(0, 0.05)
    sd = {k: v.detach().clone() for k, v in src.transformer.state_dict().items()}
    _save_sharded(sd, str(tmp_path / 'ckpt'))
    eager = _flux(transformer_path=str(tmp_path / 'ckpt'))
    assert all(torch.equal(p, sd[n]) for n, p in eager.transformer.named_parameters())
    lazy = _flux(transformer_path=str(tmp_path / 'ckpt'), lazy_layers=True)
    specs = lazy.to_layers()
    assert len(specs) == 1 + 2 + 2 + 1
    seen = set()
    for spec in specs:
        layer = spec.build()
        for n, p in layer.named_parameters():
            assert torch.equal(p, sd[p.original_name]), p.original_name
            seen.add(p.original_name)
    assert seen == set(sd)
    # with an adapter on an fp8 base: the checkpoint is read BEFORE the base is moved into fp8 storage
    lora_lazy = _flux(transformer_path=str(tmp_path / 'ckpt'), lazy_layers=True, transformer_dtype='float8')
    lora_lazy.configure_adapter({'type': 'lora', 'rank': 16, 'alpha': 16, 'dropout': 0.0})
    blk = lora_lazy.to_layers()[1].build().block
    w = blk.attn.to_q.weight
    assert w.dtype == torch.float8_e4m3fn and torch.equal(w.to(torch.bfloat16), sd['transformer_blocks.0.attn.to_q.weight'].to(torch.float8_e4m3fn).to(torch.bfloat16))


def test_missing_tensors_and_wrong_shapes_raise(tmp_path):
    src = _flux()
    sd = {k: v.detach().clone() for k, v in src.transformer.state_dict().items()}
    broken = dict(sd)
    del broken['single_transformer_blocks.1.proj_out.weight']
    _save_sharded(broken, str(tmp_path / 'missing'))
    with pytest.raises(RuntimeError, match='missing from the checkpoint'):
        _flux(transformer_path=str(tmp_path / 'missing'))
    lazy = _flux(transformer_path=str(tmp_path / 'missing'), lazy_layers=True)
    specs = lazy.to_layers()
    specs[1].build()                                            # a stage that does not own the broken layer is fine
    with pytest.raises(RuntimeError, match='missing from the checkpoint'):
        specs[4].build()
    bad = dict(sd)
    bad['x_embedder.weight'] = torch.zeros(3, 3)
    _save_sharded(bad, str(tmp_path / 'shape'))
    with pytest.raises(RuntimeError, match='shape'):
        _flux(transformer_path=str(tmp_path / 'shape'))
    os.makedirs(tmp_path / 'none')
    with pytest.raises(RuntimeError, match='no \\*.safetensors'):
        _flux(transformer_path=str(tmp_path / 'none'))


def test_wan_accepts_comfyui_keys_in_one_file(tmp_path):
    from safetensors.torch import save_file
    from diffusion_pipe_b200.wan import WanPipeline
    cfg = {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 2, 'text_dim': 64, 'text_len': 16}
    src = WanPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg}})
    with torch.no_grad():
        for p in src.transformer.parameters():
            p.normal_(0, 0.05)
    sd = {k: v.detach().clone() for k, v in src.transformer.state_dict().items()}
    f = str(tmp_path / 'wan.safetensors')
    save_file({'model.diffusion_model.' + k: v.contiguous() for k, v in sd.items()}, f)
    eager = WanPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg, 'transformer_path': f}})
    assert all(torch.equal(p, sd[n]) for n, p in eager.transformer.named_parameters())
    lazy = WanPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg, 'transformer_path': f, 'lazy_layers': True}})
    for spec in lazy.to_layers():
        for n, p in spec.build().named_parameters():
            assert torch.equal(p, sd[p.original_name]), p.original_name


def test_diffusers_directory_layout(tmp_path):
    """models/flux.py:174-191: `diffusers_path/transformer/` holds the shards and a config.json naming the architecture"""
    import json
    torch.manual_seed(1)
    cfg = dict(FLUX_CFG, guidance_embeds=False, _class_name='FluxTransformer2DModel', patch_size=1)
    src = _flux(transformer_config={k: v for k, v in cfg.items() if not k.startswith('_') and k != 'patch_size'})
    sd = {k: v.detach().clone() for k, v in src.transformer.state_dict().items()}
    assert not any('guidance_embedder' in k for k in sd)
    d = tmp_path / 'flux' / 'transformer'
    _save_sharded(sd, str(d))
    with open(d / 'config.json', 'w') as f:
        json.dump(cfg, f)
    from diffusion_pipe_b200.flux import FluxPipeline
    m = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'device': 'cpu', 'diffusers_path': str(tmp_path / 'flux')}})
    assert m.tcfg['num_layers'] == 2 and m.tcfg['guidance_embeds'] is False
    assert all(torch.equal(p, sd[n]) for n, p in m.transformer.named_parameters()) and len(sd) == len(list(m.transformer.parameters()))


def test_wan_reads_the_checkpoints_config_json(tmp_path):
    """models/wan/wan.py:80-135: the architecture comes from the config.json beside the shards; a Wan2.2 I2V checkpoint says
    model_type 'i2v' but has no CLIP image branch -> 'i2v_v2'"""
    import json
    from diffusion_pipe_b200.wan import WanPipeline
    small = {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 1, 'text_dim': 64, 'text_len': 16}
    src = WanPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': dict(small, model_type='i2v_v2')}})
    sd = {k: v.detach().clone() for k, v in src.transformer.state_dict().items()}
    assert sd['patch_embedding.weight'].shape[1] == 36
    d = tmp_path / 'wan22_i2v'
    _save_sharded(sd, str(d))
    with open(d / 'config.json', 'w') as f:
        json.dump(dict(small, model_type='i2v', in_dim=36, out_dim=16, freq_dim=256, eps=1e-6, _class_name='WanModel'), f)
    m = WanPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'ckpt_path': str(d)}})
    assert m.model_type == 'i2v_v2' and m.tcfg['dim'] == 256 and m.tcfg['num_layers'] == 1
    assert all(torch.equal(p, sd[n]) for n, p in m.transformer.named_parameters())
