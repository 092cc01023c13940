"""CPU: Flux full-model export in the BFL / ComfyUI layout (flux_export.py, FluxPipeline.save_model) against the key table
the reference builds for its own export (tests/golden/flux_bfl_map.json, produced by executing
`make_diffusers_to_bfl_map` from models/flux.py), plus the concatenation order and the final-layer half swap of
models/flux.py:257-288."""
import json
import os

import pytest
import torch


def test_rules_reproduce_the_references_key_table(golden_dir):
    from diffusion_pipe_b200.flux_export import bfl_key
    table = json.load(open(os.path.join(golden_dir, 'flux_bfl_map.json')))
    assert len(table) == 1160
    for k, (pos, bk) in table.items():
        assert bfl_key(k) == (pos, bk), k
    with pytest.raises(KeyError):
        bfl_key('transformer_blocks.0.attn.to_q.lora_A.weight')
    with pytest.raises(KeyError):
        bfl_key('something_else.weight')


def test_every_parameter_of_the_model_is_in_the_layout_and_export_round_trips(golden_dir, tmp_path):
    from safetensors.torch import load_file
    from diffusion_pipe_b200.flux import FluxPipeline
    from diffusion_pipe_b200.flux_export import to_bfl
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'num_single_layers': 2, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    m = FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg}})
    with torch.no_grad():
        for p in m.transformer.parameters():
            p.normal_(0, 0.05)
    sd = {k: v.detach().clone() for k, v in m.transformer.state_dict().items()}
    table = json.load(open(os.path.join(golden_dir, 'flux_bfl_map.json')))
    assert all(k in table for k in sd)                               # same parameter names as the model the reference exports
    m.save_model(str(tmp_path / 'bfl'), sd)
    out = load_file(str(tmp_path / 'bfl' / 'model.safetensors'))
    assert out.keys() == to_bfl(sd).keys()
    # fused tensors are [q; k; v] (double) and [q; k; v; mlp] (single), models/flux.py:22-76
    a = 'transformer_blocks.1.attn.'
    assert torch.equal(out['double_blocks.1.img_attn.qkv.weight'], torch.cat([sd[a + 'to_q.weight'], sd[a + 'to_k.weight'], sd[a + 'to_v.weight']]))
    assert torch.equal(out['double_blocks.1.txt_attn.qkv.bias'], torch.cat([sd[a + 'add_q_proj.bias'], sd[a + 'add_k_proj.bias'], sd[a + 'add_v_proj.bias']]))
    s = 'single_transformer_blocks.0.'
    assert torch.equal(out['single_blocks.0.linear1.weight'], torch.cat([sd[s + 'attn.to_q.weight'], sd[s + 'attn.to_k.weight'],
                                                                        sd[s + 'attn.to_v.weight'], sd[s + 'proj_mlp.weight']]))
    assert torch.equal(out['single_blocks.0.norm.key_norm.scale'], sd[s + 'attn.norm_k.weight'])
    # final layer: diffusers stores (scale, shift), BFL (shift, scale)   (models/flux.py:280-288)
    w = sd['norm_out.linear.weight']
    assert torch.equal(out['final_layer.adaLN_modulation.1.weight'], torch.cat([w[w.shape[0] // 2:], w[:w.shape[0] // 2]]))
    assert torch.equal(out['img_in.weight'], sd['x_embedder.weight']) and torch.equal(out['guidance_in.out_layer.bias'], sd['time_text_embed.guidance_embedder.linear_2.bias'])
    assert sum(v.numel() for v in out.values()) == sum(v.numel() for v in sd.values())
    # the engine's own names on request
    m2 = FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'lazy_layers': True, 'export_layout': 'diffusers', 'transformer_config': cfg}})
    m2.save_model(str(tmp_path / 'dif'), sd)
    assert load_file(str(tmp_path / 'dif' / 'model.safetensors')).keys() == sd.keys()


def test_bfl_checkpoints_load_into_the_eager_and_the_stage_local_model(tmp_path):
    """`transformer_path` pointing at a BFL-layout file (flux1-dev.safetensors style; the reference hands it to diffusers'
    from_single_file, models/flux.py:174-182): export -> import is the identity, also when each stage reads only its layers"""
    from diffusion_pipe_b200.flux import FluxPipeline
    cfg = {'num_attention_heads': 2, 'num_layers': 2, 'num_single_layers': 2, 'joint_attention_dim': 64, 'pooled_projection_dim': 32}
    torch.manual_seed(0)
    src = FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg}})
    with torch.no_grad():
        for p in src.transformer.parameters():
            p.normal_(0, 0.05)
    sd = {k: v.detach().clone() for k, v in src.transformer.state_dict().items()}
    src.save_model(str(tmp_path), sd)
    f = str(tmp_path / 'model.safetensors')
    eager = FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg, 'transformer_path': f}})
    assert all(torch.equal(p, sd[n]) for n, p in eager.transformer.named_parameters())
    lazy = FluxPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': cfg, 'transformer_path': f, 'lazy_layers': True}})
    seen = set()
    for spec in lazy.to_layers():
        for n, p in spec.build().named_parameters():
            assert torch.equal(p, sd[p.original_name]), p.original_name
            seen.add(p.original_name)
    assert seen == set(sd)
