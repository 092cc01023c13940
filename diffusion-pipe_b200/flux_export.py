"""Flux full-model export in the BFL / ComfyUI checkpoint layout — what the reference's FluxPipeline.save_model writes
(models/flux.py:257-288: diffusers-named tensors regrouped under BFL names, the projections a BFL block keeps fused are
concatenated in a fixed order, and the final adaLN linear stores (shift, scale) where diffusers stores (scale, shift)).

The layout is stated here as rules over the module tree rather than as a key table; tests/test_flux_export.py holds the
rules to the table the reference builds (tests/golden/flux_bfl_map.json, produced by running its own map builder).
"""
import re

import torch

_GLOBAL = {   # diffusers module -> BFL module (weight and bias alike)
    'x_embedder': 'img_in', 'context_embedder': 'txt_in', 'proj_out': 'final_layer.linear',
    'norm_out.linear': 'final_layer.adaLN_modulation.1',
    'time_text_embed.timestep_embedder.linear_1': 'time_in.in_layer', 'time_text_embed.timestep_embedder.linear_2': 'time_in.out_layer',
    'time_text_embed.text_embedder.linear_1': 'vector_in.in_layer', 'time_text_embed.text_embedder.linear_2': 'vector_in.out_layer',
    'time_text_embed.guidance_embedder.linear_1': 'guidance_in.in_layer', 'time_text_embed.guidance_embedder.linear_2': 'guidance_in.out_layer',
}
_DOUBLE = {   # inside transformer_blocks.N.  ->  double_blocks.N.   (position in the fused tensor, BFL module)
    'norm1.linear': (0, 'img_mod.lin'), 'norm1_context.linear': (0, 'txt_mod.lin'),
    'attn.to_q': (0, 'img_attn.qkv'), 'attn.to_k': (1, 'img_attn.qkv'), 'attn.to_v': (2, 'img_attn.qkv'),
    'attn.add_q_proj': (0, 'txt_attn.qkv'), 'attn.add_k_proj': (1, 'txt_attn.qkv'), 'attn.add_v_proj': (2, 'txt_attn.qkv'),
    'attn.to_out.0': (0, 'img_attn.proj'), 'attn.to_add_out': (0, 'txt_attn.proj'),
    'ff.net.0.proj': (0, 'img_mlp.0'), 'ff.net.2': (0, 'img_mlp.2'),
    'ff_context.net.0.proj': (0, 'txt_mlp.0'), 'ff_context.net.2': (0, 'txt_mlp.2'),
}
_DOUBLE_NORM = {'attn.norm_q': 'img_attn.norm.query_norm', 'attn.norm_k': 'img_attn.norm.key_norm',
                'attn.norm_added_q': 'txt_attn.norm.query_norm', 'attn.norm_added_k': 'txt_attn.norm.key_norm'}
_SINGLE = {   # inside single_transformer_blocks.N.  ->  single_blocks.N.
    'norm.linear': (0, 'modulation.lin'),
    'attn.to_q': (0, 'linear1'), 'attn.to_k': (1, 'linear1'), 'attn.to_v': (2, 'linear1'), 'proj_mlp': (3, 'linear1'),
    'proj_out': (0, 'linear2'),
}
_SINGLE_NORM = {'attn.norm_q': 'norm.query_norm', 'attn.norm_k': 'norm.key_norm'}


def bfl_key(diffusers_key):
    """(position inside the BFL tensor, BFL key) of one diffusers-named tensor; KeyError for a name outside the layout"""
    mod, _, leaf = diffusers_key.rpartition('.')
    if mod in _GLOBAL:
        return 0, f'{_GLOBAL[mod]}.{leaf}'
    m = re.fullmatch(r'(transformer_blocks|single_transformer_blocks)\.(\d+)\.(.+)', mod)
    if m is None:
        raise KeyError(f'Key not found in the diffusers -> BFL layout: {diffusers_key}')
    double = m.group(1) == 'transformer_blocks'
    prefix = f"{'double_blocks' if double else 'single_blocks'}.{m.group(2)}."
    table, norms = (_DOUBLE, _DOUBLE_NORM) if double else (_SINGLE, _SINGLE_NORM)
    if m.group(3) in table:
        pos, name = table[m.group(3)]
        return pos, f'{prefix}{name}.{leaf}'
    if m.group(3) in norms and leaf == 'weight':
        return 0, f'{prefix}{norms[m.group(3)]}.scale'
    raise KeyError(f'Key not found in the diffusers -> BFL layout: {diffusers_key}')


def to_bfl(diffusers_sd):
    """{BFL key: tensor}: fused tensors concatenated along dim 0 in position order, final adaLN linear with its two halves
    swapped ((scale, shift) -> (shift, scale)), as models/flux.py:257-288 does"""
    parts = {}
    for k, t in diffusers_sd.items():
        pos, bk = bfl_key(k)
        parts.setdefault(bk, []).append((pos, t))
    out = {}
    for bk, vals in parts.items():
        out[bk] = vals[0][1] if len(vals) == 1 else torch.cat([t for _, t in sorted(vals, key=lambda v: v[0])])
    for leaf in ('weight', 'bias'):
        k = f'final_layer.adaLN_modulation.1.{leaf}'
        if k in out:
            a, b = out[k].chunk(2, dim=0)
            out[k] = torch.cat([b, a], dim=0)
    return out


def from_bfl_plan(named_shapes):
    """How to fill diffusers-named parameters from a BFL-layout checkpoint (flux1-dev.safetensors and friends: what
    `transformer_path` usually points at; the reference lets diffusers' from_single_file do this, models/flux.py:174-182).
    named_shapes: [(diffusers name, shape)] of ONE module tree whose fused siblings are all present (a block, or the whole
    model).  Returns {name: (bfl key, first row, rows, swap_halves)}."""
    groups = {}
    for name, shape in named_shapes:
        pos, bk = bfl_key(name)
        groups.setdefault(bk, []).append((pos, name, int(shape[0])))
    plan = {}
    for bk, items in groups.items():
        off = 0
        for pos, name, rows in sorted(items):
            plan[name] = (bk, off, rows, bk.startswith('final_layer.adaLN_modulation.1.'))
            off += rows
    return plan


def read_bfl_tensor(tensor, first_row, rows, swap_halves):
    if swap_halves:                          # BFL (shift, scale) -> diffusers (scale, shift)
        a, b = tensor.chunk(2, dim=0)
        tensor = torch.cat([b, a], dim=0)
    return tensor[first_row:first_row + rows]
