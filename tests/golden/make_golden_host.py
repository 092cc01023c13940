"""Generates tests/golden/host_golden.pt by running the REFERENCE's own host-side functions of the hot path:

    FluxPipeline.prepare_inputs        models/flux.py:323-394        
    QwenImagePipeline.prepare_inputs   models/qwen_image.py:394-480
    WanPipeline.prepare_inputs         models/wan/wan.py:371-412
    BasePipeline.get_loss_fn           models/base.py:418-436        (row D5)
    get_t_distribution / slice_t_distribution / sample_t / time_shift / get_lin_function   utils/common.py:114-165

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_host.py

The modules themselves cannot be imported here (diffusers / peft / deepspeed / accelerate are absent), so each function is
taken verbatim from the source text (ast) and executed with a stand-in `self` carrying only the attributes the function
reads.  Two diffusers helpers the methods call are restated (recalled, marked below): `_prepare_latent_image_ids` and
`_pack_latents`.  `get_t_distribution(...).to('cuda')` (wan.py:78) is not executed: the table is built on the CPU.

Every case seeds torch's global RNG right before the call; the fixture stores inputs that cannot be regenerated, the
outputs, and the seed.  tests/test_host_golden.py replays the product's prepare_inputs / loss functions on the same
inputs with the same seed: integer / bool outputs and every float output must be bit-identical.
"""
import ast
import math
import os
import sys

import torch
import torch.nn.functional as F
from einops import rearrange

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import synth_tensor  # noqa: E402

REF = '/root/reference'
OUT = os.path.join(HERE, 'host_golden.pt')


def extract(path, names, cls=None):
    """{name: function} compiled from the source text of `path` (module level, or methods of class `cls`)"""
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fns = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {f.name for f in fns} == set(names), (path, names, [f.name for f in fns])
    ns = {'torch': torch, 'F': F, 'math': math, 'rearrange': rearrange}
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, 'exec'), ns)
    return ns


def latent_image_ids(batch_size, height, width, device, dtype):
    """diffusers FluxPipeline._prepare_latent_image_ids (RECALLED: diffusers is absent)"""
    ids = torch.zeros(height, width, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width)[None, :]
    return ids.reshape(height * width, 3).to(device=device, dtype=dtype)


def pack_latents(latents, batch_size, num_channels_latents, height, width):
    """diffusers QwenImagePipeline._pack_latents (RECALLED)"""
    latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
    latents = latents.permute(0, 2, 4, 1, 3, 5)
    return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)


class Self:
    pass


def main():
    common = extract(f'{REF}/utils/common.py', ['time_shift', 'get_lin_function', 'get_t_distribution', 'slice_t_distribution', 'sample_t'])
    g = {'cases': []}

    def record(family, name, seed, inputs, quantile, cfg, feats, label, extra=None):
        g['cases'].append({'family': family, 'name': name, 'seed': seed, 'inputs': inputs, 'quantile': quantile, 'model_config': cfg,
                           'features': tuple(feats), 'target': label[0], 'mask': label[1], 'extra': extra or {}})

    # ------------------------------------------------------------------------------------------------- Flux
    flux = extract(f'{REF}/models/flux.py', ['prepare_inputs'], cls='FluxPipeline')
    flux.update({k: common[k] for k in ('time_shift', 'get_lin_function')})
    flux_fn = flux['prepare_inputs']
    flux_mod = extract(f'{REF}/models/flux.py', ['time_shift', 'get_lin_function'])
    flux['time_shift'], flux['get_lin_function'] = flux_mod['time_shift'], flux_mod['get_lin_function']
    bs, h, w = 2, 8, 12
    base_in = {'latents': synth_tensor((bs, 16, h, w), 801, 1.0), 't5_embed': synth_tensor((bs, 6, 32), 802, 1.0).bfloat16(),
               'clip_embed': synth_tensor((bs, 16), 803, 1.0).bfloat16(), 'mask': None}
    pix_mask = (synth_tensor((bs, 8 * h, 8 * w), 804, 1.0) > 0).to(torch.float16)
    flux_cases = [
        ('default', {'guidance': 1.0}, dict(base_in), None),
        ('uniform', {'guidance': 3.5, 'timestep_sample_method': 'uniform'}, dict(base_in), None),
        ('sigmoid_scale_shift', {'guidance': 1.0, 'sigmoid_scale': 1.3, 'shift': 3.0}, dict(base_in), None),
        ('flux_shift', {'guidance': 1.0, 'flux_shift': True}, dict(base_in), None),
        ('quantile', {'guidance': 1.0}, dict(base_in), 0.3),
        ('mask', {'guidance': 1.0}, dict(base_in, mask=pix_mask), None),
        ('kontext', {'guidance': 1.0}, dict(base_in, control_latents=synth_tensor((bs, 16, h, w), 805, 1.0)), None),
    ]
    for i, (name, cfg, inputs, q) in enumerate(flux_cases):
        s = Self()
        s.model_config, s.is_flex2, s._prepare_latent_image_ids = cfg, False, latent_image_ids
        seed = 100 + i
        torch.manual_seed(seed)
        feats, label = flux_fn(s, inputs, timestep_quantile=q)
        record('flux', name, seed, inputs, q, cfg, feats, label)

    # ------------------------------------------------------------------------------------------------- Qwen-Image
    qwen = extract(f'{REF}/models/qwen_image.py', ['prepare_inputs'], cls='QwenImagePipeline')
    qwen.update({k: common[k] for k in ('time_shift', 'get_lin_function')})
    qwen_fn = qwen['prepare_inputs']
    qin = {'latents': synth_tensor((bs, 16, 1, h, w), 811, 1.0), 'mask': None,
           'prompt_embeds': [synth_tensor((5, 24), 812, 1.0).bfloat16(), synth_tensor((9, 24), 813, 1.0).bfloat16()]}
    qwen_cases = [
        ('ragged_prompts', {}, dict(qin), None),
        ('shift_quantile', {'shift': 2.0}, dict(qin), 0.7),
        ('mask', {}, dict(qin, mask=pix_mask), None),
        ('control', {'flux_shift': True}, dict(qin, control_latents=synth_tensor((bs, 16, 1, h, w), 814, 1.0)), None),
    ]
    for i, (name, cfg, inputs, q) in enumerate(qwen_cases):
        s = Self()
        s.model_config, s._pack_latents = cfg, pack_latents
        s.transformer = Self()
        s.transformer.config = Self()
        s.transformer.config.in_channels = 64
        seed = 200 + i
        torch.manual_seed(seed)
        feats, label = qwen_fn(s, inputs, timestep_quantile=q)
        record('qwen_image', name, seed, inputs, q, cfg, feats, label)

    # ------------------------------------------------------------------------------------------------- Wan
    wan = extract(f'{REF}/models/wan/wan.py', ['prepare_inputs'], cls='WanPipeline')
    wan.update({k: common[k] for k in ('time_shift', 'get_lin_function', 'slice_t_distribution', 'sample_t')})
    wan_fn = wan['prepare_inputs']
    win = {'latents': synth_tensor((bs, 16, 3, h, w), 821, 1.0), 'mask': None,
           'text_embeddings': synth_tensor((bs, 12, 24), 822, 1.0).bfloat16(), 'seq_lens': torch.tensor([7, 12])}
    wan_cases = [
        ('t2v_default', 't2v', {}, dict(win), None),
        ('t2v_uniform_minmax', 't2v', {'timestep_sample_method': 'uniform', 'min_t': 0.2, 'max_t': 0.9}, dict(win), None),
        ('t2v_shift_quantile', 't2v', {'shift': 5.0}, dict(win), 0.5),
        ('t2v_flux_shift_mask', 't2v', {'flux_shift': True, 'sigmoid_scale': 0.8}, dict(win, mask=pix_mask), None),
        ('i2v_v2', 'i2v_v2', {}, dict(win, y=synth_tensor((bs, 16, 3, h, w), 823, 1.0)), None),
    ]
    for i, (name, mtype, cfg, inputs, q) in enumerate(wan_cases):
        s = Self()
        s.model_config, s.model_type, s.cache_text_embeddings = cfg, mtype, True
        s.t_dist = common['get_t_distribution'](cfg)             # (wan.py:78 moves it to 'cuda'; CPU here)
        seed = 300 + i
        torch.manual_seed(seed)
        feats, label = wan_fn(s, inputs, timestep_quantile=q)
        record('wan', name, seed, inputs, q, cfg, feats, label, {'model_type': mtype, 't_dist_head': s.t_dist[:8].clone(),
                                                                 't_dist_sum': float(s.t_dist.double().sum()), 't_dist_len': len(s.t_dist)})

    # ------------------------------------------------------------------------------------------------- default loss
    loss_ns = extract(f'{REF}/models/base.py', ['get_loss_fn'], cls='BasePipeline')
    out = synth_tensor((2, 24, 64), 831, 1.0).bfloat16()
    tgt = synth_tensor((2, 24, 64), 832, 1.0)
    msk = (synth_tensor((2, 24, 64), 833, 1.0) > 0).to(torch.float16)
    g['loss'] = {'output': out, 'target': tgt, 'mask': msk, 'cases': []}
    for name, cfg, m in (('mse', {}, torch.tensor([])), ('mse_masked', {}, msk), ('huber', {'huber_delta': 0.7}, torch.tensor([])),
                         ('huber_masked', {'huber_delta': 0.7}, msk), ('smooth_l1', {'smooth_l1_beta': 0.4}, msk)):
        s = Self()
        s.config = cfg
        o = out.clone().float().requires_grad_(True)
        loss = loss_ns['get_loss_fn'](s)(o, (tgt, m))
        loss.backward()
        g['loss']['cases'].append({'name': name, 'config': cfg, 'masked': m.numel() > 0, 'loss': loss.detach(), 'dout': o.grad.clone()})

    torch.save(g, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes;', len(g['cases']), 'prepare_inputs cases,', len(g['loss']['cases']), 'loss cases')


if __name__ == '__main__':
    main()
