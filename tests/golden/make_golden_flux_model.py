"""Generates tests/golden/flux_model_golden.pt: a WHOLE Flux-architecture model (embedders, one double + one single
block, final layer), forward, run by the reference tree's in-tree BFL-layout implementation and tied to the
diffusers layout by the reference's own key map.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_flux_model.py

Sources of truth, both used as they are:
  * /root/reference/submodules/ComfyUI/comfy/ldm/flux/model.py (`Flux.forward_orig`: img_in / time_in / guidance_in /
    vector_in / txt_in, EmbedND, DoubleStreamBlock, SingleStreamBlock, LastLayer) with layers.py (`timestep_embedding`
    with time_factor 1000, Modulation chunk order, QKNorm);
  * /root/reference/models/flux.py:22-76 `BFL_TO_DIFFUSERS_MAP` (read from the source text with `ast`, not retyped) and
    the final-layer (shift, scale) <-> (scale, shift) swap of models/flux.py:280-288 — the reference's own statement of
    how a diffusers-layout Flux (what it trains, and what this repo trains) corresponds to the BFL layout.

What this pins beyond flux_blocks_golden.pt (block arithmetic): the diffusers module wiring that oracle/flux_ref.py
restates from memory — CombinedTimestepGuidanceTextProjEmbeddings (sinusoid convention, x1000, guidance, pooled text),
x_embedder / context_embedder, AdaLayerNormZero chunk order, FluxPosEmbed on [text ids; image ids], AdaLayerNormContinuous
(scale, shift) order, proj_out.

No weights are stored: the diffusers-named parameters are filled by name from tests/golden/synth.py and converted to the
BFL layout here; the test regenerates them the same way.
"""
import ast
import os
import sys
import types
from unittest.mock import MagicMock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from synth import fill_parameters, synth_tensor  # noqa: E402

COMFY = '/root/reference/submodules/ComfyUI'
REF_FLUX = '/root/reference/models/flux.py'
OUT = os.path.join(HERE, 'flux_model_golden.pt')
CFG = dict(dim=256, heads=2, num_double=1, num_single=1, in_channels=64, joint_dim=64, pooled_dim=32, B=2, h=4, w=6, Lt=10)


def reference_key_map():
    tree = ast.parse(open(REF_FLUX).read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], 'id', None) == 'BFL_TO_DIFFUSERS_MAP':
            return ast.literal_eval(node.value)
    raise RuntimeError('BFL_TO_DIFFUSERS_MAP not found')


def load_comfy_flux():
    class Stub(types.ModuleType):
        def __getattr__(self, n):
            if n.startswith('__'):
                raise AttributeError(n)
            return MagicMock()
    for n in ('comfy_aimdo', 'comfy_aimdo.host_buffer', 'comfy_aimdo.control', 'comfy_aimdo.model_vbar',
              'comfy_aimdo.torch', 'comfy_aimdo.vram_buffer', 'comfy_aimdo.model_mmap'):
        sys.modules[n] = Stub(n)
    sys.path.insert(0, COMFY)
    sys.argv = [sys.argv[0], '--cpu']
    import comfy.options
    comfy.options.enable_args_parsing()
    import comfy.ldm.flux.model as M
    import comfy.model_management
    import comfy.ops
    comfy.model_management.in_training = True
    return M, comfy.ops.disable_weight_init


def expand(key_map, n_double, n_single):
    """[(bfl_key, [diffusers keys])] with block indices filled in; the reference's `.scale` is ComfyUI's `.weight`"""
    out = []
    for bfl, dif in key_map.items():
        if '()' in bfl:
            n, prefix = (n_double, 'transformer_blocks') if bfl.startswith('double_blocks') else (n_single, 'single_transformer_blocks')
            for i in range(n):
                out.append((bfl.replace('()', str(i)), [f'{prefix}.{i}.{d}' for d in dif]))
        else:
            out.append((bfl, list(dif)))
    return [(b[:-len('.scale')] + '.weight' if b.endswith('.scale') else b, d) for b, d in out]


def swap_halves(t):
    a, b = t.chunk(2, dim=0)
    return torch.cat([b, a], dim=0)


def main():
    from oracle import flux_ref as R
    M, ops = load_comfy_flux()
    c = CFG
    ref = fill_parameters(R.RefFluxTransformer(dim=c['dim'], heads=c['heads'], num_double=c['num_double'], num_single=c['num_single'],
                                               in_channels=c['in_channels'], joint_dim=c['joint_dim'], pooled_dim=c['pooled_dim']))
    dsd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    model = M.Flux(in_channels=16, out_channels=16, vec_in_dim=c['pooled_dim'], context_in_dim=c['joint_dim'], hidden_size=c['dim'],
                   mlp_ratio=4.0, num_heads=c['heads'], depth=c['num_double'], depth_single_blocks=c['num_single'],
                   axes_dim=[16, 56, 56], theta=10000, patch_size=2, qkv_bias=True, guidance_embed=True, txt_ids_dims=[],
                   dtype=torch.float32, device='cpu', operations=ops)
    pairs = expand(reference_key_map(), c['num_double'], c['num_single'])
    bsd, used = {}, set()
    for bfl, dif in pairs:
        t = torch.cat([dsd[k] for k in dif], dim=0)
        if bfl.startswith('final_layer.adaLN_modulation.1.'):
            t = swap_halves(t)                                   # diffusers (scale, shift) -> BFL (shift, scale), :280-288
        bsd[bfl] = t
        used.update(dif)
    assert used == set(dsd), sorted(set(dsd) ^ used)[:5]
    missing, unexpected = model.load_state_dict(bsd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)

    B, h, w, Lt = c['B'], c['h'], c['w'], c['Lt']
    L = h * w
    img = synth_tensor((B, L, 64), 701, 1.0).requires_grad_(True)
    txt = synth_tensor((B, Lt, c['joint_dim']), 702, 1.0).requires_grad_(True)
    y = synth_tensor((B, c['pooled_dim']), 703, 1.0).requires_grad_(True)
    t = torch.tensor([0.25, 0.8125])
    guidance = torch.full((B,), 1.0)
    img_ids = torch.zeros(h, w, 3)
    img_ids[..., 1] += torch.arange(h)[:, None]
    img_ids[..., 2] += torch.arange(w)[None, :]
    img_ids = img_ids.reshape(1, L, 3).repeat(B, 1, 1)
    txt_ids = torch.zeros(B, Lt, 3)
    # forward only: ComfyUI's single-stream block updates its input in place (inference code), which autograd rejects; the
    # backward of the block arithmetic is pinned by flux_blocks_golden.pt (the flow repo's training blocks)
    with torch.no_grad():
        out = model.forward_orig(img, img_ids, txt, txt_ids, t, y, guidance)
        # linearity probe of the embedders: the same model on a second, unrelated timestep / guidance / pooled vector
        t2, g2 = torch.tensor([0.5, 0.0625]), torch.full((B,), 3.5)
        y2 = synth_tensor((B, c['pooled_dim']), 705, 1.0)
        out2 = model.forward_orig(img, img_ids, txt, txt_ids, t2, y2, g2)
    torch.save({'cfg': c, 'img': img.detach(), 'txt': txt.detach(), 'y': y.detach(), 't': t, 'guidance': guidance,
                'img_ids': img_ids, 'txt_ids': txt_ids, 'out': out, 't2': t2, 'guidance2': g2, 'y2': y2, 'out2': out2}, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')

if __name__ == '__main__':
    main()
