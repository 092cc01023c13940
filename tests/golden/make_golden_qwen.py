"""Generates tests/golden/qwen_golden.pt by running the REFERENCE tree's in-tree Qwen-Image transformer (fp32, CPU).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_qwen.py

Source of truth: /root/reference/submodules/ComfyUI/comfy/ldm/qwen_image/model.py (QwenImageTransformer2DModel,
QwenImageTransformerBlock, Attention, LastLayer, QwenTimestepProjEmbeddings) — imported as it is.  ComfyUI's runtime
(`comfy_aimdo`, a CUDA memory manager) is absent here, so those imports are replaced by inert stubs; nothing on the
arithmetic path touches them.  `comfy.model_management.in_training = True` selects the plain PyTorch RoPE.

The model is the same architecture the reference trains through diffusers (models/qwen_image.py:272-285) with the same
parameter names, so the fixture stores no weights: both this script and the tests fill parameters by name from the
deterministic generator in tests/golden/synth.py.  Stored: inputs, the whole-model output, one block's outputs, and
gradient fingerprints (sum, abs-sum, a fixed-probe dot product) of every parameter and of the inputs.
"""
import os
import sys
import types
from unittest.mock import MagicMock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import fill_parameters, fingerprint, synth_tensor  # noqa: E402

COMFY = '/root/reference/submodules/ComfyUI'
OUT = os.path.join(HERE, 'qwen_golden.pt')


def load_comfy_qwen():
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
    import comfy.ldm.qwen_image.model as M
    import comfy.model_management
    import comfy.ops
    comfy.model_management.in_training = True
    return M, comfy.ops.disable_weight_init


def main():
    M, ops = load_comfy_qwen()
    cfg = dict(dim=256, heads=2, num_layers=2, joint_dim=64, in_channels=64, out_channels=16, B=2, h=8, w=12, Lt=10)
    model = M.QwenImageTransformer2DModel(patch_size=2, in_channels=cfg['in_channels'], out_channels=cfg['out_channels'],
                                          num_layers=cfg['num_layers'], attention_head_dim=128,
                                          num_attention_heads=cfg['heads'], joint_attention_dim=cfg['joint_dim'],
                                          dtype=torch.float32, device='cpu', operations=ops)
    fill_parameters(model)
    B, h, w, Lt = cfg['B'], cfg['h'], cfg['w'], cfg['Lt']
    x = synth_tensor((B, 16, 1, h, w), 101, 1.0).requires_grad_(True)            # noised latents
    ctx = synth_tensor((B, Lt, cfg['joint_dim']), 102, 1.0).requires_grad_(True)  # prompt embeddings (no padding)
    t = torch.tensor([0.25, 0.8125])                                               # bf16-representable timesteps
    out = model._forward(x, t, ctx)                                                # [B, 16, 1, h, w]
    probe = synth_tensor(tuple(out.shape), 103, 1.0)
    (out * probe).sum().backward()
    g = {'cfg': cfg, 'x': x.detach(), 'ctx': ctx.detach(), 't': t, 'out': out.detach(), 'probe': probe,
         'dx': x.grad.clone(), 'dctx': ctx.grad.clone(),
         'param_grads': {n: fingerprint(p.grad, 200 + i) for i, (n, p) in enumerate(model.named_parameters())}}

    # one block on its own (inputs after the embedders), incl. the rope table ComfyUI builds for these positions
    blk = model.transformer_blocks[0]
    Li = (h // 2) * (w // 2)
    hid = synth_tensor((B, Li, cfg['dim']), 104, 1.0).requires_grad_(True)
    enc = synth_tensor((B, Lt, cfg['dim']), 105, 1.0).requires_grad_(True)
    temb = synth_tensor((B, cfg['dim']), 106, 1.0).requires_grad_(True)
    _, img_ids, _ = model.process_img(x.detach())
    txt_start = max((w // 2) // 2, (h // 2) // 2)
    txt_ids = torch.arange(txt_start, txt_start + Lt).reshape(1, -1, 1).repeat(B, 1, 3)
    pe = model.pe_embedder(torch.cat((txt_ids, img_ids), dim=1)).contiguous()     # [B, 1, L, 64, 2, 2]
    model.zero_grad()
    eo, ho = blk(hid, enc, None, temb, pe)
    p1 = synth_tensor(tuple(ho.shape), 107, 1.0)
    p2 = synth_tensor(tuple(eo.shape), 108, 1.0)
    ((ho * p1).sum() + (eo * p2).sum()).backward()
    g['block'] = {'hid': hid.detach(), 'enc': enc.detach(), 'temb': temb.detach(), 'hid_out': ho.detach(),
                  'enc_out': eo.detach(), 'p_hid': p1, 'p_enc': p2, 'dhid': hid.grad.clone(), 'denc': enc.grad.clone(),
                  'dtemb': temb.grad.clone(), 'rope_cos': pe[0, 0, :, :, 0, 0].clone(), 'rope_sin': pe[0, 0, :, :, 1, 0].clone(),
                  'param_grads': {n: fingerprint(p.grad, 300 + i) for i, (n, p) in enumerate(blk.named_parameters())}}
    torch.save(g, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
