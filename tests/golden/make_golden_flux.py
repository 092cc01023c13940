"""Generates tests/golden/flux_blocks_golden.pt by running the REFERENCE's in-tree Flux-architecture blocks.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_flux.py

Source of truth: /root/reference/submodules/flow/src/models/chroma/{math.py, module/layers.py}
(DoubleStreamBlock :256-392, SingleStreamBlock :395-466, EmbedND :12-26, QKNorm, RMSNorm) — imported as they are.
Those blocks take the modulation (shift, scale, gate) as an input; the fixture stores the modulation tensors so that
the oracle (oracle/flux_ref.py) is checked on exactly the same block arithmetic, fp32, forward and backward.
Weights are stored under diffusers names using the layout map of models/flux.py:22-76.
"""
import importlib.util
import os
import sys
import types

import torch

FLOW = '/root/reference/submodules/flow/src/models/chroma'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'flux_blocks_golden.pt')


def load_flow():
    pkg = types.ModuleType('flowchroma')
    pkg.__path__ = [FLOW]
    sys.modules['flowchroma'] = pkg
    sub = types.ModuleType('flowchroma.module')
    sub.__path__ = [os.path.join(FLOW, 'module')]
    sys.modules['flowchroma.module'] = sub
    for name, path in (('flowchroma.math', os.path.join(FLOW, 'math.py')),
                       ('flowchroma.module.layers', os.path.join(FLOW, 'module', 'layers.py'))):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return sys.modules['flowchroma.module.layers']


def main():
    L = load_flow()
    torch.manual_seed(1234)
    dim, heads, Lt, Li, B = 256, 2, 24, 40, 2
    mlp_ratio = 2.0
    axes = (16, 56, 56)
    ids = torch.zeros(B, Lt + Li, 3)
    ids[:, Lt:, 1] = torch.arange(Li)[None] // 8
    ids[:, Lt:, 2] = torch.arange(Li)[None] % 8
    pe = L.EmbedND(dim=dim // heads, theta=10000, axes_dim=list(axes))(ids)
    out = {'dim': dim, 'heads': heads, 'Lt': Lt, 'Li': Li, 'B': B, 'mlp_ratio': int(mlp_ratio), 'ids': ids[0].clone()}

    def mods(n):
        return [r16(0.3 * torch.randn(B, 1, dim)) for _ in range(n)]

    def rand_init(m):
        for p in m.parameters():
            if p.ndim > 1:
                torch.nn.init.normal_(p, std=0.05)
            else:
                torch.nn.init.normal_(p, mean=1.0 if 'scale' in [n for n, q in m.named_parameters() if q is p][0] else 0.0, std=0.1)
        for p in m.parameters():
            p.data = p.data.to(torch.bfloat16).float()   # bf16-representable so the same fixture feeds the bf16 kernels

    def r16(t):
        return t.to(torch.bfloat16).float()

    def gsum(named):
        # full gradients for vectors, (sum, l2) fingerprints for matrices (keeps the fixture small)
        out = {}
        for k, p in named:
            g = p.grad.detach()
            out[k] = g.clone() if g.ndim == 1 else torch.stack([g.sum(), g.norm(), g[0].sum(), g[:, 0].sum()])
        return out

    # ---------------- double block ----------------
    blk = L.DoubleStreamBlock(dim, heads, mlp_ratio=mlp_ratio, qkv_bias=True)
    rand_init(blk)
    img = r16(torch.randn(B, Li, dim)).requires_grad_(True)
    txt = r16(torch.randn(B, Lt, dim)).requires_grad_(True)
    im, tm = mods(6), mods(6)
    for t in im + tm:
        t.requires_grad_(True)
    vec = ((L.ModulationOut(*im[:3]), L.ModulationOut(*im[3:])), (L.ModulationOut(*tm[:3]), L.ModulationOut(*tm[3:])))
    oi, ot = blk(img=img, txt=txt, pe=pe, distill_vec=vec, mask=None)
    gi, gt = r16(torch.randn_like(oi)), r16(torch.randn_like(ot))
    (oi * gi).sum().add((ot * gt).sum()).backward()
    sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}

    def split3(w):
        return w.chunk(3, dim=0)
    d = {}
    for s, a, n in (('img', 'attn.to_', ''), ('txt', 'attn.add_', '_proj')):
        for j, nm in enumerate('qkv'):
            d[f'{a}{nm}{n}.weight'] = split3(sd[f'{s}_attn.qkv.weight'])[j]
            d[f'{a}{nm}{n}.bias'] = split3(sd[f'{s}_attn.qkv.bias'])[j]
    d['attn.norm_q.weight'] = sd['img_attn.norm.query_norm.scale']
    d['attn.norm_k.weight'] = sd['img_attn.norm.key_norm.scale']
    d['attn.norm_added_q.weight'] = sd['txt_attn.norm.query_norm.scale']
    d['attn.norm_added_k.weight'] = sd['txt_attn.norm.key_norm.scale']
    d['attn.to_out.0.weight'], d['attn.to_out.0.bias'] = sd['img_attn.proj.weight'], sd['img_attn.proj.bias']
    d['attn.to_add_out.weight'], d['attn.to_add_out.bias'] = sd['txt_attn.proj.weight'], sd['txt_attn.proj.bias']
    for s, f in (('img', 'ff'), ('txt', 'ff_context')):
        d[f'{f}.net.0.proj.weight'], d[f'{f}.net.0.proj.bias'] = sd[f'{s}_mlp.0.weight'], sd[f'{s}_mlp.0.bias']
        d[f'{f}.net.2.weight'], d[f'{f}.net.2.bias'] = sd[f'{s}_mlp.2.weight'], sd[f'{s}_mlp.2.bias']
    grads = gsum(blk.named_parameters())
    out['double'] = {
        'weights': d, 'img': img.detach(), 'txt': txt.detach(),
        'img_mod': [t.detach().squeeze(1) for t in im], 'txt_mod': [t.detach().squeeze(1) for t in tm],   # shift,scale,gate x2
        'out_img': oi.detach(), 'out_txt': ot.detach(), 'g_img': gi, 'g_txt': gt,
        'd_img': img.grad.clone(), 'd_txt': txt.grad.clone(),
        'd_img_mod': [t.grad.squeeze(1).clone() for t in im], 'd_txt_mod': [t.grad.squeeze(1).clone() for t in tm],
        'flow_param_grads': grads,
    }

    # ---------------- single block ----------------
    sb = L.SingleStreamBlock(dim, heads, mlp_ratio=mlp_ratio)
    rand_init(sb)
    x = r16(torch.randn(B, Lt + Li, dim)).requires_grad_(True)
    sm = mods(3)
    for t in sm:
        t.requires_grad_(True)
    y = sb(x, pe=pe, distill_vec=L.ModulationOut(*sm), mask=None)
    gy = r16(torch.randn_like(y))
    (y * gy).sum().backward()
    sd = {k: v.detach().clone() for k, v in sb.state_dict().items()}
    w1, b1 = sd['linear1.weight'], sd['linear1.bias']
    d = {
        'attn.to_q.weight': w1[0:dim], 'attn.to_k.weight': w1[dim:2 * dim], 'attn.to_v.weight': w1[2 * dim:3 * dim],
        'proj_mlp.weight': w1[3 * dim:],
        'attn.to_q.bias': b1[0:dim], 'attn.to_k.bias': b1[dim:2 * dim], 'attn.to_v.bias': b1[2 * dim:3 * dim],
        'proj_mlp.bias': b1[3 * dim:],
        'attn.norm_q.weight': sd['norm.query_norm.scale'], 'attn.norm_k.weight': sd['norm.key_norm.scale'],
        'proj_out.weight': sd['linear2.weight'], 'proj_out.bias': sd['linear2.bias'],
    }
    out['single'] = {
        'weights': d, 'x': x.detach(), 'mod': [t.detach().squeeze(1) for t in sm],
        'out': y.detach(), 'g': gy, 'd_x': x.grad.clone(), 'd_mod': [t.grad.squeeze(1).clone() for t in sm],
        'flow_param_grads': gsum(sb.named_parameters()),
    }
    for blkname in ('double', 'single'):
        w = out[blkname]['weights']
        for k in w:
            w[k] = w[k].to(torch.bfloat16)   # exact (values are bf16-representable)
    torch.save(out, OUT)
    print('wrote', OUT, os.path.getsize(OUT) // 1024, 'KiB')


if __name__ == '__main__':
    main()
