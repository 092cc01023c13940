"""Generates tests/golden/wan_golden.pt by running the REFERENCE's own Wan model code (fp32, CPU).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_wan.py

Source of truth: /root/reference/models/wan/model.py, imported as it is (WanModel, WanAttentionBlock, Head, rope_apply,
sinusoidal_embedding_1d, unpatchify).  Two stand-ins make the import possible here:
  * `diffusers.configuration_utils.{ConfigMixin, register_to_config}` and `diffusers.models.modeling_utils.ModelMixin`
    (base classes only; diffusers is not installed) — inert;
  * `flash_attention` (models/wan/attention.py asserts CUDA, :49) is replaced by softmax(q k^T / sqrt(d)) v over the
    first k_lens keys — the function flash_attn's varlen kernel computes.
The pipeline-layer glue around the model (models/wan/wan.py:414-546; the module cannot be imported: it pulls in the VAE,
T5, CLIP and DeepSpeed) is replayed here line by line on the reference's modules AND executed from its own source text
(`load_reference_layers`): both must give the identical output (asserted bit for bit) before the fixture is written.

The fixture stores no weights: parameters are filled by name from tests/golden/synth.py.  Stored: inputs, the model
output, one block's output, and gradient fingerprints of every parameter and of the inputs.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import fill_parameters, fingerprint, synth_tensor  # noqa: E402

REF = '/root/reference/models/wan'
OUT = os.path.join(HERE, 'wan_golden.pt')


def load_reference_wan():
    d = types.ModuleType('diffusers'); d.__path__ = []
    cu = types.ModuleType('diffusers.configuration_utils')
    cu.ConfigMixin = type('ConfigMixin', (), {})
    cu.register_to_config = lambda f: f
    mm = types.ModuleType('diffusers.models'); mm.__path__ = []
    mu = types.ModuleType('diffusers.models.modeling_utils')
    mu.ModelMixin = type('ModelMixin', (nn.Module,), {})
    for n, m in (('diffusers', d), ('diffusers.configuration_utils', cu), ('diffusers.models', mm),
                 ('diffusers.models.modeling_utils', mu)):
        sys.modules[n] = m
    pkg = types.ModuleType('refwan'); pkg.__path__ = [REF]
    sys.modules['refwan'] = pkg
    for name in ('attention', 'model'):
        spec = importlib.util.spec_from_file_location('refwan.' + name, os.path.join(REF, name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules['refwan.' + name] = mod
        spec.loader.exec_module(mod)
    M = sys.modules['refwan.model']

    def attention_standin(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
                          window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
        qq, kk, vv = (t.permute(0, 2, 1, 3) for t in (q, k, v))
        s = torch.matmul(qq, kk.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
        if k_lens is not None:
            mask = torch.arange(k.shape[1])[None, :] < k_lens[:, None]
            s = s.masked_fill(~mask[:, None, None, :], float('-inf'))
        return torch.matmul(torch.softmax(s, dim=-1), vv).permute(0, 2, 1, 3)
    M.flash_attention = attention_standin
    return M


def load_reference_layers(M):
    """The pipeline layers of models/wan/wan.py:414-546 (InitialLayer, TransformerLayer, FinalLayer), taken verbatim from
    the source text (the module itself imports the VAE, T5, CLIP, accelerate and DeepSpeed-dependent base classes) together
    with `make_contiguous` (models/base.py:37-38).  Stand-ins: AUTOCAST_DTYPE (autocast('cuda') is inert on the CPU, so the
    layers run in fp32) and a no-op offloader."""
    import ast
    ns = {'torch': torch, 'nn': nn, 'AUTOCAST_DTYPE': torch.bfloat16, 'sinusoidal_embedding_1d': M.sinusoidal_embedding_1d}
    tree = ast.parse(open('/root/reference/models/base.py').read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'make_contiguous']
    exec(compile(ast.Module(body=fns, type_ignores=[]), 'models/base.py', 'exec'), ns)
    tree = ast.parse(open('/root/reference/models/wan/wan.py').read())
    classes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ('InitialLayer', 'TransformerLayer', 'FinalLayer')]
    assert len(classes) == 3
    exec(compile(ast.Module(body=classes, type_ignores=[]), 'models/wan/wan.py', 'exec'), ns)
    return ns


class _NoOffload:
    def wait_for_block(self, i):
        pass

    def submit_move_blocks_forward(self, i):
        pass


def run_reference_layers(M, model, x, y, t, text, text_lens):
    """the reference's own layer stack on the tuple its prepare_inputs emits (None -> empty tensor, utils/dataset.py:1277-1279)"""
    import warnings
    L = load_reference_layers(M)
    none = torch.tensor([])
    layers = [L['InitialLayer'](model, None)] + [L['TransformerLayer'](b, i, _NoOffload()) for i, b in enumerate(model.blocks)] \
        + [L['FinalLayer'](model)]
    h = (x, none if y is None else y, t, text, text_lens, none)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')          # "CUDA is not available" from the autocast decorators
        for layer in layers:
            h = layer(h)
    return h


def main():
    M = load_reference_wan()
    cfg = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16, B=2, f=3, h=8, w=12)
    model = M.WanModel(model_type='t2v', dim=cfg['dim'], ffn_dim=cfg['ffn_dim'], num_heads=cfg['num_heads'],
                       num_layers=cfg['num_layers'], text_dim=cfg['text_dim'], text_len=cfg['text_len'])
    fill_parameters(model)
    B, f, h, w = cfg['B'], cfg['f'], cfg['h'], cfg['w']
    x = synth_tensor((B, 16, f, h, w), 401, 1.0).requires_grad_(True)
    text = synth_tensor((B, cfg['text_len'], cfg['text_dim']), 402, 1.0).requires_grad_(True)
    text_lens = torch.tensor([10, 16])
    t = torch.tensor([250.0, 812.5])

    # ---- models/wan/wan.py:432-511 InitialLayer.forward (t2v, cached text embeddings) ----
    context = [emb[:length] for emb, length in zip(text, text_lens)]
    xs = [model.patch_embedding(u.unsqueeze(0)) for u in x]
    grid_sizes = torch.stack([torch.tensor(u.shape[2:], dtype=torch.long) for u in xs])
    xs = [u.flatten(2).transpose(1, 2) for u in xs]
    seq_lens = torch.tensor([u.size(1) for u in xs], dtype=torch.long)
    seq_len = seq_lens.max()
    xe = torch.cat([torch.cat([u, u.new_zeros(1, seq_len - u.size(1), u.size(2))], dim=1) for u in xs])
    tt = t.unsqueeze(-1)
    bt = tt.size(0)
    e = model.time_embedding(M.sinusoidal_embedding_1d(model.freq_dim, tt.flatten()).unflatten(0, (bt, 1)).to(torch.float32))
    e0 = model.time_projection(e).unflatten(2, (6, model.dim))
    ctx = model.text_embedding(torch.stack([torch.cat([u, u.new_zeros(model.text_len - u.size(0), u.size(1))]) for u in context]))
    # ---- :521-529 TransformerLayer.forward ----
    hcur = xe
    for blk in model.blocks:
        hcur = blk(hcur, e0, seq_lens, grid_sizes, model.freqs, ctx, None)
    # ---- :541-546 FinalLayer.forward ----
    out = torch.stack(model.unpatchify(model.head(hcur, e), grid_sizes), dim=0)
    # the same through the reference's own InitialLayer / TransformerLayer / FinalLayer text: must be the replay above
    out_layers = run_reference_layers(M, model, x.detach().clone(), None, t, text.detach().clone(), text_lens)
    assert torch.equal(out_layers, out), (out_layers - out).abs().max()
    probe = synth_tensor(tuple(out.shape), 403, 1.0)
    (out * probe).sum().backward()
    g = {'cfg': cfg, 'x': x.detach(), 'text': text.detach(), 'text_lens': text_lens, 't': t, 'out': out.detach(),
         'probe': probe, 'dx': x.grad.clone(), 'dtext': text.grad.clone(),
         'param_grads': {n: fingerprint(p.grad, 500 + i) for i, (n, p) in enumerate(model.named_parameters())}}

    # ---- one block on its own ----
    blk = model.blocks[0]
    model.zero_grad()
    L = f * (h // 2) * (w // 2)
    bx = synth_tensor((B, L, cfg['dim']), 404, 1.0).requires_grad_(True)
    be0 = synth_tensor((B, 1, 6, cfg['dim']), 405, 0.3).requires_grad_(True)
    bctx = synth_tensor((B, cfg['text_len'], cfg['dim']), 406, 1.0).requires_grad_(True)
    by = blk(bx, be0, seq_lens, grid_sizes, model.freqs, bctx, None)
    bp = synth_tensor(tuple(by.shape), 407, 1.0)
    (by * bp).sum().backward()
    # the rope multipliers the reference applies for this grid, as real tables (for the oracle's table test)
    c = (cfg['dim'] // cfg['num_heads']) // 2
    fr = model.freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    hh, ww = h // 2, w // 2
    freqs_i = torch.cat([fr[0][:f].view(f, 1, 1, -1).expand(f, hh, ww, -1), fr[1][:hh].view(1, hh, 1, -1).expand(f, hh, ww, -1),
                         fr[2][:ww].view(1, 1, ww, -1).expand(f, hh, ww, -1)], dim=-1).reshape(L, -1)
    g['block'] = {'x': bx.detach(), 'e0': be0.detach(), 'ctx': bctx.detach(), 'y': by.detach(), 'probe': bp,
                  'dx': bx.grad.clone(), 'de0': be0.grad.clone(), 'dctx': bctx.grad.clone(),
                  'rope_cos': freqs_i.real.clone(), 'rope_sin': freqs_i.imag.clone(),
                  'param_grads': {n: fingerprint(p.grad, 600 + i) for i, (n, p) in enumerate(blk.named_parameters())}}
    # ---- Wan2.2 I2V ('i2v_v2'): the same model with 36 input channels = [x | first-frame mask | y] (wan.py:459-465) ----
    m2 = M.WanModel(model_type='i2v_v2', in_dim=36, dim=cfg['dim'], ffn_dim=cfg['ffn_dim'], num_heads=cfg['num_heads'],
                    num_layers=1, text_dim=cfg['text_dim'], text_len=cfg['text_len'])
    fill_parameters(m2)
    x2 = synth_tensor((B, 16, f, h, w), 411, 1.0).requires_grad_(True)
    y2 = synth_tensor((B, 16, f, h, w), 412, 1.0).requires_grad_(True)
    fmask = torch.zeros((B, 4, f, h, w))
    fmask[:, :, 0, ...] = 1
    yy = torch.cat([fmask, y2], dim=1)
    xs = [torch.cat([u, v], dim=0) for u, v in zip(x2, yy)]
    xs = [m2.patch_embedding(u.unsqueeze(0)) for u in xs]
    gs = torch.stack([torch.tensor(u.shape[2:], dtype=torch.long) for u in xs])
    xs = [u.flatten(2).transpose(1, 2) for u in xs]
    sl = torch.tensor([u.size(1) for u in xs], dtype=torch.long)
    xe2 = torch.cat(xs)
    e_2 = m2.time_embedding(M.sinusoidal_embedding_1d(m2.freq_dim, t).unflatten(0, (B, 1)).to(torch.float32))
    e0_2 = m2.time_projection(e_2).unflatten(2, (6, m2.dim))
    ctx2 = m2.text_embedding(torch.stack([torch.cat([u, u.new_zeros(m2.text_len - u.size(0), u.size(1))])
                                          for u in [emb[:n] for emb, n in zip(text.detach(), text_lens)]]))
    h2 = m2.blocks[0](xe2, e0_2, sl, gs, m2.freqs, ctx2, None)
    out2 = torch.stack(m2.unpatchify(m2.head(h2, e_2), gs), dim=0)
    out2_layers = run_reference_layers(M, m2, x2.detach().clone(), y2.detach().clone(), t, text.detach().clone(), text_lens)
    assert torch.equal(out2_layers, out2), (out2_layers - out2).abs().max()
    probe2 = synth_tensor(tuple(out2.shape), 413, 1.0)
    (out2 * probe2).sum().backward()
    g['i2v_v2'] = {'x': x2.detach(), 'y': y2.detach(), 'out': out2.detach(), 'probe': probe2, 'dx': x2.grad.clone(),
                   'dy': y2.grad.clone(),
                   'param_grads': {n: fingerprint(p.grad, 900 + i) for i, (n, p) in enumerate(m2.named_parameters())}}
    torch.save(g, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
