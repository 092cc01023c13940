"""Generates tests/golden/flux_layers_golden.pt: the REFERENCE's own Flux pipeline wrappers over the oracle's modules.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_flux_layers.py

`EmbeddingWrapper`, `TransformerWrapper`, `SingleTransformerWrapper`, `OutputWrapper` (models/flux.py:456-548) and
`make_contiguous` (models/base.py:37-38) are taken from the source text (ast) and run over the ORACLE's modules
(oracle/flux_ref.py; pinned by flux_blocks_golden.pt / flux_model_golden.pt) — diffusers, whose modules they wrap, is
absent.  The oracle's modules already have diffusers' call signatures; two adapters: `pos_embed(ids)` -> the oracle's rope
tables, and the time/text embedder wears diffusers' class name, which the wrapper dispatches on (dev vs schnell,
models/flux.py:475-479).

Pins the glue of oracle/flux_ref.py's RefEmbeddingWrapper / RefTransformerWrapper / RefOutputWrapper: tuple order, x1000 on
timestep and guidance, [text ids; image ids] for the rope, (encoder, hidden) return order, and the `[:, :img_seq_len]` cut
that drops the Kontext control tokens.  Inputs: the features the reference's own prepare_inputs produced
(tests/golden/host_golden.pt: 'default', 'kontext'), widened to the oracle model's channel counts.
"""
import ast
import os
import sys
import warnings

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from synth import fill_parameters  # noqa: E402
from oracle import flux_ref as R  # noqa: E402

REF = '/root/reference'
OUT = os.path.join(HERE, 'flux_layers_golden.pt')
DIMS = dict(dim=256, heads=2, num_double=1, num_single=1, joint_dim=32, pooled_dim=16)      # t5 / clip widths of host_golden.pt


def load_reference_wrappers():
    ns = {'torch': torch, 'nn': nn, 'AUTOCAST_DTYPE': torch.bfloat16}
    tree = ast.parse(open(f'{REF}/models/base.py').read())
    exec(compile(ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'make_contiguous'],
                            type_ignores=[]), 'models/base.py', 'exec'), ns)
    tree = ast.parse(open(f'{REF}/models/flux.py').read())
    names = ('EmbeddingWrapper', 'TransformerWrapper', 'SingleTransformerWrapper', 'OutputWrapper')
    classes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in names]
    assert len(classes) == 4
    exec(compile(ast.Module(body=classes, type_ignores=[]), 'models/flux.py', 'exec'), ns)
    return ns


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):      # the class NAME is what models/flux.py:476 looks at
    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, timestep, guidance, pooled):
        return self.inner(timestep, guidance, pooled)


class NoOffload:
    def wait_for_block(self, i):
        pass

    def submit_move_blocks_forward(self, i):
        pass


def build(t):
    W = load_reference_wrappers()
    pos_embed = lambda ids: R.flux_rope_tables(ids, t.axes_dim)
    layers = [W['EmbeddingWrapper'](t.x_embedder, CombinedTimestepGuidanceTextProjEmbeddings(t.time_text_embed), t.context_embedder, pos_embed)]
    layers += [W['TransformerWrapper'](b, i, NoOffload()) for i, b in enumerate(t.transformer_blocks)]
    layers += [W['SingleTransformerWrapper'](b, i, NoOffload()) for i, b in enumerate(t.single_transformer_blocks)]
    layers.append(W['OutputWrapper'](t.norm_out, t.proj_out))
    return layers


def main():
    host = torch.load(os.path.join(HERE, 'host_golden.pt'), weights_only=False)
    t = fill_parameters(R.RefFluxTransformer(**DIMS))
    layers = build(t)
    g = {'dims': DIMS, 'cases': {}}
    for c in host['cases']:
        if c['family'] != 'flux' or c['name'] not in ('default', 'kontext'):
            continue
        h = tuple(f.clone() for f in c['features'])
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for layer in layers:
                h = layer(h)
        g['cases'][c['name']] = {'out': h.detach()}
        print(c['name'], tuple(h.shape))
    torch.save(g, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
