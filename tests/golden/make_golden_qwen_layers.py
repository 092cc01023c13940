"""Generates tests/golden/qwen_layers_golden.pt: the REFERENCE's own Qwen-Image pipeline layers over the oracle's modules.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_qwen_layers.py

`InitialLayer`, `TransformerLayer`, `FinalLayer` (models/qwen_image.py:519-605) and `make_contiguous` (models/base.py:37-38)
are taken from the source text (ast).  They are thin glue around diffusers modules, which are absent; the modules they call
are therefore the ORACLE's (oracle/qwen_ref.py, each already pinned: qwen_golden.pt, qwen_attn_golden.pt) behind
adapters with diffusers' call signatures:

    model.time_text_embed(timestep, hidden_states)          -> oracle time_text_embed(timestep)
    model.pos_embed(img_shapes, txt_seq_lens, device=...)   -> oracle rope tables as complex `freqs_cis`
    block(hidden_states=, encoder_hidden_states=, encoder_hidden_states_mask=None, temb=, image_rotary_emb=(vid, txt),
          joint_attention_kwargs={'attention_mask': m})     -> oracle block on the real cos / sin form of the same tables

What this pins is the glue of oracle/qwen_ref.py's RefInitialLayer / RefTransformerLayer / RefFinalLayer: tuple order,
which tensor goes where, the (encoder, hidden) return order of the block, the `extra` element and the
`output[:, :img_seq_len]` cut of the control-latents case, None / empty placeholders.  Inputs are the features the
reference's own prepare_inputs produced (tests/golden/host_golden.pt, cases 'ragged_prompts' and 'control').
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
from oracle import qwen_ref as Q  # noqa: E402

REF = '/root/reference'
OUT = os.path.join(HERE, 'qwen_layers_golden.pt')


def load_reference_layers():
    ns = {'torch': torch, 'nn': nn, 'AUTOCAST_DTYPE': torch.bfloat16}
    tree = ast.parse(open(f'{REF}/models/base.py').read())
    exec(compile(ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'make_contiguous'],
                            type_ignores=[]), 'models/base.py', 'exec'), ns)
    tree = ast.parse(open(f'{REF}/models/qwen_image.py').read())
    classes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ('InitialLayer', 'TransformerLayer', 'FinalLayer')]
    assert len(classes) == 3
    exec(compile(ast.Module(body=classes, type_ignores=[]), 'models/qwen_image.py', 'exec'), ns)
    return ns


class DiffusersFacade(nn.Module):
    """the attributes InitialLayer / FinalLayer read of a diffusers QwenImageTransformer2DModel, served by the oracle"""

    def __init__(self, t):
        super().__init__()
        self.t = t
        self.img_in, self.txt_norm, self.txt_in, self.norm_out, self.proj_out = t.img_in, t.txt_norm, t.txt_in, t.norm_out, t.proj_out

    def time_text_embed(self, timestep, hidden_states):
        return self.t.time_text_embed(timestep)

    def pos_embed(self, img_shapes, txt_seq_lens, device=None):
        vc, vs, tc, ts = Q.qwen_rope_tables([tuple(s) for s in img_shapes[0]], max(txt_seq_lens), self.t.axes_dim)
        cx = lambda c, s: torch.complex(c[:, 0::2].contiguous(), s[:, 0::2].contiguous())
        return cx(vc, vs), cx(tc, ts)


class BlockFacade(nn.Module):
    def __init__(self, block):
        super().__init__()
        self.block = block

    def forward(self, hidden_states, encoder_hidden_states, encoder_hidden_states_mask, temb, image_rotary_emb, joint_attention_kwargs):
        assert encoder_hidden_states_mask is None
        vid, txt = image_rotary_emb
        rep = lambda z: (z.real.repeat_interleave(2, dim=1), z.imag.repeat_interleave(2, dim=1))
        return self.block(hidden_states, encoder_hidden_states, temb, (rep(vid), rep(txt)), joint_attention_kwargs['attention_mask'])


class NoOffload:
    def wait_for_block(self, i):
        pass

    def submit_move_blocks_forward(self, i):
        pass


def main():
    L = load_reference_layers()
    host = torch.load(os.path.join(HERE, 'host_golden.pt'), weights_only=False)
    t = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=24))
    facade = DiffusersFacade(t)
    layers = [L['InitialLayer'](facade)] + [L['TransformerLayer'](BlockFacade(b), i, NoOffload()) for i, b in enumerate(t.transformer_blocks)] \
        + [L['FinalLayer'](facade)]
    g = {'cases': {}}
    for c in host['cases']:
        if c['family'] != 'qwen_image' or c['name'] not in ('ragged_prompts', 'control'):
            continue
        feats = tuple(f.clone() for f in c['features'])
        h = feats
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for layer in layers:
                h = layer(h)
        g['cases'][c['name']] = {'out': h.detach(), 'n_features': len(feats)}
        print(c['name'], tuple(h.shape))
    torch.save(g, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
