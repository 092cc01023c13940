"""CPU: host-side logic of the Qwen-Image plugin (diffusion-pipe_b200/qwen_image.py) against the oracle
(oracle/qwen_ref.py, itself pinned to the reference tree's in-tree model by tests/test_oracle_qwen_golden.py):
parameter names, rope tables, prepare_inputs, lazy LayerSpecs.  No kernel is launched here."""
import torch

from diffusion_pipe_b200 import qwen_image as P
from oracle import qwen_ref as Q

CFG = {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}


def test_parameter_names_match_diffusers_layout():
    m = P.QwenImageTransformer2DModel(CFG, device='cpu')
    ref = Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64)
    mine = {n: tuple(p.shape) for n, p in m.named_parameters()}
    theirs = {n: tuple(p.shape) for n, p in ref.named_parameters()}
    assert mine == theirs
    assert all(p.original_name == n for n, p in m.named_parameters())
    # the aliases the fused block function reads do not register parameters twice
    blk = m.transformer_blocks[0]
    assert blk.norm1.linear is blk.img_mod[1] and blk.ff is blk.img_mlp
    assert len(list(blk.parameters())) == len({id(p) for p in blk.parameters()}) == 32


def test_rope_tables_match_oracle():
    for shapes, lt in (([(1, 4, 6)], 10), ([(1, 8, 8), (1, 8, 8)], 7), ([(1, 5, 3)], 3)):
        vid, txt = P.qwen_rope_tables(shapes, lt)
        vc, vs, tc, ts = Q.qwen_rope_tables(shapes, lt)
        torch.testing.assert_close(vid, torch.stack([vc, vs]), rtol=0, atol=0)
        torch.testing.assert_close(txt, torch.stack([tc, ts]), rtol=0, atol=0)
        assert vid.dtype == torch.float32 and vid.shape == (2, sum(f * h * w for f, h, w in shapes), 128)


def test_prepare_inputs_matches_oracle_on_the_same_draws():
    pipe = P.QwenImagePipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': CFG}}, device='cpu')
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(2, 16, 1, 8, 12, generator=g)
    pe = [torch.randn(5, 64, generator=g), torch.randn(9, 64, generator=g)]
    ctrl = torch.randn(2, 16, 1, 8, 12, generator=g)
    for control in (None, ctrl):
        batch = {'latents': lat, 'prompt_embeds': pe, 'mask': None}
        if control is not None:
            batch['control_latents'] = control
        torch.manual_seed(11)
        feats, (target, mask) = pipe.prepare_inputs(batch)
        torch.manual_seed(11)
        t = torch.sigmoid(torch.distributions.normal.Normal(0, 1).sample((2,)))
        noise_packed = torch.randn(2, 24, 64)            # the reference draws x_0 in the packed layout (:450)
        rf, (rt, rm) = Q.prepare_inputs(lat, pe, t, torch.zeros_like(lat), control_latents=control)
        x1 = Q.pack_latents(lat)
        te = t.view(-1, 1, 1)
        x_t = (1 - te) * x1 + te * noise_packed
        if control is not None:
            x_t = torch.cat([x_t, Q.pack_latents(control)], dim=1)
        torch.testing.assert_close(feats[0], x_t)
        torch.testing.assert_close(target, noise_packed - x1)
        assert mask is None and rm is None
        for a, b in zip(feats[1:], rf[1:]):
            assert a.dtype == b.dtype and a.shape == b.shape
            if a.dtype != torch.float32 or a.ndim != 1:
                assert torch.equal(a, b)
        torch.testing.assert_close(feats[3], t)
        assert len(feats) == (7 if control is not None else 6)
    # eval quantile path (train.py:176-242): deterministic t
    feats, _ = pipe.prepare_inputs({'latents': lat, 'prompt_embeds': pe, 'mask': None}, timestep_quantile=0.5)
    torch.testing.assert_close(feats[3], torch.full((2,), 0.5))


def test_lazy_layer_specs_cover_the_model():
    pipe = P.QwenImagePipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': CFG}}, device='cpu')
    specs = pipe.to_layers()
    assert [s.typename.__name__ for s in specs] == ['InitialLayer', 'TransformerLayer', 'TransformerLayer', 'FinalLayer']
    full = P.QwenImageTransformer2DModel(CFG, device='cpu')
    assert sum(s.param_count for s in specs) == sum(p.numel() for p in full.parameters())
    names = set()
    for s in specs:
        names |= {p.original_name for p in s.build().parameters()}
    assert names == {n for n, _ in full.named_parameters()}


def test_ragged_prompts_produce_a_key_mask_with_trailing_padding():
    pipe = P.QwenImagePipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': CFG}}, device='cpu')
    lat = torch.randn(2, 16, 1, 8, 8)
    feats, _ = pipe.prepare_inputs({'latents': lat, 'prompt_embeds': [torch.randn(3, 64), torch.randn(6, 64)], 'mask': None})
    assert not bool(feats[2].all())
