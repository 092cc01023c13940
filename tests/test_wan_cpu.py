"""CPU: host-side logic of the Wan plugin (diffusion-pipe_b200/wan.py) against the oracle (oracle/wan_ref.py, itself
pinned to the reference's own model code by tests/test_oracle_wan_golden.py): parameter names and shapes, rope tables,
the timestep table, prepare_inputs, unpatchify, lazy LayerSpecs.  No kernel is launched here."""
import pytest
import torch

from diffusion_pipe_b200 import wan as P
from oracle import wan_ref as W

CFG = {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 2, 'text_dim': 64, 'text_len': 16}


def test_parameter_names_and_shapes_match_reference_layout():
    m = P.WanModel(CFG, device='cpu')
    ref = W.RefWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16)
    assert {n: tuple(p.shape) for n, p in m.named_parameters()} == {n: tuple(p.shape) for n, p in ref.named_parameters()}
    assert all(p.original_name == n for n, p in m.named_parameters())
    blk = m.blocks[0]
    assert len(list(blk.parameters())) == len({id(p) for p in blk.parameters()}) == 27
    # q/k/v of self-attention and k/v of cross-attention alias one fused allocation each
    assert blk.self_attn.k.weight.data_ptr() == blk.self_attn.qkv.weight.data_ptr() + 256 * 256 * 2
    assert blk.cross_attn.v.weight.data_ptr() == blk.cross_attn.kv.weight.data_ptr() + 256 * 256 * 2


def test_rope_tables_and_sinusoid_match_oracle():
    for grid in ((3, 4, 6), (1, 8, 8), (5, 3, 7)):
        torch.testing.assert_close(P.wan_rope_tables(grid), torch.stack(W.wan_rope_tables(grid)), rtol=0, atol=0)
    t = torch.tensor([0.0, 250.0, 999.0])
    torch.testing.assert_close(P.sinusoidal_embedding_1d(256, t), W.sinusoidal_embedding_1d(256, t), rtol=0, atol=0)


def test_t_table_and_prepare_inputs_match_oracle_on_the_same_draws():
    pipe = P.WanPipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': CFG}}, device='cpu')
    torch.testing.assert_close(pipe.t_dist, W.t_distribution(), rtol=0, atol=0)
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(2, 16, 3, 8, 12, generator=g)
    text = torch.randn(2, 16, 64, generator=g)
    lens = torch.tensor([10, 16])
    torch.manual_seed(21)
    feats, (target, mask) = pipe.prepare_inputs({'latents': lat, 'mask': None, 'text_embeddings': text, 'seq_lens': lens})
    torch.manual_seed(21)
    t = pipe.t_dist[torch.randint(0, 10000, size=(2,))]
    noise = torch.randn_like(lat)
    rf, (rt, rm) = W.prepare_inputs(lat, text, lens, t, noise)
    torch.testing.assert_close(feats[0], rf[0])
    torch.testing.assert_close(feats[2], rf[2])
    torch.testing.assert_close(target, rt)
    assert feats[1] is None and feats[5] is None and mask is None       # split_batch turns these into empty tensors
    # eval quantile (train.py:176-242): t = table[int(q * len)]
    feats, _ = pipe.prepare_inputs({'latents': lat, 'mask': None, 'text_embeddings': text, 'seq_lens': lens}, timestep_quantile=0.5)
    torch.testing.assert_close(feats[2], pipe.t_dist[5000].repeat(2) * 1000)
    # shift and min_t / max_t slicing
    pipe2 = P.WanPipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': CFG, 'shift': 3.0,
                                     'min_t': 0.2, 'max_t': 0.9}}, device='cpu')
    feats, _ = pipe2.prepare_inputs({'latents': lat, 'mask': None, 'text_embeddings': text, 'seq_lens': lens}, timestep_quantile=0.0)
    assert 200.0 <= float(feats[2][0]) < 201.0


def test_unpatchify_matches_oracle():
    ref = W.RefWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=64, text_len=16)
    x = torch.randn(2, 3 * 4 * 6, 64)
    grid = torch.tensor([[3, 4, 6]] * 2)
    want = torch.stack(ref.unpatchify(x, grid))
    torch.testing.assert_close(P.unpatchify(x, (3, 4, 6), (1, 2, 2), 16), want, rtol=0, atol=0)


def test_lazy_layer_specs_cover_the_model():
    pipe = P.WanPipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': CFG}}, device='cpu')
    specs = pipe.to_layers()
    assert [s.typename.__name__ for s in specs] == ['InitialLayer', 'TransformerLayer', 'TransformerLayer', 'FinalLayer']
    full = P.WanModel(CFG, device='cpu')
    assert sum(s.param_count for s in specs) == sum(p.numel() for p in full.parameters())
    names = set()
    for s in specs:
        names |= {p.original_name for p in s.build().parameters()}
    assert names == {n for n, _ in full.named_parameters()}


def test_unsupported_variants_fail_loudly():
    with pytest.raises(NotImplementedError):
        P.WanModel(dict(CFG, model_type='i2v'), device='cpu')
    with pytest.raises(NotImplementedError):
        P.WanPipeline({'model': {'dtype': 'bfloat16', 'cache_text_embeddings': False, 'transformer_config': CFG}}, device='cpu')
    with pytest.raises(NotImplementedError):
        P.WanPipeline({'model': {'dtype': 'float16', 'transformer_config': CFG}}, device='cpu')
