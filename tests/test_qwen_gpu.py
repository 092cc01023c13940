"""GPU: a small Qwen-Image-architecture model through the reference-facing plugin API (QwenImagePipeline.to_layers /
prepare_inputs-shaped tuples / get_loss_fn) against the oracle (oracle/qwen_ref.py, pinned to the reference tree's
in-tree model by tests/test_oracle_qwen_golden.py) on identical inputs and weights, and against the stored output of
that in-tree model itself (tests/golden/qwen_golden.pt).

Tolerance: loss within 1e-3 relative of the oracle with the reference's bf16 rounding points emulated, 5e-3 of the
pure-fp32 oracle; parameter gradients within 5e-2 relative L2 of the emulating oracle."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))

CFG = {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}


def _make():
    from synth import fill_parameters
    from diffusion_pipe_b200.qwen_image import QwenImagePipeline
    from oracle import qwen_ref as Q
    model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'transformer_config': CFG}})
    ref = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))            # synth values are bf16-representable: exact
    return model, ref


def _batch(bs, seed, control=False):
    from oracle import qwen_ref as Q
    g = torch.Generator().manual_seed(seed)
    latents = torch.randn(bs, 16, 1, 16, 24, generator=g)
    pe = [torch.randn(12, 64, generator=g).bfloat16().float() for _ in range(bs)]
    t = torch.sigmoid(torch.randn(bs, generator=g))
    noise = torch.randn(bs, 16, 1, 16, 24, generator=g)
    ctrl = torch.randn(bs, 16, 1, 16, 24, generator=g) if control else None
    feats, (target, mask) = Q.prepare_inputs(latents, pe, t, noise, control_latents=ctrl)
    return feats, (target, torch.tensor([]))


@pytest.mark.parametrize('control', [False, True])
def test_layers_and_loss_match_oracle(control):
    from oracle import flux_ref as R
    from oracle import qwen_ref as Q
    model, ref = _make()
    feats, label = _batch(2, 1, control)
    x = tuple(f.cuda() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    loss = model.get_loss_fn()(x, tuple(l.cuda() for l in label))
    loss.backward()
    for emu, tol in ((True, 1e-3), (False, 5e-3)):
        ref.set_emulate_bf16(emu)
        ref.zero_grad()
        y = tuple(f.clone() for f in feats)
        for layer in Q.to_layers(ref):
            y = layer(y)
        rloss = R.loss_fn(y, label)
        rel = abs(loss.item() - rloss.item()) / abs(rloss.item())
        assert rel <= tol, (emu, loss.item(), rloss.item(), rel)
        if emu:
            rloss.backward()
            rg = {n: p.grad for n, p in ref.named_parameters()}
            errs = {}
            for n, p in model.transformer.named_parameters():
                if rg[n] is None:          # text stream of the last block: not on the path to the loss
                    continue
                assert p.grad is not None, n
                errs[n] = ((p.grad.float().cpu() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
            bad = sorted(((v, k) for k, v in errs.items() if v > 5e-2), reverse=True)
            assert not bad, bad[:8]


def test_forward_matches_the_reference_trees_own_model(golden_dir):
    """the stored output of submodules/ComfyUI/comfy/ldm/qwen_image/model.py (fp32) on the fixture's inputs"""
    from oracle import qwen_ref as Q
    g = torch.load(os.path.join(golden_dir, 'qwen_golden.pt'), weights_only=False)
    cfg = g['cfg']
    model, _ = _make()
    B, h, w, Lt = cfg['B'], cfg['h'], cfg['w'], cfg['Lt']
    x = Q.pack_latents(g['x']).cuda()
    mask = torch.ones(B, 1, 1, Lt + x.shape[1], dtype=torch.bool, device='cuda')
    img_shapes = torch.tensor([[(1, h // 2, w // 2)]], dtype=torch.int32, device='cuda').repeat(B, 1, 1)
    txt_seq_lens = torch.tensor([Lt], dtype=torch.int32, device='cuda').repeat(B)
    out = (x, g['ctx'].cuda(), mask, g['t'].cuda(), img_shapes, txt_seq_lens)
    with torch.no_grad():
        for layer in model.to_layers():
            out = layer(out)
    want = Q.pack_latents(g['out'])
    err = (out.float().cpu() - want).norm() / want.norm()
    assert err <= 2e-2, err.item()
