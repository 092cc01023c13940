"""CPU: the HOST logic of the Wan block / model (diffusion-pipe_b200/wan.py: buffer plumbing, gradient routing, what is
saved for backward, parameter-gradient accumulation over micro-batches, deferred weight gradients) with the kernel
wrappers replaced by the PyTorch test doubles of tests/kernel_doubles.py, against the oracle.  The real kernels are
checked by tests/test_wan_gpu.py."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

CFG = {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 2, 'text_dim': 64, 'text_len': 16}


@pytest.fixture
def doubles(monkeypatch):
    import kernel_doubles
    from diffusion_pipe_b200 import ops
    kernel_doubles.install(monkeypatch, ops)
    return ops


def _make():
    from synth import fill_parameters
    from diffusion_pipe_b200.wan import WanPipeline
    from oracle import wan_ref as W
    model = WanPipeline({'model': {'dtype': 'bfloat16', 'transformer_config': CFG}}, device='cpu')
    ref = fill_parameters(W.RefWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    return model, ref


def _batch(bs, seed):
    from oracle import wan_ref as W
    g = torch.Generator().manual_seed(seed)
    latents = torch.randn(bs, 16, 2, 8, 8, generator=g)
    text = torch.randn(bs, 16, 64, generator=g).bfloat16().float()
    lens = torch.tensor([10, 16][:bs])
    t = torch.sigmoid(torch.randn(bs, generator=g))
    noise = torch.randn(bs, 16, 2, 8, 8, generator=g)
    feats, (target, mask) = W.prepare_inputs(latents, text, lens, t, noise)
    return feats, (target, torch.tensor([]))


def _run_product(model, feats, label):
    x = tuple(f.clone() for f in feats)
    for layer in model.to_layers():
        x = layer(x)
    loss = model.get_loss_fn()(x, label)
    loss.backward()
    return loss


def _run_oracle(ref, feats, label):
    from oracle import flux_ref as R
    from oracle import wan_ref as W
    y = tuple(f.clone() for f in feats)
    for layer in W.to_layers(ref):
        y = layer(y)
    loss = R.loss_fn(y, label)
    loss.backward()
    return loss


def test_model_forward_backward_matches_oracle(doubles):
    model, ref = _make()
    ref.set_emulate_bf16(True)
    feats, label = _batch(2, 1)
    loss = _run_product(model, feats, label)
    rloss = _run_oracle(ref, feats, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    errs = {}
    for n, p in model.transformer.named_parameters():
        assert p.grad is not None, n
        errs[n] = ((p.grad.float() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
    bad = sorted(((v, k) for k, v in errs.items() if v > 5e-2), reverse=True)
    assert not bad, bad[:8]


def test_gradients_accumulate_over_micro_batches_and_deferral_is_equivalent(doubles):
    """two micro-batches: .grad accumulates (fused q/k/v buffers included); running the weight-gradient closures later
    (zero-bubble W pass) gives the same gradients as running them in place"""
    ops = doubles
    model, ref = _make()
    b1, b2 = _batch(1, 5), _batch(1, 6)
    _run_product(model, *b1)
    _run_product(model, *b2)
    want = {n: p.grad.float().clone() for n, p in model.transformer.named_parameters()}
    model2, _ = _make()
    for feats, label in (b1, b2):
        queue = []
        ops.WGRAD_DEFER = queue
        try:
            _run_product(model2, feats, label)
        finally:
            ops.WGRAD_DEFER = None
        assert queue, 'no weight-gradient work was deferred'
        for fn in queue:
            fn()
    for n, p in model2.transformer.named_parameters():
        assert p.grad is not None, n
        torch.testing.assert_close(p.grad.float(), want[n], rtol=2e-2, atol=2e-3 * want[n].abs().max().item() + 1e-6)
    # and they are the sum of the oracle's two micro-batch gradients
    ref.set_emulate_bf16(True)
    _run_oracle(ref, *b1)
    _run_oracle(ref, *b2)
    for n, p in ref.named_parameters():
        rel = ((want[n] - p.grad).norm() / (p.grad.norm() + 1e-12)).item()
        assert rel <= 6e-2, (n, rel)


@pytest.mark.parametrize('family', ['wan', 'flux_double', 'flux_single', 'wan+lora', 'flux_double+lora', 'flux_single+lora'])
def test_deferred_weight_gradients_do_not_read_boundary_tensors(doubles, family):
    """zero-bubble contract (pipe/engine.py:_exec_backward_input): block inputs received from another stage are views of
    a mailbox slot that is handed back right after the input-gradient pass; emulate the sender overwriting the slot
    (NaN-poison every input in place) before the queued weight-gradient closures run"""
    ops = doubles
    torch.manual_seed(0)
    family, _, adapter = family.partition('+')
    if family == 'wan':
        from diffusion_pipe_b200.wan import WanAttentionBlock, wan_rope_tables
        mk = lambda: WanAttentionBlock(256, 512, 2, 1e-6, torch.bfloat16, 'cpu')
        freqs = wan_rope_tables((2, 4, 4), 128)
        ins = lambda: [(0.5 * torch.randn(2, 32, 256)).bfloat16().requires_grad_(True), (0.1 * torch.randn(2, 1, 6, 256)).bfloat16().requires_grad_(True),
                       (0.5 * torch.randn(2, 16, 256)).bfloat16().requires_grad_(True)]
        call = lambda blk, t: blk(t[0], t[1], None, None, freqs, t[2], None)
    else:
        from diffusion_pipe_b200.flux_blocks import FluxSingleTransformerBlock, FluxTransformerBlock
        from oracle.flux_ref import flux_rope_tables
        cls = FluxTransformerBlock if family == 'flux_double' else FluxSingleTransformerBlock
        mk = lambda: cls(256, 2, 2, torch.bfloat16, 'cpu')
        ids = torch.zeros(24, 3)
        ids[8:, 1] = torch.arange(16) // 4
        ids[8:, 2] = torch.arange(16) % 4
        cos, sin = flux_rope_tables(ids)
        ins = lambda: [(0.5 * torch.randn(2, 16, 256)).bfloat16().requires_grad_(True), (0.5 * torch.randn(2, 8, 256)).bfloat16().requires_grad_(True),
                       (0.5 * torch.randn(2, 256)).bfloat16().requires_grad_(True)]

        def call(blk, t):
            e, h = blk(t[0], t[1], t[2], (cos, sin))
            return torch.cat([e, h], dim=1)
    torch.manual_seed(1)
    blk = mk()
    with torch.no_grad():
        for p in blk.parameters():
            if p.ndim == 1:
                p.add_(0.05 * torch.randn_like(p.float()).to(p.dtype))
    state = {k: v.clone() for k, v in blk.state_dict().items()}
    torch.manual_seed(2)
    t = ins()
    gout = None

    def run(poison):
        nonlocal gout
        b = mk()
        b.load_state_dict(state)
        if adapter:
            from diffusion_pipe_b200 import lora
            assert lora.attach(b, 16) == 1
            gl = torch.Generator().manual_seed(7)
            with torch.no_grad():
                for n, p in b.named_parameters():
                    if '.lora_' in n:
                        p.copy_((0.05 * torch.randn(p.shape, generator=gl)).bfloat16())
        torch.manual_seed(2)
        tt = ins()
        y = call(b, tt)
        if gout is None:
            gout = torch.randn_like(y.float()).to(y.dtype)
        q = []
        if poison:
            ops.WGRAD_DEFER = q
        try:
            y.backward(gout)
        finally:
            ops.WGRAD_DEFER = None
        if poison:
            assert q
            with torch.no_grad():
                for x in tt:
                    x.fill_(float('nan'))                 # the previous stage reuses the slot
                for fn in q:
                    fn()
        return {n: p.grad.float().clone() for n, p in b.named_parameters() if p.grad is not None}
    base, deferred = run(False), run(True)
    assert base.keys() == deferred.keys() and base
    for n in base:
        assert torch.isfinite(deferred[n]).all(), n
        torch.testing.assert_close(deferred[n], base[n], rtol=1e-5, atol=1e-6, msg=n)


def make_i2v_pair(device='cpu'):
    from synth import fill_parameters
    from diffusion_pipe_b200.wan import WanPipeline
    from oracle import wan_ref as W
    cfg = dict(CFG, model_type='i2v_v2', num_layers=1)
    model = WanPipeline({'model': {'dtype': 'bfloat16', 'device': device, 'transformer_config': cfg}})
    assert model.transformer.patch_embedding.weight.shape[1] == 36
    ref = fill_parameters(W.RefWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=64, text_len=16, in_dim=36, model_type='i2v_v2'))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    ref.set_emulate_bf16(True)
    return model, ref


def i2v_batch(model):
    g = torch.Generator().manual_seed(8)
    batch = {'latents': torch.randn(2, 16, 2, 8, 8, generator=g), 'y': torch.randn(2, 16, 2, 8, 8, generator=g),
             'text_embeddings': torch.randn(2, 16, 64, generator=g).bfloat16().float(), 'seq_lens': torch.tensor([7, 16]), 'mask': None}
    torch.manual_seed(3)
    feats, (target, _) = model.prepare_inputs(batch)
    assert feats[1] is batch['y']
    none = torch.tensor([])
    return tuple(none if f is None else f for f in feats), (target, none)


def test_wan22_i2v_forward_backward_matches_oracle(doubles):
    """model_type 'i2v_v2': first-frame mask + conditioning latents as extra patch-embedding channels (K = 144 GEMM)"""
    model, ref = make_i2v_pair()
    feats, label = i2v_batch(model)
    loss = _run_product(model, feats, label)
    rloss = _run_oracle(ref, feats, label)
    assert abs(loss.item() - rloss.item()) / abs(rloss.item()) <= 1e-3, (loss.item(), rloss.item())
    rg = {n: p.grad for n, p in ref.named_parameters()}
    for n, p in model.transformer.named_parameters():
        rel = ((p.grad.float() - rg[n]).norm() / (rg[n].norm() + 1e-12)).item()
        assert rel <= 5e-2, (n, rel)
    # a t2v model must not be fed `y`, an i2v_v2 model must be
    from diffusion_pipe_b200.wan import WanPipeline
    t2v = WanPipeline({'model': {'dtype': 'bfloat16', 'device': 'cpu', 'transformer_config': dict(CFG, num_layers=1)}})
    with pytest.raises(NotImplementedError):
        t2v.to_layers()[0](tuple(f.clone() for f in feats))
