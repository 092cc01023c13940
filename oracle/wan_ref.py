"""ORACLE — test infrastructure only.  Plain-PyTorch restatement of the reference's Wan (t2v) training path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file; the product package
(diffusion-pipe_b200/) never does.

What is restated, and from where (paths relative to the reference repo root):
  * sinusoidal_embedding_1d, rope_params, rope_apply ......... models/wan/model.py:14-68
  * WanRMSNorm (full-width), WanLayerNorm .................... models/wan/model.py:71-103
  * WanSelfAttention / WanCrossAttention ..................... models/wan/model.py:106-183
  * WanAttentionBlock (modulation, gated residuals, ffn) ..... models/wan/model.py:242-318
  * Head (fp32 autocast) and unpatchify ...................... models/wan/model.py:321-349,492-517
  * WanModel parameter tree (t2v) ............................ models/wan/model.py:368-489
  * pipeline layers InitialLayer / TransformerLayer / FinalLayer and their tuple protocol
    (x, e, e0, seq_lens, grid_sizes, freqs, context) ......... models/wan/wan.py:414-546
  * prepare_inputs (t from the 10 000-bucket table, noising) . models/wan/wan.py:332-373, utils/common.py:124-160
  * attention itself is flash_attn's varlen kernel (models/wan/attention.py:17-140, third-party, CUDA only): restated
    as softmax(q k^T / sqrt(d)) v over the first k_lens keys.

PARITY PIN: unlike Flux / Qwen (diffusers), this reference model is in-tree and importable (with a 4-line stand-in for
the two diffusers base classes it inherits from and an SDPA stand-in for flash_attention).
tests/golden/make_golden_wan.py runs the reference's own WanModel pieces in fp32 on CPU and stores inputs, outputs and
gradient fingerprints in tests/golden/wan_golden.pt; tests/test_oracle_wan_golden.py checks this file against them.

`emulate_bf16=True` rounds where the reference's bf16 parameters + `torch.autocast('cuda', bf16)` produce bf16 tensors.
For Wan that is after EVERY elementwise op of the block (all operands are bf16 tensors: norm(x).type_as(x),
1 + e, the products and sums), unlike Flux where LayerNorm's fp32 output keeps the modulation in fp32.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .flux_ref import RefLinear, _r, make_contiguous


def sinusoidal_embedding_1d(dim, position):
    half = dim // 2
    position = position.type(torch.float32)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_angles(max_seq_len, dim, theta=10000):
    """angle table of rope_params (the reference stores torch.polar(1, angle))."""
    return torch.outer(torch.arange(max_seq_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))


def wan_rope_tables(grid, head_dim=128):
    """rope_apply's per-token multipliers for one (f, h, w) grid as real tables: (cos, sin) fp32 [f*h*w, head_dim], every
    frequency repeated twice.  Frequencies: d - 4(d//6) dims for the frame axis, 2(d//6) each for height and width
    (models/wan/model.py:478-484), positions 0..n-1 on each axis (:57-61)."""
    f, h, w = grid
    d = head_dim
    c = d // 2
    a = torch.cat([rope_angles(1024, d - 4 * (d // 6)), rope_angles(1024, 2 * (d // 6)), rope_angles(1024, 2 * (d // 6))], dim=1)
    parts = a.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    ang = torch.cat([parts[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), parts[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                     parts[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, -1)
    return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)


def apply_rope(x, cos, sin):
    """complex multiply of interleaved pairs, fp32 (rope_apply :64-65); x [B, L, H, D], tables [L, D]."""
    xr, xi = x.float().reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(-2)
    return x.float() * cos[None, :, None, :] + rot * sin[None, :, None, :]


def attention(q, k, v, emulate, k_lens=None):
    """flash_attention(q, k, v, k_lens): [B, L, H, D] in, [B, L, H, D] out (bf16 probabilities / output when emulating)."""
    q, k, v = (t.permute(0, 2, 1, 3).float() for t in (q, k, v))
    s = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    if k_lens is not None:
        mask = torch.arange(k.shape[2])[None, :] < torch.as_tensor(k_lens)[:, None]
        s = s.masked_fill(~mask[:, None, None, :], float('-inf'))
    m = s.amax(-1, keepdim=True)
    p = torch.exp(s - m)
    o = torch.matmul(_r(p, emulate), v) / p.sum(-1, keepdim=True)
    return _r(o.permute(0, 2, 1, 3), emulate)


class RefWanRMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.emulate_bf16 = False

    def forward(self, x):
        xf = x.float()
        xh = _r(xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + self.eps), self.emulate_bf16)
        return _r(xh * self.weight.float(), self.emulate_bf16)


def wan_layer_norm(x, eps, emulate, weight=None, bias=None):
    y = F.layer_norm(x.float(), (x.shape[-1],), weight.float() if weight is not None else None,
                     bias.float() if bias is not None else None, eps)
    return _r(y, emulate)


class RefWanSelfAttention(nn.Module):
    def __init__(self, dim, num_heads, eps=1e-6):
        super().__init__()
        self.num_heads = num_heads
        self.q, self.k, self.v, self.o = RefLinear(dim, dim), RefLinear(dim, dim), RefLinear(dim, dim), RefLinear(dim, dim)
        self.norm_q, self.norm_k = RefWanRMSNorm(dim, eps), RefWanRMSNorm(dim, eps)
        self.emulate_bf16 = False

    def forward(self, x, seq_lens, cos, sin):
        b, s, n = x.shape[0], x.shape[1], self.num_heads
        e = self.emulate_bf16
        q = self.norm_q(self.q(x)).view(b, s, n, -1)
        k = self.norm_k(self.k(x)).view(b, s, n, -1)
        v = self.v(x).view(b, s, n, -1)
        q, k = _r(apply_rope(q, cos, sin), e), _r(apply_rope(k, cos, sin), e)
        return self.o(attention(q, k, v, e, seq_lens).flatten(2))


class RefWanCrossAttention(RefWanSelfAttention):
    def forward(self, x, context, context_lens=None):
        b, n = x.shape[0], self.num_heads
        q = self.norm_q(self.q(x)).view(b, -1, n, x.shape[-1] // n)
        k = self.norm_k(self.k(context)).view(b, -1, n, x.shape[-1] // n)
        v = self.v(context).view(b, -1, n, x.shape[-1] // n)
        return self.o(attention(q, k, v, self.emulate_bf16, context_lens).flatten(2))


class RefWanAttentionBlock(nn.Module):
    def __init__(self, dim, ffn_dim, num_heads, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.self_attn = RefWanSelfAttention(dim, num_heads, eps)
        self.norm3 = nn.LayerNorm(dim, eps, elementwise_affine=True)        # cross_attn_norm=True (WanModel default)
        self.cross_attn = RefWanCrossAttention(dim, num_heads, eps)
        self.ffn = nn.Sequential(RefLinear(dim, ffn_dim), nn.GELU(approximate='tanh'), RefLinear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)
        self.emulate_bf16 = False

    def forward(self, x, e, seq_lens, cos, sin, context, context_lens=None):
        r = lambda t: _r(t, self.emulate_bf16)
        x = x.float()
        e = r(self.modulation.float().unsqueeze(0) + e.float()).chunk(6, dim=2)          # [B, 1, 1, D] each
        e = [t.squeeze(2) for t in e]

        def mod(xn, scale, shift):
            return r(r(xn * r(1 + scale)) + shift)
        y = self.self_attn(mod(wan_layer_norm(x, self.eps, self.emulate_bf16), e[1], e[0]), seq_lens, cos, sin)
        x = r(x + r(y * e[2]))
        x = r(x + self.cross_attn(wan_layer_norm(x, self.eps, self.emulate_bf16, self.norm3.weight, self.norm3.bias),
                                  context, context_lens))
        h = self.ffn[0](mod(wan_layer_norm(x, self.eps, self.emulate_bf16), e[4], e[3]))
        y = self.ffn[2](r(F.gelu(h, approximate='tanh')))
        return r(x + r(y * e[5]))


class RefWanHead(nn.Module):
    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)
        self.emulate_bf16 = False

    def forward(self, x, e):
        r = lambda t: _r(t, self.emulate_bf16)
        e = r(self.modulation.float().unsqueeze(0) + e.float().unsqueeze(2)).chunk(2, dim=2)
        xm = r(r(wan_layer_norm(x, self.eps, self.emulate_bf16) * r(1 + e[1].squeeze(2))) + e[0].squeeze(2))
        return F.linear(xm, self.head.weight.float(), self.head.bias.float())             # fp32 autocast: no rounding


class RefWanModel(nn.Module):
    """Parameter container with the reference's names (WanModel, model_type 't2v')."""

    def __init__(self, dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, in_dim=16, out_dim=16, text_dim=4096,
                 text_len=512, freq_dim=256, patch_size=(1, 2, 2), eps=1e-6, model_type='t2v'):
        super().__init__()
        assert model_type in ('t2v', 'i2v_v2')       # i2v_v2 = Wan2.2 I2V: first-frame latents + mask as extra input channels
        self.model_type = model_type
        self.dim, self.num_heads, self.freq_dim, self.text_len = dim, num_heads, freq_dim, text_len
        self.patch_size, self.out_dim = tuple(patch_size), out_dim
        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
        self.text_embedding = nn.Sequential(RefLinear(text_dim, dim), nn.GELU(approximate='tanh'), RefLinear(dim, dim))
        self.time_embedding = nn.Sequential(RefLinear(freq_dim, dim), nn.SiLU(), RefLinear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), RefLinear(dim, dim * 6))
        self.blocks = nn.ModuleList([RefWanAttentionBlock(dim, ffn_dim, num_heads, eps) for _ in range(num_layers)])
        self.head = RefWanHead(dim, out_dim, patch_size, eps)
        for name, p in self.named_parameters():
            p.original_name = name

    def set_emulate_bf16(self, flag):
        self._emulate_flag = flag
        for m in self.modules():
            if hasattr(m, 'emulate_bf16'):
                m.emulate_bf16 = flag
        return self

    def unpatchify(self, x, grid_sizes):
        c = self.out_dim
        out = []
        for u, v in zip(x, grid_sizes.tolist()):
            u = u[:math.prod(v)].view(*v, *self.patch_size, c)
            u = torch.einsum('fhwpqrc->cfphqwr', u)
            out.append(u.reshape(c, *[i * j for i, j in zip(v, self.patch_size)]))
        return out


# ---- pipeline layers (models/wan/wan.py:414-546), t2v with cached text embeddings ---------------------------------------
class RefInitialLayer(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.m = [model]
        self.patch_embedding, self.time_embedding = model.patch_embedding, model.time_embedding
        self.text_embedding, self.time_projection = model.text_embedding, model.time_projection
        self.emulate_bf16 = False

    def forward(self, inputs):
        for item in inputs:
            if torch.is_floating_point(item) and item.numel() > 0:
                item.requires_grad_(True)
        m = self.m[0]
        emu = self.emulate_bf16
        x, y, t, text_embeddings, text_seq_lens, clip_fea = inputs
        context = [emb[:length] for emb, length in zip(text_embeddings, text_seq_lens)]
        if m.model_type == 'i2v_v2':                  # models/wan/wan.py:459-465: [x | mask (first frame = 1) | y] channels
            bs, _, f, h, w_ = x.shape
            mask = torch.zeros((bs, 4, f, h, w_), dtype=x.dtype)
            mask[:, :, 0, ...] = 1
            x = torch.cat([x, mask, y], dim=1)
        w, b = self.patch_embedding.weight.float(), self.patch_embedding.bias.float()
        x = [_r(F.conv3d(_r(u.unsqueeze(0).float(), emu), w, b, stride=m.patch_size), emu) for u in x]
        grid_sizes = torch.stack([torch.tensor(u.shape[2:], dtype=torch.long) for u in x])
        x = [u.flatten(2).transpose(1, 2) for u in x]
        seq_lens = torch.tensor([u.size(1) for u in x], dtype=torch.long)
        seq_len = seq_lens.max()
        x = torch.cat([torch.cat([u, u.new_zeros(1, seq_len - u.size(1), u.size(2))], dim=1) for u in x])
        t = t.unsqueeze(-1) if t.dim() == 1 else t
        bt = t.size(0)
        e = self.time_embedding(sinusoidal_embedding_1d(m.freq_dim, t.flatten()).unflatten(0, (bt, 1)).to(torch.float32))
        e0 = self.time_projection(e).unflatten(2, (6, m.dim))
        context = self.text_embedding(torch.stack([torch.cat([u, u.new_zeros(m.text_len - u.size(0), u.size(1))]) for u in context]).float())
        cos, sin = wan_rope_tables(tuple(grid_sizes[0].tolist()), m.dim // m.num_heads)
        return make_contiguous(x, e, e0, seq_lens, grid_sizes, torch.stack([cos, sin]), context)


class RefTransformerLayer(nn.Module):
    def __init__(self, block):
        super().__init__()
        self.block = block

    def forward(self, inputs):
        x, e, e0, seq_lens, grid_sizes, freqs, context = inputs
        x = self.block(x, e0, seq_lens.tolist(), freqs[0], freqs[1], context, None)
        return make_contiguous(x, e, e0, seq_lens, grid_sizes, freqs, context)


class RefFinalLayer(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.m = [model]
        self.head = model.head

    def forward(self, inputs):
        x, e, e0, seq_lens, grid_sizes, freqs, context = inputs
        x = self.head(x, e)
        return torch.stack(self.m[0].unpatchify(x, grid_sizes), dim=0)


def to_layers(model):
    """models/wan/wan.py:375-382 (cache_text_embeddings=True)."""
    first = RefInitialLayer(model)
    first.emulate_bf16 = getattr(model, '_emulate_flag', False)
    return [first] + [RefTransformerLayer(b) for b in model.blocks] + [RefFinalLayer(model)]


def t_distribution(method='logit_normal', sigmoid_scale=1.0):
    """utils/common.py:124-146."""
    dist = torch.distributions.normal.Normal(0, 1) if method == 'logit_normal' else torch.distributions.uniform.Uniform(0, 1)
    n = 10_000
    t = dist.icdf(torch.linspace(1 / n, 1 - 1 / n, n))
    return torch.sigmoid(t * sigmoid_scale) if method == 'logit_normal' else t


def prepare_inputs(latents, text_embeddings, seq_lens, t, noise, mask=None, y=None):
    """models/wan/wan.py:332-373 (t2v; i2v_v2 when the first-frame conditioning latents `y` are given) with the random
    draws (t in [0,1], x_0) passed in."""
    latents = latents.float()
    bs, c, f, h, w = latents.shape
    if mask is not None:
        mask = F.interpolate(mask.unsqueeze(1), size=(h, w), mode='nearest-exact').unsqueeze(2)
    te = t.view(-1, 1, 1, 1, 1)
    x_t = (1 - te) * latents + te * noise
    target = noise - latents
    none = torch.tensor([])           # utils/dataset.py:1277-1279: None travels through the pipeline as an empty tensor
    return (x_t, none if y is None else y, t * 1000, text_embeddings, seq_lens, none), (target, mask)
