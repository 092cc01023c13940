"""ORACLE — test infrastructure only.  Plain-PyTorch restatement of the reference's Qwen-Image training path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file; the product package
(diffusion-pipe_b200/) never does.

What is restated, and from where (paths relative to the reference repo root):
  * pipeline layers and their tuple protocol ........... models/qwen_image.py:519-605 (InitialLayer, TransformerLayer,
    FinalLayer), models/base.py:37-38 (make_contiguous)
  * joint attention (separate q/k/v linears with bias, per-head RMSNorm, complex RoPE with separate image / text
    tables, [text, image] order, bool key mask) ......... models/qwen_image.py:66-71,91-174
  * prepare_inputs (ragged prompt embeddings -> padded + bool mask, noising, packing, img_shapes / txt_seq_lens,
    Qwen-Image-Edit control latents) .................... models/qwen_image.py:394-488
  * the block / embedder / rope arithmetic those layers call lives in `diffusers` (requirements.txt:4, not vendored,
    not installed here): QwenImageTransformerBlock, QwenTimestepProjEmbeddings, QwenEmbedRope(scale_rope=True),
    AdaLayerNormContinuous.  Restated from the published diffusers 0.35 transformer_qwenimage.py and cross-checked
    against the in-tree statement of the same architecture,
        submodules/ComfyUI/comfy/ldm/qwen_image/model.py:68-93   (QwenTimestepProjEmbeddings)
        submodules/ComfyUI/comfy/ldm/qwen_image/model.py:96-209  (Attention)
        submodules/ComfyUI/comfy/ldm/qwen_image/model.py:212-313 (QwenImageTransformerBlock)
        submodules/ComfyUI/comfy/ldm/qwen_image/model.py:316-338 (LastLayer), :405-431 (positions), :493-494 (text positions)
    with configs/qwen_image/transformer/config.json for the sizes.

PARITY PIN: tests/golden/make_golden_qwen.py runs the ComfyUI block / last layer / embedders above (fp32, CPU) and
stores inputs, weights, outputs and gradient fingerprints in tests/golden/qwen_block_golden.pt;
tests/test_oracle_golden.py checks this restatement against them.  The diffusers-only facts that ComfyUI states
differently (odd latent heights: diffusers centres with h - h//2, ComfyUI with h//2) are noted where they occur and
are "parity unpinned".

SECOND PIN (the reference's own code): tests/golden/make_golden_qwen_attn.py executes `apply_rotary_emb_qwen` and
`QwenDoubleStreamAttnProcessor2_0.__call__` from the source text of models/qwen_image.py:26-174 — with the bool key mask
of a ragged micro-batch, which ComfyUI's model does not have — and tests/test_oracle_qwen_golden.py holds
`RefQwenAttention` (outputs, input and parameter gradients) to tests/golden/qwen_attn_golden.pt.

`emulate_bf16=True` rounds where the reference's autocast produces bf16 tensors (see oracle/flux_ref.py).
"""
import torch
import torch.nn.functional as F
from torch import nn

from .flux_ref import RefFeedForward, RefLinear, RefRMSNorm, _r, apply_rope, make_contiguous, timestep_sinusoid


def qwen_rope_tables(img_shapes, txt_len, axes_dim=(16, 56, 56), theta=10000.0):
    """diffusers QwenEmbedRope(theta, axes_dim, scale_rope=True).forward(img_shapes, txt_seq_lens): position tables
    for the image tokens (every (frame, h, w) entry of img_shapes in turn; entry idx sits at frame position idx) and
    for max(txt_seq_lens) text tokens, which start at max(h // 2, w // 2) on all three axes.  Returns
    (vid_cos, vid_sin, txt_cos, txt_sin), each [tokens, sum(axes_dim)] fp32 with every frequency repeated twice — the
    real/imag form of the complex `freqs_cis` the reference multiplies by (models/qwen_image.py:66-71)."""
    def angles(pos, d):
        freqs = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float32), torch.arange(0, d, 2).to(torch.float32).div(d))
        return torch.outer(pos.to(torch.float32), freqs)            # models/qwen_image.py:547-555 (rope_params)

    vid, max_idx = [], 0
    for idx, (frame, h, w) in enumerate(img_shapes):
        pf = torch.arange(idx, idx + frame)
        ph = torch.arange(h) - (h - h // 2)                          # centred ("scale_rope"); ComfyUI uses h // 2: same for even h
        pw = torch.arange(w) - (w - w // 2)
        a = torch.cat([angles(pf, axes_dim[0]).view(frame, 1, 1, -1).expand(frame, h, w, -1),
                       angles(ph, axes_dim[1]).view(1, h, 1, -1).expand(frame, h, w, -1),
                       angles(pw, axes_dim[2]).view(1, 1, w, -1).expand(frame, h, w, -1)], dim=-1)
        vid.append(a.reshape(frame * h * w, -1))
        max_idx = max(max_idx, h // 2, w // 2)
    vid = torch.cat(vid, dim=0)
    pt = torch.arange(max_idx, max_idx + txt_len)
    txt = torch.cat([angles(pt, d) for d in axes_dim], dim=-1)
    rep = lambda a: a.repeat_interleave(2, dim=1)
    return rep(vid.cos()), rep(vid.sin()), rep(txt.cos()), rep(txt.sin())


def masked_sdpa(q, k, v, key_mask, emulate):
    """softmax(q k^T / sqrt(d) + mask) v on [B, L, H, D]; key_mask bool [B, L] (True = attend) or None."""
    q, k, v = (t.permute(0, 2, 1, 3).float() for t in (q, k, v))
    s = torch.matmul(q, k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float('-inf'))
    m = s.amax(-1, keepdim=True)
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    o = torch.matmul(_r(p, emulate), v) / l
    return _r(o.permute(0, 2, 1, 3), emulate)


class RefQwenAttention(nn.Module):
    """diffusers Attention(added_kv_proj_dim=dim, bias=True, qk_norm='rms_norm', eps=1e-6) driven by the reference's
    QwenDoubleStreamAttnProcessor2_0 (models/qwen_image.py:91-174)."""

    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        hd = dim // heads
        self.to_q, self.to_k, self.to_v = RefLinear(dim, dim), RefLinear(dim, dim), RefLinear(dim, dim)
        self.add_q_proj, self.add_k_proj, self.add_v_proj = RefLinear(dim, dim), RefLinear(dim, dim), RefLinear(dim, dim)
        self.norm_q, self.norm_k = RefRMSNorm(hd), RefRMSNorm(hd)
        self.norm_added_q, self.norm_added_k = RefRMSNorm(hd), RefRMSNorm(hd)
        self.to_out = nn.ModuleList([RefLinear(dim, dim), nn.Identity()])
        self.to_add_out = RefLinear(dim, dim)
        self.emulate_bf16 = False

    def forward(self, img, txt, vid_freqs, txt_freqs, attention_mask=None):
        B, Li, _ = img.shape
        Lt = txt.shape[1]
        H, e = self.heads, self.emulate_bf16
        iq = self.norm_q(self.to_q(img).view(B, Li, H, -1))
        ik = self.norm_k(self.to_k(img).view(B, Li, H, -1))
        iv = self.to_v(img).view(B, Li, H, -1)
        tq = self.norm_added_q(self.add_q_proj(txt).view(B, Lt, H, -1))
        tk = self.norm_added_k(self.add_k_proj(txt).view(B, Lt, H, -1))
        tv = self.add_v_proj(txt).view(B, Lt, H, -1)
        iq, ik = _r(apply_rope(iq, *vid_freqs), e), _r(apply_rope(ik, *vid_freqs), e)
        tq, tk = _r(apply_rope(tq, *txt_freqs), e), _r(apply_rope(tk, *txt_freqs), e)
        q, k, v = torch.cat([tq, iq], 1), torch.cat([tk, ik], 1), torch.cat([tv, iv], 1)      # order: [text, image]
        km = attention_mask.reshape(B, -1) if attention_mask is not None else None
        o = masked_sdpa(q, k, v, km, e).flatten(2)
        return self.to_out[0](o[:, Lt:]), self.to_add_out(o[:, :Lt])


class _Mod(nn.Sequential):
    def __init__(self, dim):
        super().__init__(nn.SiLU(), RefLinear(dim, 6 * dim))


class RefQwenImageTransformerBlock(nn.Module):
    """diffusers QwenImageTransformerBlock.  Returns (encoder_hidden_states, hidden_states)."""

    def __init__(self, dim, heads, mlp_ratio=4):
        super().__init__()
        self.img_mod, self.txt_mod = _Mod(dim), _Mod(dim)
        self.attn = RefQwenAttention(dim, heads)
        self.img_mlp = RefFeedForward(dim, dim * mlp_ratio)
        self.txt_mlp = RefFeedForward(dim, dim * mlp_ratio)
        self.emulate_bf16 = False

    def _modulate(self, x, mod):
        shift, scale, gate = mod.chunk(3, dim=-1)
        xn = F.layer_norm(x, (x.shape[-1],), eps=1e-6)
        return xn * _r(1 + scale, self.emulate_bf16)[:, None] + shift[:, None], gate[:, None]

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb, attention_mask=None):
        e = self.emulate_bf16
        vid_freqs, txt_freqs = image_rotary_emb
        x, c = hidden_states.float(), encoder_hidden_states.float()
        st = F.silu(temb.float())
        img_mod1, img_mod2 = self.img_mod[1](st).chunk(2, dim=-1)
        txt_mod1, txt_mod2 = self.txt_mod[1](st).chunk(2, dim=-1)
        xm, xg1 = self._modulate(x, img_mod1)
        cm, cg1 = self._modulate(c, txt_mod1)
        xa, ca = self.attn(xm, cm, vid_freqs, txt_freqs, attention_mask)
        x = _r(x + _r(xg1 * xa, e), e)
        c = _r(c + _r(cg1 * ca, e), e)
        xm2, xg2 = self._modulate(x, img_mod2)
        x = _r(x + _r(xg2 * self.img_mlp(xm2), e), e)
        cm2, cg2 = self._modulate(c, txt_mod2)
        c = _r(c + _r(cg2 * self.txt_mlp(cm2), e), e)
        return c, x


class RefQwenTimestepProjEmbeddings(nn.Module):
    """diffusers QwenTimestepProjEmbeddings: Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000)
    -> TimestepEmbedding(256, dim)."""

    class _TE(nn.Module):
        def __init__(self, dim):
            super().__init__()
            self.linear_1, self.linear_2 = RefLinear(256, dim), RefLinear(dim, dim)

        def forward(self, x):
            return self.linear_2(F.silu(self.linear_1(x)))

    def __init__(self, dim):
        super().__init__()
        self.timestep_embedder = self._TE(dim)

    def forward(self, timestep):
        half = 128
        import math
        exponent = -math.log(10000.0) * torch.arange(0, half, dtype=torch.float32) / half
        emb = 1000.0 * (timestep[:, None].float() * torch.exp(exponent)[None, :])     # `scale * emb`
        return self.timestep_embedder(torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1))


class RefAdaLayerNormContinuous(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.linear = RefLinear(dim, 2 * dim)
        self.emulate_bf16 = False

    def forward(self, x, temb):
        scale, shift = self.linear(F.silu(temb.float())).chunk(2, dim=1)
        xn = F.layer_norm(x.float(), (x.shape[-1],), eps=1e-6)
        return xn * _r(1 + scale, self.emulate_bf16)[:, None, :] + shift[:, None, :]


class RefQwenImageTransformer(nn.Module):
    """Parameter container with diffusers' module names (QwenImageTransformer2DModel)."""

    def __init__(self, dim=3072, heads=24, num_layers=60, in_channels=64, out_channels=16, joint_dim=3584,
                 axes_dim=(16, 56, 56), mlp_ratio=4):
        super().__init__()
        assert dim // heads == sum(axes_dim)
        self.axes_dim = tuple(axes_dim)
        self.img_in = RefLinear(in_channels, dim)
        self.txt_norm = RefRMSNorm(joint_dim, eps=1e-6)
        self.txt_in = RefLinear(joint_dim, dim)
        self.time_text_embed = RefQwenTimestepProjEmbeddings(dim)
        self.transformer_blocks = nn.ModuleList([RefQwenImageTransformerBlock(dim, heads, mlp_ratio) for _ in range(num_layers)])
        self.norm_out = RefAdaLayerNormContinuous(dim)
        self.proj_out = RefLinear(dim, 4 * out_channels)
        for name, p in self.named_parameters():
            p.original_name = name   # models/qwen_image.py:281-282

    def set_emulate_bf16(self, flag):
        self._emulate_flag = flag
        for m in self.modules():
            if hasattr(m, 'emulate_bf16'):
                m.emulate_bf16 = flag
        return self


# ---- pipeline layers (models/qwen_image.py:519-605) ------------------------------------------------------------------
class RefInitialLayer(nn.Module):
    def __init__(self, t):
        super().__init__()
        self.img_in, self.txt_norm, self.txt_in, self.time_text_embed = t.img_in, t.txt_norm, t.txt_in, t.time_text_embed
        self.axes_dim = t.axes_dim
        self.emulate_bf16 = False

    def forward(self, inputs):
        for item in inputs:
            if torch.is_floating_point(item):
                item.requires_grad_(True)
        hidden_states, encoder_hidden_states, attention_mask, timestep, img_shapes, txt_seq_lens, *extra = inputs
        hidden_states = self.img_in(hidden_states)
        timestep = _r(timestep.float(), self.emulate_bf16)       # `timestep.to(hidden_states.dtype)` :532
        encoder_hidden_states = self.txt_in(self.txt_norm(encoder_hidden_states))
        temb = self.time_text_embed(timestep)
        shapes = [tuple(s) for s in img_shapes.tolist()[0]]       # per-sample lists are identical (:467)
        vc, vs, tc, ts = qwen_rope_tables(shapes, max(txt_seq_lens.tolist()), self.axes_dim)
        return make_contiguous(hidden_states, encoder_hidden_states, attention_mask, temb,
                               torch.stack([vc, vs]), torch.stack([tc, ts])) + tuple(extra)


class RefTransformerLayer(nn.Module):
    def __init__(self, block):
        super().__init__()
        self.block = block

    def forward(self, inputs):
        hidden_states, encoder_hidden_states, attention_mask, temb, vid_freqs, txt_freqs, *extra = inputs
        encoder_hidden_states, hidden_states = self.block(hidden_states, encoder_hidden_states, temb,
                                                          ((vid_freqs[0], vid_freqs[1]), (txt_freqs[0], txt_freqs[1])),
                                                          attention_mask)
        return make_contiguous(hidden_states, encoder_hidden_states, attention_mask, temb, vid_freqs, txt_freqs) + tuple(extra)


class RefFinalLayer(nn.Module):
    def __init__(self, t):
        super().__init__()
        self.norm_out, self.proj_out = t.norm_out, t.proj_out

    def forward(self, inputs):
        hidden_states, encoder_hidden_states, attention_mask, temb, vid_freqs, txt_freqs, *extra = inputs
        output = self.proj_out(self.norm_out(hidden_states, temb))
        if len(extra) > 0:
            assert len(extra) == 1
            output = output[:, :int(extra[0][0].item()), ...]
        return output


def to_layers(transformer):
    """models/qwen_image.py:490-496."""
    first = RefInitialLayer(transformer)
    first.emulate_bf16 = getattr(transformer, '_emulate_flag', False)
    return [first] + [RefTransformerLayer(b) for b in transformer.transformer_blocks] + [RefFinalLayer(transformer)]


def pack_latents(x):
    """diffusers QwenImagePipeline._pack_latents on [bs, C, 1, h, w]."""
    b, c = x.shape[0], x.shape[1]
    h, w = x.shape[-2], x.shape[-1]
    return x.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(b, (h // 2) * (w // 2), c * 4)


def prepare_inputs(latents, prompt_embeds, t, noise, mask=None, control_latents=None):
    """models/qwen_image.py:394-488 with the random draws (t, x_0) passed in."""
    latents = latents.float()
    bs, c, _, h, w = latents.shape
    lens = [e.size(0) for e in prompt_embeds]
    max_len = max(lens)
    pe = torch.stack([torch.cat([u, u.new_zeros(max_len - u.size(0), u.size(1))]) for u in prompt_embeds])
    pm = torch.stack([torch.cat([torch.ones(n, dtype=torch.bool), torch.zeros(max_len - n, dtype=torch.bool)]) for n in lens])
    x_1 = pack_latents(latents)
    if mask is not None:
        mask = mask.unsqueeze(1).expand((-1, c, -1, -1))
        mask = F.interpolate(mask, size=(h, w), mode='nearest-exact').unsqueeze(2)
        mask = pack_latents(mask)
    x_0 = pack_latents(noise.float())
    te = t.view(-1, 1, 1)
    x_t = (1 - te) * x_1 + te * x_0
    target = x_0 - x_1
    img_shapes = [(1, h // 2, w // 2)]
    extra = tuple()
    if control_latents is not None:
        cl = pack_latents(control_latents.float())
        extra = (torch.tensor(x_t.shape[1]).repeat((bs,)),)
        x_t = torch.cat([x_t, cl], dim=1)
        img_shapes.append((1, h // 2, w // 2))
    img_shapes = torch.tensor([img_shapes], dtype=torch.int32).repeat((bs, 1, 1))
    txt_seq_lens = torch.tensor([max_len], dtype=torch.int32).repeat((bs,))
    attention_mask = torch.cat([pm, torch.ones((bs, x_t.shape[1]), dtype=torch.bool)], dim=1).view(bs, 1, 1, -1)
    return (x_t, pe, attention_mask, t, img_shapes, txt_seq_lens) + extra, (target, mask)
