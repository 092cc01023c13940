"""Times one Wan attention block (forward + backward, fused autograd function) and its full-width RMSNorm+RoPE kernels at
real shapes on one GPU.  Not a bench line: per-block evidence for DESIGN.md / profiles.

    python tools/probe_wan_block.py [--dim 5120 --ffn 13824 --heads 40 --frames 9 --h 32 --w 32 --text 512]

Wan2.1-14B t2v, 33 frames 512x512: latent grid (9, 32, 32) after the (1,2,2) patchify -> L = 9216 tokens.
Algorithmic FLOPs per block forward: self-attn GEMMs 8 L D^2, self-attention 4 L^2 D, cross-attention
4 L D^2 + 4 Lc D^2 + 4 L Lc D, FFN 4 L D F; training = 3x.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dim', type=int, default=5120)
    ap.add_argument('--ffn', type=int, default=13824)
    ap.add_argument('--heads', type=int, default=40)
    ap.add_argument('--frames', type=int, default=9)
    ap.add_argument('--h', type=int, default=32)
    ap.add_argument('--w', type=int, default=32)
    ap.add_argument('--text', type=int, default=512)
    ap.add_argument('--iters', type=int, default=5)
    a = ap.parse_args()
    from diffusion_pipe_b200 import ops
    from diffusion_pipe_b200.wan import WanAttentionBlock, wan_rope_tables
    dev = 'cuda'
    D, F, H, Lc = a.dim, a.ffn, a.heads, a.text
    L = a.frames * a.h * a.w
    torch.manual_seed(0)
    blk = WanAttentionBlock(D, F, H, 1e-6, torch.bfloat16, dev)
    x = (torch.randn(1, L, D, device=dev) * 0.5).bfloat16().requires_grad_(True)
    e0 = (torch.randn(1, 1, 6, D, device=dev) * 0.1).bfloat16().requires_grad_(True)
    ctx = (torch.randn(1, Lc, D, device=dev) * 0.5).bfloat16().requires_grad_(True)
    freqs = wan_rope_tables((a.frames, a.h, a.w), D // H, device=dev)
    gout = torch.randn(1, L, D, device=dev).bfloat16()

    def step():
        y = blk(x, e0, None, None, freqs, ctx, None)
        y.backward(gout)
        return y

    for _ in range(2):
        y = step()
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(y.float()).all()) and bool(torch.isfinite(x.grad.float()).all())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(a.iters):
        step()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / a.iters
    fwd = 8 * L * D * D + 4 * L * L * D + 4 * L * D * D + 4 * Lc * D * D + 4 * L * Lc * D + 4 * L * D * F
    out = {'probe': 'wan_block_fwd_bwd', 'L': L, 'D': D, 'ffn': F, 'heads': H, 'text_len': Lc, 'ms': ms,
           'tflop_fwd': fwd / 1e12, 'tflops_training': 3 * fwd / ms / 1e9, 'finite': ok}
    # the HBM-bound pre-processing kernel alone: q, k, v of self-attention (forward), algorithmic bytes = 3 * (2 + 2) * L * D
    qkv = torch.randn(L, 3 * D, device=dev).bfloat16()
    wq = torch.ones(D, device=dev).bfloat16()
    projs = [{'src': qkv[:, 0:D], 'weight': wq, 'rope': True}, {'src': qkv[:, D:2 * D], 'weight': wq, 'rope': True},
             {'src': qkv[:, 2 * D:]}]
    for _ in range(2):
        ops.wan_norm_rope_fwd(projs, 1, L, H, freqs[0], freqs[1])
    ev[0].record()
    for _ in range(10):
        ops.wan_norm_rope_fwd(projs, 1, L, H, freqs[0], freqs[1])
    ev[1].record()
    torch.cuda.synchronize()
    nms = ev[0].elapsed_time(ev[1]) / 10
    nbytes = (3 * 2 + 3 * 2 + 2 * 2) * L * D          # read q,k,v; write q,k,v head-major; write xhat_q, xhat_k
    out['wan_norm_rope_fwd_ms'] = nms
    out['wan_norm_rope_fwd_GBps'] = nbytes / nms / 1e6
    print('RESULT ' + json.dumps(out), flush=True)


if __name__ == '__main__':
    try:
        main()
    except Exception as e:   # a probe must never take the rest of a gpurun command down with it
        import traceback
        traceback.print_exc()
        print('RESULT ' + json.dumps({'probe': 'wan_block_fwd_bwd', 'error': repr(e)}), flush=True)
