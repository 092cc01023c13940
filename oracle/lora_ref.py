"""ORACLE — test infrastructure only.  LoRA adapters on the restated Flux / Qwen-Image / Wan blocks, as the reference
configures them through PEFT (models/base.py:263-303):

    peft.LoraConfig(r=rank, lora_alpha=rank, lora_dropout=0, bias='none',
                    target_modules=<every nn.Linear inside the classes listed in adapter_target_modules>)

train.py:115-133 forces alpha = rank, so `scaling = lora_alpha / r = 1`.  A PEFT `lora.Linear` computes
`result = base_layer(x); result = result + lora_B(lora_A(dropout(x))) * scaling` with lora_A ~ kaiming_uniform(a=sqrt(5)),
lora_B = 0 and every parameter except the factors frozen.

Pinning: `peft` itself is a third-party dependency (requirements.txt, not vendored, not installed here) and the reference
ships no test or golden vector for it; the forward rule above is PEFT's published `lora.Linear.forward`.  What the
reference tree does contain is the consumer of the exported adapters — submodules/ComfyUI/comfy/weight_adapter/lora.py
(`LoRAAdapter.load` :148-214 for the key layout, `calculate_weight` :224-285 for W' = W + (alpha/rank) * up @ down).
tests/golden/make_golden_lora.py runs that code on a synthetic adapter and tests/test_oracle_lora_golden.py holds this
file (forward, factor gradients) and the product's K-extended operands to its output.  Pinned to the reference TREE, not
to PEFT's own code.
Under `emulate_bf16` each Linear output (base, lora_A, lora_B) and their sum are rounded to bf16, as autocast does.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .flux_ref import RefLinear, _r

TARGET_BLOCKS = ('RefFluxTransformerBlock', 'RefFluxSingleTransformerBlock', 'RefQwenImageTransformerBlock',
                 'RefWanAttentionBlock')


class RefLoraLinear(nn.Module):
    """base Linear (frozen; keeps the names `weight` / `bias`) + lora_A / lora_B"""

    def __init__(self, base, rank):
        super().__init__()
        self.weight, self.bias = base.weight, base.bias
        self.weight.requires_grad_(False)
        if self.bias is not None:
            self.bias.requires_grad_(False)
        self.lora_A = RefLinear(base.in_features, rank, bias=False)
        self.lora_B = RefLinear(rank, base.out_features, bias=False)
        nn.init.kaiming_uniform_(self.lora_A.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B.weight)
        self.emulate_bf16 = False

    def forward(self, x):
        e = self.emulate_bf16
        xin = _r(x, e) if e else x
        base = _r(F.linear(xin.float(), self.weight.float(), self.bias.float() if self.bias is not None else None), e)
        self.lora_A.emulate_bf16 = self.lora_B.emulate_bf16 = e
        return _r(base + self.lora_B(self.lora_A(x)), e)


def add_lora(model, rank):
    """wraps every RefLinear inside the target block classes; freezes everything else.  Returns the wrapped modules."""
    for p in model.parameters():
        p.requires_grad_(False)
    wrapped = []

    def visit(parent):
        for name, child in list(parent.named_children()):
            if isinstance(child, RefLinear):
                new = RefLoraLinear(child, rank)
                if isinstance(parent, (nn.ModuleList, nn.Sequential)):
                    parent[int(name)] = new
                else:
                    setattr(parent, name, new)
                wrapped.append(new)
            else:
                visit(child)
    for m in model.modules():
        if type(m).__name__ in TARGET_BLOCKS:
            visit(m)
    return wrapped


# ---- `transformer_dtype = 'float8'`: which weights the reference stores in fp8, restated per family ----------------------
# The reference casts the stored tensor (`p.data = p.data.to(transformer_dtype)`) and autocast widens it back to bf16
# inside nn.Linear, so the arithmetic sees W rounded to the nearest fp8 value.
FP8_RULES = {
    # models/flux.py:79,203-205
    'flux': lambda name, p: not (any(k in name for k in ('time_text_embed', 'context_embedder', 'x_embedder'))
                                 or name.startswith('proj_out') or name.startswith('norm_out') or p.ndim == 1),
    # models/qwen_image.py:23,261-263
    'qwen_image': lambda name, p: not (any(k in name for k in ('time_text_embed', 'img_in', 'txt_in', 'norm_out', 'proj_out'))
                                       or p.ndim == 1),
    # models/wan/wan.py:25,233-235
    'wan': lambda name, p: not any(k in name for k in ('norm', 'bias', 'patch_embedding', 'text_embedding', 'time_embedding',
                                                       'time_projection', 'head', 'modulation')),
}


def fp8_stored_names(model, family):
    """names of the parameters the reference would hold in fp8 (adapter factors are created afterwards in the adapter
    dtype: utils/patches.py:61-69 keeps them out of float8)"""
    rule = FP8_RULES[family]
    return {n for n, p in model.named_parameters() if '.lora_A.' not in n and '.lora_B.' not in n and rule(n, p)}


def round_base_through_fp8(model, family, fp8_dtype=torch.float8_e4m3fn):
    """W <- widen(fp8(W)) for every parameter of fp8_stored_names: what the reference computes with under transformer_dtype"""
    names = fp8_stored_names(model, family)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n in names:
                p.copy_(p.to(fp8_dtype).to(p.dtype))
    return names
