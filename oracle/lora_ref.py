"""ORACLE — test infrastructure only.  LoRA adapters on the restated Flux / Qwen-Image / Wan blocks, as the reference
configures them through PEFT (models/base.py:263-303):

    peft.LoraConfig(r=rank, lora_alpha=rank, lora_dropout=0, bias='none',
                    target_modules=<every nn.Linear inside the classes listed in adapter_target_modules>)

train.py:115-133 forces alpha = rank, so `scaling = lora_alpha / r = 1`.  A PEFT `lora.Linear` computes
`result = base_layer(x); result = result + lora_B(lora_A(dropout(x))) * scaling` with lora_A ~ kaiming_uniform(a=sqrt(5)),
lora_B = 0 and every parameter except the factors frozen.

PARITY UNPINNED for this file: `peft` is a third-party dependency (requirements.txt, not vendored, not installed here) and
the reference ships no test or golden vector for it; the forward rule above is PEFT's published `lora.Linear.forward`.
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
