"""Generates tests/golden/lora_golden.pt by running the REFERENCE TREE's own LoRA consumer on a synthetic adapter.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_lora.py

PEFT (the library the reference trains adapters with, models/base.py:263-303) is absent from this image, so the adapter
path cannot be pinned to it.  What the reference tree DOES contain is the code that consumes the adapters this path
exports: submodules/ComfyUI/comfy/weight_adapter/lora.py — `LoRAAdapter.load` recognises the exported key layout
(`<module>.lora_B.weight` = up, `<module>.lora_A.weight` = down, :162,177-180) and `calculate_weight` (:224-285) states
the arithmetic:  W' = W + strength * (alpha / rank) * (up @ down), alpha = None -> 1.  That file is imported as it is;
`comfy.model_management` (pulls in an absent native module) is replaced by a stand-in providing `cast_to_device`.

Stored: W, down (A), up (B), the merged weight W' computed by the reference tree's code, an input x and
y = x W'^T + b, the key names the loader accepted.  tests/test_oracle_lora_golden.py checks oracle/lora_ref.py (and the
key names the product exports) against it.
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import synth_tensor  # noqa: E402

COMFY = '/root/reference/submodules/ComfyUI/comfy'
OUT = os.path.join(HERE, 'lora_golden.pt')


def load_reference_adapter_code():
    comfy = types.ModuleType('comfy'); comfy.__path__ = [COMFY]
    mm = types.ModuleType('comfy.model_management')
    mm.cast_to_device = lambda t, device, dtype, copy=False: t.to(device=device, dtype=dtype)
    wa = types.ModuleType('comfy.weight_adapter'); wa.__path__ = [os.path.join(COMFY, 'weight_adapter')]
    comfy.model_management = mm
    for n, m in (('comfy', comfy), ('comfy.model_management', mm), ('comfy.weight_adapter', wa)):
        sys.modules[n] = m
    for name in ('base', 'lora'):
        spec = importlib.util.spec_from_file_location('comfy.weight_adapter.' + name, os.path.join(COMFY, 'weight_adapter', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules['comfy.weight_adapter.' + name] = mod
        spec.loader.exec_module(mod)
    return sys.modules['comfy.weight_adapter.lora']


def main():
    L = load_reference_adapter_code()
    N, K, r = 48, 32, 8
    W = synth_tensor((N, K), 701, 0.2)
    bias = synth_tensor((N,), 702, 0.1)
    A = synth_tensor((r, K), 703, 0.3)          # lora_A.weight  (down)
    B = synth_tensor((N, r), 704, 0.3)          # lora_B.weight  (up)
    x = synth_tensor((5, K), 705, 1.0)
    module = 'diffusion_model.transformer_blocks.0.attn.to_q'
    sd = {module + '.lora_A.weight': A, module + '.lora_B.weight': B}
    accepted = set()
    adapter = L.LoRAAdapter.load(module, sd, None, None, accepted)          # alpha None -> scale 1 (train.py:115-133: alpha = rank)
    assert adapter is not None and accepted == set(sd)
    merged = adapter.calculate_weight(W.clone(), module + '.weight', 1.0, 1.0, None, lambda a: a)
    # the same adapter with an explicit alpha = rank stored next to it (what PEFT's config says): identical
    adapter2 = L.LoRAAdapter.load(module, sd, float(r), None, set())
    merged2 = adapter2.calculate_weight(W.clone(), module + '.weight', 1.0, 1.0, None, lambda a: a)
    assert torch.equal(merged, merged2)
    y = torch.nn.functional.linear(x, merged, bias)
    torch.save({'N': N, 'K': K, 'r': r, 'W': W, 'bias': bias, 'A': A, 'B': B, 'x': x, 'merged': merged, 'y': y,
                'module': module, 'accepted_keys': sorted(accepted)}, OUT)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
