"""Generates tests/golden/flux_bfl_map.json: the diffusers -> BFL (ComfyUI) key map the REFERENCE builds for its Flux export.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_flux_export.py

`BFL_TO_DIFFUSERS_MAP` (models/flux.py:22-76), the block counts (:19-20) and `make_diffusers_to_bfl_map` (:84-112) are taken from the source text (ast)
and executed; the result {diffusers key: [index, bfl key]} for the 19 + 38 block model is stored as it is.
tests/test_flux_export.py holds diffusion-pipe_b200/flux_export.py (a rule-based statement of the same layout) to it and
checks the concatenation order and the final-layer (scale, shift) swap of FluxPipeline.save_model (:257-288).
"""
import ast
import json
import os

REF = '/root/reference/models/flux.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'flux_bfl_map.json')


def main():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if (isinstance(n, ast.Assign) and getattr(n.targets[0], 'id', '') in ('BFL_TO_DIFFUSERS_MAP', 'NUM_DOUBLE_BLOCKS', 'NUM_SINGLE_BLOCKS'))
            or (isinstance(n, ast.FunctionDef) and n.name == 'make_diffusers_to_bfl_map')]
    assert len(body) == 4
    ns = {}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, 'exec'), ns)
    m = ns['make_diffusers_to_bfl_map']()
    with open(OUT, 'w') as f:
        json.dump({k: [int(i), b] for k, (i, b) in sorted(m.items())}, f, indent=0)
    print('wrote', OUT, len(m), 'keys', os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
