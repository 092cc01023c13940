"""CPU: the engine's global-norm gradient clipping (pipe/engine.py:_clip_grad_norm) against the
reference's own `clip_grad_norm_` (utils/patches.py:175-246), executed from its source text by
tests/golden/make_golden_clip.py -> clip_golden.json: one process, and S pipeline stages x D replicas with the collectives
of both sides replaced by the same in-process all-reduce."""
import json
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))


def _grads(ci, stage, shapes, scale):
    from synth import synth_tensor
    return [synth_tensor(tuple(s), 9000 + 100 * ci + 10 * stage + j, scale) for j, s in enumerate(shapes)]


class _FakeDist:
    """all_reduce over lists of ranks, resolved in passes (pass k collects the operands of collective k)"""
    class ReduceOp:
        SUM = 'sum'

    def __init__(self, world):
        self.w = world

    def all_reduce(self, t, op='sum', group=None):
        if group is None or len(group) == 1:
            return
        idx = self.w['call']
        key = (idx, tuple(group))
        self.w['call'] += 1
        if idx < self.w['phase']:
            vals = self.w['store'][key]
            assert set(vals) == set(group)
            t.copy_(torch.stack([vals[r] for r in group]).sum(0))
        elif idx == self.w['phase']:
            self.w['store'].setdefault(key, {})[self.w['rank']] = t.clone()


def test_clipping_matches_the_references_function(golden_dir, monkeypatch):
    from diffusion_pipe_b200.pipe import engine as E
    cases = json.load(open(os.path.join(golden_dir, 'clip_golden.json')))
    assert len(cases) == 5
    for ci, (name, c) in enumerate(cases.items()):
        S, D = c['stages'], c['replicas']
        store, got = {}, {}
        for phase in range(3):
            for stage in range(S):
                for rep in range(D):
                    rank = stage * D + rep
                    world = {'rank': rank, 'phase': phase, 'store': store, 'call': 0}
                    monkeypatch.setattr(E, 'dist', _FakeDist(world))
                    pp, dp = [s * D + rep for s in range(S)], [stage * D + r for r in range(D)]
                    grid = types.SimpleNamespace(data_parallel_size=D, get_pipe_parallel_group=lambda pp=pp: pp,
                                                 get_data_parallel_group=lambda dp=dp: dp)
                    eng = types.SimpleNamespace(device=torch.device('cpu'), is_pipe_parallel=S > 1, is_data_parallel=D > 1, grid=grid)
                    gs = _grads(ci, stage, c['shapes'][stage], c['scale'])
                    ps = [torch.nn.Parameter(torch.zeros_like(g)) for g in gs] + [torch.nn.Parameter(torch.zeros(3))]   # last: no grad
                    for p, g in zip(ps, gs):
                        p.grad = g.clone()
                    norm = E.PipelineEngine._clip_grad_norm(eng, ps, c['max_norm'])
                    if phase == 2:
                        got[rank] = (float(norm), ps)
        for rank, (norm, ps) in got.items():
            assert norm == pytest.approx(c['total_norm'], rel=1e-6), (name, rank)
            stage = rank // D
            for p, g in zip(ps, _grads(ci, stage, c['shapes'][stage], c['scale'])):
                assert torch.allclose(p.grad, g * c['clip_coef'], rtol=2e-6, atol=1e-9), (name, rank)
            assert ps[-1].grad is None
    assert cases['below_threshold']['clip_coef'] == 1.0 and cases['above_threshold']['clip_coef'] < 0.1
