"""CPU / gloo: the pipeline engine (C++ 1F1B planner + torch.distributed links) reproduces the single-process oracle
engine (oracle/engine_ref.py) for 1 and 2 pipeline stages and for 2-way data parallelism — loss, clipped gradient
norm and post-step weights in fp32 to 1e-5.  Multi-process cases spawn world_size-2 gloo groups on 127.0.0.1."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

GAS, MBS, STEPS = 4, 2, 2


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, stages, outdir, partition, schedule='1f1b', perturb=False, gas=GAS):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import toy_model
    from diffusion_pipe_b200.pipe import ManualPipelineModule, dist, initialize
    torch.set_num_threads(1)
    dist.init_distributed('gloo')
    layers = toy_model.make_layers()
    if perturb and rank > 0:      # a replica that was initialised differently: the engine must overwrite it (E11)
        with torch.no_grad():
            for l in layers:
                for p in (l.parameters() if hasattr(l, 'parameters') else []):
                    if p.requires_grad:
                        p.add_(0.37)
    pm = ManualPipelineModule(layers=layers, num_stages=stages, partition_method=partition,
                              manual_partition_split=[3] if partition == 'manual' else None, loss_fn=toy_model.loss_fn,
                              dynamic_shape=True, device=torch.device('cpu'))
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': MBS, 'gradient_accumulation_steps': gas,
                                                   'gradient_clipping': 0.5, 'steps_per_print': 0, 'stage_link': 'dist',
                                                   'pipeline_schedule': schedule})
    params = [p for p in pm.parameters() if p.requires_grad]
    engine._configure_optimizer(lambda ps: torch.optim.SGD(ps, lr=0.1) if ps else None, params)
    dp_rank = engine.grid.get_data_parallel_rank()
    losses, norms = [], []
    for step in range(STEPS):
        engine.reset_activation_shape()
        mbs = toy_model.make_micro_batches(gas, MBS, seed=100 * step + dp_rank)
        it = iter(mbs) if (engine.is_first_stage() or engine.is_last_stage()) else None
        losses.append(float(engine.train_batch(it)))
        norms.append(float(engine._grad_norm))
    ev = float(engine.eval_batch(iter(toy_model.make_micro_batches(3, MBS, seed=999)), num_micro_batches=3)) \
        if (engine.is_first_stage() or engine.is_last_stage()) else float(engine.eval_batch(None, num_micro_batches=3))
    sd = {p.original_name: p.detach().clone() for p in pm.parameters()}
    torch.save({'losses': losses, 'norms': norms, 'eval': ev, 'params': sd, 'stage': engine.stage_id, 'dp': dp_rank,
                'parts': pm.parts, 'early': engine.dp_early_layers, 'nlayers': len(pm.forward_funcs)}, os.path.join(outdir, f'rank{rank}.pt'))
    dist.barrier()


def _reference(dp_world, gas=GAS):
    import toy_model
    from oracle.engine_ref import RefPipelineEngine
    layers = toy_model.make_layers()
    params = [p for l in layers for p in l.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.1)
    eng = RefPipelineEngine(layers, toy_model.loss_fn, opt, None, gas * dp_world, 0.5)
    losses, norms = [], []
    for step in range(STEPS):
        mbs = []
        for d in range(dp_world):
            mbs += toy_model.make_micro_batches(gas, MBS, seed=100 * step + d)
        losses.append(float(eng.train_batch(mbs)))
        norms.append(float(eng.grad_norm))
    with torch.no_grad():
        ev = sum(float(toy_model.loss_fn(eng.forward(f), l)) for f, l in toy_model.make_micro_batches(3, MBS, seed=999)) / 3
    sd = {p.original_name: p.detach().clone() for l in layers for p in l.parameters()}
    return losses, norms, ev, sd


def _run(world, stages, partition='uniform', schedule='1f1b', perturb=False, gas=GAS):
    with tempfile.TemporaryDirectory() as d:
        port = _free_port()
        if world == 1:
            _worker(0, 1, port, stages, d, partition, schedule, False, gas)
            import torch.distributed as tdist
            if tdist.is_initialized():
                tdist.destroy_process_group()
        else:
            mp.spawn(_worker, args=(world, port, stages, d, partition, schedule, perturb, gas), nprocs=world, join=True)
        return [torch.load(os.path.join(d, f'rank{r}.pt'), weights_only=False) for r in range(world)]


def _check(results, dp_world, gas=GAS):
    losses, norms, ev, sd = _reference(dp_world, gas)
    for r in results:
        assert r['losses'] == pytest.approx(losses, rel=1e-5, abs=1e-7), (r['losses'], losses)
        assert r['norms'] == pytest.approx(norms, rel=1e-5)
        assert r['eval'] == pytest.approx(ev, rel=1e-5)
        for k, v in r['params'].items():
            assert torch.allclose(v, sd[k], rtol=1e-5, atol=1e-6), k


def test_single_stage_world_1():
    res = _run(1, 1)
    assert res[0]['parts'] == [0, 6]
    _check(res, 1)


def test_two_stages_world_2():
    res = _run(2, 2, 'manual')
    assert [r['stage'] for r in res] == [0, 1] and res[0]['parts'] == [0, 3, 6]
    assert set(res[0]['params']) | set(res[1]['params']) == set(_reference(1)[3])   # each stage owns only its layers
    assert not (set(res[0]['params']) & set(res[1]['params']))
    _check(res, 1)


def test_two_stages_parameter_partition():
    res = _run(2, 2, 'parameters')
    _check(res, 1)


def test_data_parallel_world_2():
    res = _run(2, 1)
    assert [r['dp'] for r in res] == [0, 1]
    _check(res, 2)
    # the all-reduce of (all but at most the first) layers was started from inside the last micro-batch's backward pass
    assert all(r['early'] >= r['nlayers'] - 1 for r in res), [(r['early'], r['nlayers']) for r in res]


def test_engine_broadcasts_trainable_parameters_over_the_dp_group():
    """utils/patches.py:163-172: replica 1 starts from different weights and must end up training rank 0's model"""
    res = _run(2, 1, 'uniform', '1f1b', perturb=True)
    _check(res, 2)


def test_zero_bubble_schedule_single_and_two_stages():
    """the split-backward order changes when things run, not what is computed"""
    _check(_run(1, 1, 'uniform', 'zb'), 1)
    _check(_run(2, 2, 'manual', 'zb'), 1)


@pytest.mark.parametrize('schedule', ['1f1b', 'zb'])
def test_two_stages_times_two_replicas_world_4(schedule):
    """BASELINE.json configs[4] shape (pipeline x data parallel): rank = stage * dp + dp_rank, per-stage DP groups for the
    gradient all-reduce, per-replica pipe groups for the boundary traffic, the norm and the loss"""
    res = _run(4, 2, 'uniform', schedule)
    assert sorted((r['stage'], r['dp']) for r in res) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    assert [(r['stage'], r['dp']) for r in res] == [(0, 0), (0, 1), (1, 0), (1, 1)]      # pipe is the outer axis
    losses, norms, ev, sd = _reference(2)
    for r in res:
        assert r['losses'] == pytest.approx(losses, rel=1e-5, abs=1e-7), (r['losses'], losses)
        assert r['norms'] == pytest.approx(norms, rel=1e-5)
        for k, v in r['params'].items():
            assert torch.allclose(v, sd[k], rtol=1e-5, atol=1e-6), k


@pytest.mark.parametrize('stages,gas,schedule', [(4, 8, 'zb'), (4, 3, 'zb'), (6, 12, 'zb'), (6, 7, '1f1b')])
def test_deep_pipelines_match_the_single_process_engine(stages, gas, schedule):
    """4 and 6 stages (6 toy layers: 'uniform' gives uneven stages at 4, one layer per stage at 6), more and fewer
    micro-batches than the zero-bubble order may hold (2 x stages): the order of F / B / W across real processes changes
    nothing in the loss, the clipped norm or the updated weights"""
    res = _run(stages, stages, 'uniform', schedule, gas=gas)
    assert [r['stage'] for r in res] == list(range(stages))
    _check(res, 1, gas)


def _worker_shape_change(rank, world, port, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import toy_model
    from diffusion_pipe_b200.pipe import ManualPipelineModule, dist, initialize
    torch.set_num_threads(1)
    dist.init_distributed('gloo')
    pm = ManualPipelineModule(layers=toy_model.make_layers(), num_stages=2, partition_method='uniform', loss_fn=toy_model.loss_fn,
                              dynamic_shape=True, device=torch.device('cpu'))
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': MBS, 'gradient_accumulation_steps': 2,
                                                   'gradient_clipping': 0.5, 'steps_per_print': 0, 'stage_link': 'dist'})
    engine._configure_optimizer(lambda ps: torch.optim.SGD(ps, lr=0.1) if ps else None, [p for p in pm.parameters() if p.requires_grad])
    engine.reset_activation_shape()
    engine.train_batch(iter(toy_model.make_micro_batches(2, MBS, seed=1)))
    # a step with another micro-batch size WITHOUT reset_activation_shape(): the sending stage must refuse
    engine.train_batch(iter(toy_model.make_micro_batches(2, MBS + 1, seed=2)))


def test_changed_boundary_shapes_without_reset_are_refused():
    """train.py:916 / :181 call reset_activation_shape() before every step; forgetting it used to receive the new tuple
    into buffers sized for the old one without any error from the transport"""
    with tempfile.TemporaryDirectory() as d:
        with pytest.raises(Exception, match='reset_activation_shape'):
            mp.spawn(_worker_shape_change, args=(2, _free_port(), d), nprocs=2, join=True)


def _passthrough_stack(n=4):
    """layers of the shape every model definition of the reference has: a stream that each layer rewrites and a tensor that
    every layer reads and hands on UNCHANGED (the time embedding: models/flux.py:497-533, qwen_image.py:519-605)"""
    events = []

    class Layer(torch.nn.Module):
        def __init__(self, i):
            super().__init__()
            self.i, self.lin, self.mod = i, torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)

        def forward(self, x):
            h, temb = x
            layer = self

            class Tap(torch.autograd.Function):
                @staticmethod
                def forward(ctx, a):
                    return a.clone()

                @staticmethod
                def backward(ctx, g):
                    events.append(('backward', layer.i))
                    return g
            return Tap.apply(self.lin(h) * self.mod(temb)), temb
    torch.manual_seed(0)
    return [Layer(i) for i in range(n)], events


def test_grads_ready_callback_runs_right_after_each_layers_backward_despite_a_handed_through_tensor():
    """A hook on the handed-through tensor itself would complete only after the FIRST layer's backward (its gradient sums
    over every layer) and every data-parallel all-reduce would start when nothing is left to overlap with."""
    from diffusion_pipe_b200.pipe.module import PipelineModule
    layers, events = _passthrough_stack()
    x = (torch.randn(2, 4, requires_grad=True), torch.randn(2, 4, requires_grad=True))
    y = x
    for i, f in enumerate(layers):
        y = PipelineModule._arm_grads_ready(y, i, i + 1, lambda a, b: events.append(('ready', a)))
        y = f(y)
    y[0].sum().backward()
    assert events == [(k, i) for i in (3, 2, 1, 0) for k in ('backward', 'ready')], events
    # parameter gradients of a layer exist when its callback runs (AccumulateGrad nodes run before any other ready node)
    layers, events = _passthrough_stack()
    seen = {}
    y = (x[0].detach().requires_grad_(True), x[1].detach().requires_grad_(True))
    for i, f in enumerate(layers):
        y = PipelineModule._arm_grads_ready(y, i, i + 1, lambda a, b: seen.setdefault(a, [p.grad is not None for p in layers[a].parameters()]))
        y = f(y)
    y[0].sum().backward()
    assert sorted(seen) == [0, 1, 2, 3] and all(all(v) for v in seen.values()), seen


@pytest.mark.parametrize('interval', [0, 2])
def test_arming_the_callbacks_changes_no_gradient(interval):
    """the markers are aliases: outputs, input gradients and parameter gradients are bit-identical with and without them,
    also with activation checkpointing (one marker per checkpointed segment); ints / non-grad tensors pass untouched"""
    from diffusion_pipe_b200.pipe.module import PipelineModule
    import toy_model
    res = []
    for armed in (False, True):
        layers = toy_model.make_layers()
        pm = PipelineModule(layers=layers, num_stages=1, loss_fn=toy_model.loss_fn, activation_checkpoint_interval=interval,
                            device=torch.device('cpu'))
        fired = []
        pm._grads_ready_cb = (lambda a, b: fired.append((a, b))) if armed else None
        (feats, label), = toy_model.make_micro_batches(1, MBS, seed=7)
        feats = tuple(t.clone().requires_grad_(t.is_floating_point()) for t in feats)
        loss = toy_model.loss_fn(pm(feats if len(feats) > 1 else feats[0]), label)
        loss.backward()
        res.append((loss.detach(), [t.grad for t in feats if t.is_floating_point()],
                    {n: p.grad.clone() for n, p in pm.named_parameters() if p.grad is not None}, fired))
    (l0, g0, p0, f0), (l1, g1, p1, f1) = res
    assert torch.equal(l0, l1) and not f0 and f1
    assert len(g0) == len(g1) and all((a is None and b is None) or torch.equal(a, b) for a, b in zip(g0, g1))
    assert p0.keys() == p1.keys() and all(torch.equal(p0[k], p1[k]) for k in p0)
    spans = sorted(f1)
    n = len(pm.forward_funcs)
    assert spans[0][0] <= 1 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:])), spans
