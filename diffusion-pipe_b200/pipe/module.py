"""PipelineModule / ManualPipelineModule: the layer container the reference builds at train.py:608-617
(`ManualPipelineModule(layers=..., num_stages=..., partition_method=..., manual_partition_split=..., loss_fn=...,
dynamic_shape=True, activation_checkpoint_interval=1, checkpointable_layers=..., activation_checkpoint_func=...)`,
reference: utils/pipeline.py:11-53 on top of deepspeed.pipe.PipelineModule).

Rank layout follows DeepSpeed's PipeDataParallelTopology(axes=['pipe','data']): global_rank = stage * dp + dp_rank.
"""
import ctypes
import functools
import re

import torch
from torch import nn

from .. import _lib
from . import dist


class LayerSpec:
    """Deferred layer construction (deepspeed.runtime.pipe.LayerSpec)."""

    def __init__(self, typename, *module_args, **module_kwargs):
        self.typename = typename
        self.module_args = module_args
        self.module_kwargs = module_kwargs

    def build(self):
        return self.typename(*self.module_args, **self.module_kwargs)

    def __repr__(self):
        return f'LayerSpec({self.typename.__name__})'


class PipeTopology:
    def __init__(self, num_pp, num_dp):
        self.num_pp, self.num_dp = num_pp, num_dp

    def get_dim(self, axis):
        return {'pipe': self.num_pp, 'data': self.num_dp}[axis]

    def get_rank(self, pipe, data):
        return pipe * self.num_dp + data

    def get_coord(self, rank):
        class _C:
            pass
        c = _C()
        c.pipe, c.data = rank // self.num_dp, rank % self.num_dp
        return c


class PipelineGrid:
    """The subset of deepspeed's PipelineParallelGrid the reference touches (train.py:632,822-835; utils/saver.py:59-60,
    88-89; utils/dataset.py:1393-1398)."""

    def __init__(self, topology, global_rank, make_groups=True):
        self._topo = topology
        self.global_rank = global_rank
        self.pipe_parallel_size = topology.num_pp
        self.data_parallel_size = topology.num_dp
        c = topology.get_coord(global_rank)
        self.stage_id, self.data_parallel_id = c.pipe, c.data
        self.pp_group = [topology.get_rank(s, self.data_parallel_id) for s in range(topology.num_pp)]
        self.dp_group = [topology.get_rank(self.stage_id, d) for d in range(topology.num_dp)]
        self.pp_proc_group = None
        self.dp_proc_group = None
        if make_groups and dist.is_initialized() and dist.get_world_size() > 1:
            # every rank must create every group, in the same order
            for d in range(topology.num_dp):
                ranks = [topology.get_rank(s, d) for s in range(topology.num_pp)]
                g = dist.new_group(ranks)
                if d == self.data_parallel_id:
                    self.pp_proc_group = g
            for s in range(topology.num_pp):
                ranks = [topology.get_rank(s, d) for d in range(topology.num_dp)]
                g = dist.new_group(ranks)
                if s == self.stage_id:
                    self.dp_proc_group = g

    def get_stage_id(self):
        return self.stage_id

    def get_pipe_parallel_rank(self):
        return self.stage_id

    def get_pipe_parallel_world_size(self):
        return self.pipe_parallel_size

    def get_data_parallel_rank(self):
        return self.data_parallel_id

    def get_data_parallel_world_size(self):
        return self.data_parallel_size

    def get_global_rank(self):
        return self.global_rank

    def stage_to_global(self, stage_id):
        return self._topo.get_rank(stage_id, self.data_parallel_id)

    def get_pipe_parallel_group(self):
        return self.pp_proc_group

    def get_data_parallel_group(self):
        return self.dp_proc_group

    # DeepSpeed calls the per-pipeline group the "model parallel" group
    def get_model_parallel_group(self):
        return self.pp_proc_group

    def get_model_parallel_rank(self):
        return self.stage_id

    def get_model_parallel_world_size(self):
        return self.pipe_parallel_size


def partition_uniform(num_items, num_parts):
    """DeepSpeed ds_utils.partition_uniform: residual items go one each to the first parts."""
    parts = [0] * (num_parts + 1)
    if num_items <= num_parts:
        for p in range(num_parts + 1):
            parts[p] = min(p, num_items)
        return parts
    chunk = num_items // num_parts
    residual = num_items - chunk * num_parts
    for p in range(1, num_parts + 1):
        parts[p] = parts[p - 1] + chunk + (1 if p - 1 < residual else 0)
    return parts


def partition_balanced(weights, num_parts):
    """Contiguous min-max partition via the C++ planner (dpipe_partition_balanced)."""
    n = len(weights)
    w = (ctypes.c_int64 * n)(*[int(x) for x in weights])
    b = (ctypes.c_int * (num_parts + 1))()
    _lib.check(_lib.lib().dpipe_partition_balanced(w, n, num_parts, b), 'dpipe_partition_balanced')
    return list(b)


class _GradsReady(torch.autograd.Function):
    """Identity on a layer's inputs (aliases, no copy).  Its backward runs when the gradients of ALL its outputs are
    complete, i.e. after every backward node of the layer(s) consuming them has been launched; it calls back and hands the
    gradients on untouched.  See PipelineModule._arm_grads_ready."""

    @staticmethod
    def forward(ctx, cb, first, last, *xs):
        ctx.cb, ctx.span = cb, (first, last)
        ctx.set_materialize_grads(False)
        return tuple(t.view_as(t) for t in xs)

    @staticmethod
    def backward(ctx, *grads):
        ctx.cb(*ctx.span)
        return (None, None, None) + grads


class PipelineModule(nn.Module):
    def __init__(self, layers, num_stages=None, topology=None, loss_fn=None, seed_layers=False, seed_fn=None,
                 base_seed=1234, partition_method='parameters', activation_checkpoint_interval=0,
                 activation_checkpoint_func=None, checkpointable_layers=None, dynamic_shape=False, device=None):
        super().__init__()
        if num_stages is None and topology is None:
            raise RuntimeError('must provide num_stages or topology')
        self.loss_fn = loss_fn
        self.checkpointable_layers = checkpointable_layers
        self.activation_checkpoint_interval = activation_checkpoint_interval
        # default: the re-entrant form, like DeepSpeed's checkpointing.checkpoint (train.py passes its own, train.py:588-603)
        self.activation_checkpoint_func = activation_checkpoint_func or functools.partial(torch.utils.checkpoint.checkpoint,
                                                                                          use_reentrant=True)
        self.dynamic_shape = dynamic_shape
        self.global_rank = dist.get_rank()
        self.world_size = dist.get_world_size()
        if topology is None:
            if self.world_size % num_stages != 0:
                raise RuntimeError(f'num_stages ({num_stages}) must divide distributed world size ({self.world_size})')
            topology = PipeTopology(num_stages, self.world_size // num_stages)
        self._topo = topology
        self.num_stages = topology.get_dim('pipe')
        self.stage_id = topology.get_coord(self.global_rank).pipe
        self._grid = PipelineGrid(topology, self.global_rank)
        self._layer_specs = list(layers)
        self._num_layers = len(self._layer_specs)
        self._local_start = 0
        self._local_stop = None
        self.parts = None
        self._partition_layers(method=partition_method)
        self.forward_funcs = []
        self._build()
        if device is None:
            if torch.cuda.is_available():
                device = torch.device('cuda', torch.cuda.current_device())
            else:
                device = torch.device('cpu')
        self.to(device)

    # ---- partitioning ----
    def _count_layer_params(self):
        """all parameters, trainable or not (the reference's monkeypatch, train.py:81-90)."""
        counts = [0] * len(self._layer_specs)
        for idx, layer in enumerate(self._layer_specs):
            if isinstance(layer, LayerSpec):
                if getattr(layer, 'param_count', None) is not None:
                    counts[idx] = int(layer.param_count)      # known without materialising the layer
                else:
                    counts[idx] = sum(p.numel() for p in layer.build().parameters())
            elif isinstance(layer, nn.Module):
                counts[idx] = sum(p.numel() for p in layer.parameters())
        return counts

    def _find_layer_type(self, layername):
        idxs = []
        typeregex = re.compile(layername, re.IGNORECASE)
        for idx, layer in enumerate(self._layer_specs):
            if isinstance(layer, LayerSpec):
                name = layer.typename.__name__
            elif isinstance(layer, nn.Module):
                name = layer.__class__.__name__
            else:
                name = getattr(layer, '__name__', '')
            if typeregex.search(name):
                idxs.append(idx)
        if len(idxs) == 0:
            raise RuntimeError(f"Partitioning '{layername}' found no valid layers to partition.")
        return idxs

    def _partition_layers(self, method='uniform'):
        num_stages = self._topo.get_dim('pipe')
        stage_id = self._topo.get_coord(self.global_rank).pipe
        method = method.lower()
        if method == 'uniform':
            self.parts = partition_uniform(len(self._layer_specs), num_stages)
        elif method == 'parameters':
            self.parts = partition_balanced(self._count_layer_params(), num_stages)
        elif method.startswith('type:'):
            layertype = method.split(':')[1]
            binary = [0] * len(self._layer_specs)
            for idx in self._find_layer_type(layertype):
                binary[idx] = 1
            self.parts = partition_balanced(binary, num_stages)
        else:
            raise NotImplementedError(f'Partitioning method {method} not implemented.')
        self._print_partition()
        self._set_bounds(start=self.parts[stage_id], stop=self.parts[stage_id + 1])

    def _print_partition(self):
        if self.global_rank != 0:
            return
        for stage in range(self._topo.get_dim('pipe')):
            start, stop = self.parts[stage], self.parts[stage + 1]
            print(f'stage={stage} layers={stop - start}')
            for idx, layer in enumerate(self._layer_specs[start:stop]):
                if isinstance(layer, LayerSpec):
                    name = layer.typename.__name__
                elif isinstance(layer, nn.Module):
                    name = layer.__class__.__name__
                else:
                    name = getattr(layer, '__name__', str(layer))
                print(f'    {idx + start:2d}: {name}')
        if self.loss_fn:
            print(f"  loss: {getattr(self.loss_fn, '__name__', self.loss_fn.__class__.__name__)}")

    def _set_bounds(self, start=None, stop=None):
        self._local_start = start
        self._local_stop = stop

    def _build(self):
        for local_idx, layer in enumerate(self._layer_specs[self._local_start:self._local_stop]):
            layer_idx = local_idx + self._local_start
            if isinstance(layer, nn.Module):
                self.forward_funcs.append(layer)
                self.add_module(str(layer_idx), layer)
            elif isinstance(layer, LayerSpec):
                module = layer.build()
                self.forward_funcs.append(module)
                self.add_module(str(layer_idx), module)
            else:
                self.forward_funcs.append(layer)   # plain callable (models/chroma.py:284 style)

    # ---- execution ----
    def _is_checkpointable(self, funcs):
        if self.checkpointable_layers is not None:
            return all(f.__class__.__name__ in self.checkpointable_layers for f in funcs)
        params = [f.parameters() for f in funcs if isinstance(f, nn.Module)]
        return any(len(list(p)) > 0 for p in params)

    @staticmethod
    def _arm_grads_ready(x, first, last, cb):
        """Returns x with every tensor that takes part in autograd passed through `_GradsReady`: cb(first, last) runs in the
        backward pass as soon as the gradient of every input of layers [first, last) is complete — from then on no kernel
        of these layers' backward is still to be launched, so their parameter gradients are final for this micro-batch
        (the engine starts their data-parallel all-reduce, pipe/engine.py).

        Why not tensor hooks on x itself: a tensor the layers hand through unchanged (the time embedding, the text stream
        of the single blocks) is ONE autograd tensor consumed by every layer after it; its gradient — and with it a hook
        on it — completes only when the FIRST of these layers has run its backward, i.e. when nothing is left to overlap
        with.  The marker gives each layer its own aliases, whose only consumers are that layer and the next marker."""
        single = not isinstance(x, (tuple, list))
        xs = [x] if single else list(x)
        idx = [k for k, t in enumerate(xs) if torch.is_tensor(t) and t.requires_grad]
        if not idx:
            return x
        marked = _GradsReady.apply(cb, first, last, *[xs[k] for k in idx])
        for k, m in zip(idx, marked):
            xs[k] = m
        return xs[0] if single else tuple(xs)

    def forward(self, forward_input):
        x = forward_input
        interval = self.activation_checkpoint_interval
        ready_cb = getattr(self, '_grads_ready_cb', None) if torch.is_grad_enabled() else None
        if interval == 0 or not torch.is_grad_enabled():
            for i, f in enumerate(self.forward_funcs):
                if ready_cb is not None:
                    x = self._arm_grads_ready(x, i, i + 1, ready_cb)
                x = f(x)
            return x
        n = len(self.forward_funcs)
        for start in range(0, n, interval):
            funcs = self.forward_funcs[start:min(start + interval, n)]
            if ready_cb is not None:
                x = self._arm_grads_ready(x, start, min(start + interval, n), ready_cb)

            def run(*inputs, _funcs=funcs):
                y = inputs if len(inputs) > 1 else inputs[0]
                for f in _funcs:
                    y = f(y)
                return y
            if not isinstance(x, tuple):
                x = (x,)
            if self._is_checkpointable(funcs):
                x = self.activation_checkpoint_func(run, *x)
            else:
                x = run(*x)
        return x

    def mpu(self):
        return self._grid

    def compile(self, *args, **kwargs):
        """train.py:620-621 calls `pipeline_model.compile(dynamic=True)` when `compile = true`: accepted and ignored — the
        layers already run on hand-written kernels, and a tracing compiler has nothing to fuse across the C-ABI calls."""
        return self

    def topology(self):
        return self._topo

    def num_pipeline_stages(self):
        return self.num_stages


class ManualPipelineModule(PipelineModule):
    """utils/pipeline.py:11-53: partition_method='manual' with explicit stage boundaries."""

    def __init__(self, *args, manual_partition_split=None, **kwargs):
        self.manual_partition_split = manual_partition_split
        super().__init__(*args, **kwargs)

    def _partition_layers(self, method='uniform'):
        if method.lower() == 'manual' and self.manual_partition_split is not None:
            num_stages = self._topo.get_dim('pipe')
            stage_id = self._topo.get_coord(self.global_rank).pipe
            num_partitions = len(self.manual_partition_split)
            assert num_partitions == num_stages - 1, (
                f'partition_split must be length {num_stages - 1} (pipeline_stages-1), was actually {num_partitions}')
            self.parts = [0] + list(self.manual_partition_split) + [len(self._layer_specs)]
            self._print_partition()
            self._set_bounds(start=self.parts[stage_id], stop=self.parts[stage_id + 1])
        else:
            super()._partition_layers(method)
