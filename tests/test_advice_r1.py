"""Regression tests for the round-1 advisor findings that are not covered next to the code they touch
(data-loader state: tests/test_data_feed.py)."""
import time

import torch

from diffusion_pipe_b200 import saver as S
from diffusion_pipe_b200.pipe import dist


def test_checkpoint_every_n_minutes_in_a_single_process_run():
    """a 1-GPU run never initialises torch.distributed (train.py: only when WORLD_SIZE > 1); every reference example sets
    `checkpoint_every_n_minutes` (examples/main_example.toml), so the minute clock must work without a process group"""
    assert not dist.is_initialized()
    sv = S.Saver(args=None, config={'checkpoint_every_n_minutes': 1e-9}, is_adapter=False, save_root='/tmp/unused', model=None,
                 train_dataloader=None, model_engine=None, pipeline_model=None)
    assert sv.need_to_checkpoint() is False          # first call only starts the clock
    time.sleep(0.01)
    assert sv.need_to_checkpoint() is True
    holder = ['x']
    dist.broadcast_object_list(holder, src=0)
    assert holder == ['x']
    seen = [None]
    dist.all_gather_object(seen, 5)
    assert seen == [5]


def test_in_place_broadcast_moves_the_parameter_version():
    """engine._broadcast_model writes the parameter itself (not `.data`): caches keyed on `p._version` (lora.py site buffers)
    must see the update.  Same two statements as the engine, on a one-process gloo group."""
    import os
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK='0', WORLD_SIZE='1')
    torch.distributed.init_process_group('gloo', rank=0, world_size=1)
    try:
        p = torch.nn.Parameter(torch.randn(4, 4))
        v0 = p._version
        with torch.no_grad():
            dist.broadcast(p, 0)
            p.add_(0)
        assert p._version > v0
    finally:
        torch.distributed.destroy_process_group()


def test_deferred_scalars_log_the_same_values_one_step_later():
    """train.py: the per-step `.item()` is replaced by an asynchronous copy that is written to the log after the
    next step has been enqueued — same tags, values and x axis, in order; flush() drains everything."""
    import train
    seen = []
    sc = train.DeferredScalars(lambda tag, value, x: seen.append((tag, round(float(value), 6), x)))
    sc.push('train/loss', torch.tensor(1.5), 1)
    sc.push('train/grad_norm', 0.25, 1)
    assert seen == []                                   # nothing is written in the step that produced it
    sc.flush(keep=0)
    sc.push('train/loss', torch.tensor(1.25), 2)
    assert seen == [('train/loss', 1.5, 1), ('train/grad_norm', 0.25, 1)]
    sc.flush(keep=1)
    assert len(seen) == 2
    sc.flush(keep=0)
    assert seen[-1] == ('train/loss', 1.25, 2) and sc.pending == []


def test_clip_sees_every_gradient_byte_once_through_fused_buffers():
    """engine._clip_grad_norm sums the fused gradient allocation behind per-projection views once (flux_blocks.FusedParam) and
    falls back to the per-tensor path when the views do not cover it"""
    from diffusion_pipe_b200.pipe.engine import PipelineEngine as E
    base = torch.arange(12, dtype=torch.float32)
    a, b = torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(2, 3))
    a.grad, b.grad = base[:6].view(2, 3), base[6:].view(2, 3)
    c = torch.nn.Parameter(torch.zeros(4))
    c.grad = torch.ones(4)
    bufs = E._unique_grad_buffers([a, b, c])
    assert len(bufs) == 2 and bufs[0].data_ptr() == base.data_ptr() and bufs[0].numel() == 12
    assert E._fused_views_cover([a, b, c], bufs)
    assert not E._fused_views_cover([a, c], E._unique_grad_buffers([a, c]))      # half of the fused buffer is not a gradient here
