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
