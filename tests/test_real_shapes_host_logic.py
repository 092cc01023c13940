"""CPU twin of tests/test_real_shapes_gpu.py: the SAME checking code (block construction, weight transfer to the oracle,
call signatures, loss, tolerances) on the PyTorch kernel doubles at small shapes, so that the GPU test itself is exercised
before it reaches a GPU box."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def doubles(monkeypatch):
    import kernel_doubles
    from diffusion_pipe_b200 import ops
    kernel_doubles.install(monkeypatch, ops)
    return ops


@pytest.mark.parametrize('kind', ['double', 'single'])
def test_flux_checker(doubles, kind):
    import test_real_shapes_gpu as T
    T.check_flux(kind, D=256, H=2, Lt=24, side=6, dev='cpu')


def test_qwen_checker(doubles):
    import test_real_shapes_gpu as T
    T.check_qwen(D=256, H=2, Lt=16, side=6, dev='cpu')


def test_wan_checker(doubles):
    import test_real_shapes_gpu as T
    T.check_wan(D=256, F=512, H=2, Lc=16, grid=(2, 4, 4), dev='cpu')


def test_eval_and_checkpointing_checkers(doubles, monkeypatch):
    """tests/test_eval_checkpoint_gpu.py on the kernel doubles"""
    import test_eval_checkpoint_gpu as T
    monkeypatch.setattr(T, 'DEVICE', 'cpu')
    T.test_eval_batch_with_quantile_timesteps_matches_the_oracle()
    T.test_activation_checkpointing_on_the_real_kernels(False)
    T.test_activation_checkpointing_on_the_real_kernels(True)
