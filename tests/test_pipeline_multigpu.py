"""GPU (>= 2 devices; skipped otherwise): the same small Flux model trained for two optimizer steps on 1 stage and on
2 pipeline stages, over both stage links (torch.distributed p2p and the CUDA-IPC peer-copy link), must give the same
losses — the partition and the transport do not change the arithmetic (same kernels, same order per micro-batch)."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CFG = {'num_attention_heads': 2, 'num_layers': 2, 'num_single_layers': 2, 'joint_attention_dim': 64,
       'pooled_projection_dim': 32}
GAS = 4


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batches(step):
    from oracle import flux_ref as R
    out = []
    for i in range(GAS):
        g = torch.Generator().manual_seed(1000 * step + i)
        feats, (target, _) = R.prepare_inputs(torch.randn(1, 16, 16, 16, generator=g), torch.randn(1, 32, 64, generator=g).bfloat16(),
                                              torch.randn(1, 32, generator=g).bfloat16(), torch.sigmoid(torch.randn(1, generator=g)),
                                              torch.randn(1, 16, 16, 16, generator=g))
        out.append((feats, (target, torch.tensor([]))))
    return out


def _worker(rank, world, port, link, outdir, schedule='1f1b'):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import faulthandler
    faulthandler.dump_traceback_later(420, exit=True)     # a stuck worker must not outlive the test holding a GPU
    torch.cuda.set_device(rank)
    from diffusion_pipe_b200.flux import FluxPipeline
    from diffusion_pipe_b200.pipe import ManualPipelineModule, dist, initialize
    if world > 1:
        dist.init_distributed('nccl')
    torch.manual_seed(7)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': CFG}},
                         device=torch.device('cuda', rank))
    pm = ManualPipelineModule(layers=model.to_layers(), num_stages=world, partition_method='manual' if world > 1 else 'uniform',
                              manual_partition_split=[3] if world > 1 else None, loss_fn=model.get_loss_fn(), dynamic_shape=True)
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': 1, 'gradient_accumulation_steps': GAS,
                                                   'gradient_clipping': 1.0, 'steps_per_print': 0, 'stage_link': link,
                                                   'pipeline_schedule': schedule})
    engine._configure_optimizer(lambda ps: torch.optim.SGD(ps, lr=0.02), [p for p in pm.parameters()])
    losses = []
    for step in range(3):
        engine.reset_activation_shape()
        it = iter(_batches(step)) if (engine.is_first_stage() or engine.is_last_stage()) else None
        losses.append(float(engine.train_batch(it)))
    ev = float(engine.eval_batch(iter(_batches(99)) if (engine.is_first_stage() or engine.is_last_stage()) else None,
                                 num_micro_batches=GAS))
    torch.save({'losses': losses, 'eval': ev, 'link': type(engine.link).__name__}, os.path.join(outdir, f'r{rank}.pt'))
    if world > 1:
        dist.barrier()


def _run(world, link, schedule='1f1b'):
    with tempfile.TemporaryDirectory() as d:
        port = _free_port()
        mp.spawn(_worker, args=(world, port, link, d, schedule), nprocs=world, join=True)
        return [torch.load(os.path.join(d, f'r{r}.pt'), weights_only=False) for r in range(world)]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('link', ['dist', 'ipc'])
def test_two_stage_pipeline_matches_single_stage(link):
    base = _run(1, 'dist')[0]
    res = _run(2, link)
    assert res[0]['link'] == ('IpcLink' if link == 'ipc' else 'DistLink')
    for r in res:
        assert r['losses'] == pytest.approx(base['losses'], rel=2e-3), (r['losses'], base['losses'])
        assert r['eval'] == pytest.approx(base['eval'], rel=2e-3)
    assert res[0]['losses'] == res[1]['losses']     # the loss is broadcast to every stage


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('link', ['ipc', 'dist'])
def test_zero_bubble_schedule_matches_single_stage(link):
    """split backward (deferred weight gradients) + one-sided sends: same losses as the fused 1-stage run"""
    base = _run(1, 'dist')[0]
    zb1 = _run(1, 'dist', 'zb')[0]                       # deferral alone, no pipeline
    assert zb1['losses'] == pytest.approx(base['losses'], rel=1e-6)
    res = _run(2, link, 'zb')
    for r in res:
        assert r['losses'] == pytest.approx(base['losses'], rel=2e-3), (r['losses'], base['losses'])
        assert r['eval'] == pytest.approx(base['eval'], rel=2e-3)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('kind', ['qwen_image', 'wan'])
def test_qwen_and_wan_tuples_cross_the_ipc_link(kind):
    """the other families' boundary tuples (bool key mask, int32 / int64 shape tensors, fp32 rope tables, text context)
    over CUDA-IPC mailboxes, 1F1B and zero-bubble, against one stage (CPU twin: tests/test_families_pipeline_cpu.py)"""
    sys.path.insert(0, HERE)
    import test_families_pipeline_cpu as F
    base = F._run(kind, 1, '1f1b', gpu=True)[0]
    for schedule in ('1f1b', 'zb'):
        for r in F._run(kind, 2, schedule, gpu=True, link='ipc'):
            assert r['losses'] == pytest.approx(base['losses'], rel=2e-3), (kind, schedule, r['losses'], base['losses'])
            assert r['eval'] == pytest.approx(base['eval'], rel=2e-3)
