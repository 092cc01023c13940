"""GPU (ONE device is enough): the CUDA-IPC stage link (pipe/ipc_link.py + csrc/p2p_ipc.cu: exported mailboxes,
cudaMemcpyPeerAsync, release/acquire sequence flags) carrying a 2- and a 3-stage pipeline whose stage processes all sit on
cuda:0.  The processes are real — separate CUDA contexts, cudaIpcGetMemHandle / cudaIpcOpenMemHandle between them, the
gloo control group — only the peer copy degenerates to a device-local copy, so the link's protocol (slot reuse, flow
control, per-step handshake, forward-only release) is exercised on the 1-GPU box the driver's GPU tier runs on.
(tests/test_pipeline_multigpu.py is the same comparison across 2 physical GPUs over NVLink.)

Compared against the ORACLE engine (oracle/engine_ref.py: single-process fp32 restatement of the DeepSpeed step, SURVEY.md
8a E3-E10) on identical micro-batches, weights and optimizer: every step's loss within 5e-3 relative (bf16 kernels vs fp32),
the clipped global gradient norm within 3e-2, the evaluation loss within 5e-3; and against the product's own 1-stage run
to 2e-3 (the partition and the transport do not change the arithmetic).  NCCL refuses two ranks on one device, so the
scalar collectives of these runs (loss broadcast, gradient-norm sum) ride on gloo."""
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
STEPS = 3
LR = 0.02


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


def _model(device):
    from diffusion_pipe_b200.flux import FluxPipeline
    torch.manual_seed(7)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'transformer_config': CFG}}, device=device)
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            if p.ndim == 1 and 'norm_' not in n:
                p.normal_(0, 0.05)
    return model


def _worker(rank, world, port, schedule, split, outdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK='0', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      DPIPE_LINK_TIMEOUT_S='120')
    sys.path.insert(0, ROOT)
    import faulthandler
    faulthandler.dump_traceback_later(420, exit=True)     # a stuck worker must not outlive the test holding a GPU
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    from diffusion_pipe_b200.pipe import ManualPipelineModule, dist, initialize
    if world > 1:
        dist.init_distributed('gloo')            # two ranks on one device: NCCL would refuse ("duplicate GPU")
    model = _model(dev)
    pm = ManualPipelineModule(layers=model.to_layers(), num_stages=world, partition_method='manual' if world > 1 else 'uniform',
                              manual_partition_split=split if world > 1 else None, loss_fn=model.get_loss_fn(), dynamic_shape=True)
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': 1, 'gradient_accumulation_steps': GAS,
                                                   'gradient_clipping': 1.0, 'steps_per_print': 0,
                                                   'stage_link': 'ipc' if world > 1 else 'dist', 'pipeline_schedule': schedule})
    engine._configure_optimizer(lambda ps: torch.optim.SGD(ps, lr=LR), [p for p in pm.parameters()])
    losses, norms = [], []
    pulls = engine.is_first_stage() or engine.is_last_stage()
    for step in range(STEPS):
        engine.reset_activation_shape()
        losses.append(float(engine.train_batch(iter(_batches(step)) if pulls else None)))
        norms.append(float(engine._grad_norm))
    ev = float(engine.eval_batch(iter(_batches(99)) if pulls else None, num_micro_batches=GAS))
    torch.save({'losses': losses, 'norms': norms, 'eval': ev, 'link': type(engine.link).__name__},
               os.path.join(outdir, f'r{rank}.pt'))
    if world > 1:
        dist.barrier()


def _run(world, schedule='1f1b', split=None):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), schedule, split, d), nprocs=world, join=True)
        return [torch.load(os.path.join(d, f'r{r}.pt'), weights_only=False) for r in range(world)]


_CACHE = {}


def _oracle():
    """the same three steps + evaluation on the fp32 oracle engine (CPU)"""
    if 'oracle' in _CACHE:
        return _CACHE['oracle']
    from oracle import flux_ref as R
    from oracle.engine_ref import RefPipelineEngine
    model = _model(torch.device('cuda', 0))
    ref = R.RefFluxTransformer(dim=256, heads=2, num_double=2, num_single=2, joint_dim=64, pooled_dim=32)
    missing, unexpected = ref.load_state_dict({k: v.detach().float().cpu() for k, v in model.transformer.state_dict().items()},
                                              strict=False)
    assert not missing and not unexpected
    del model
    ref.set_emulate_bf16(False)
    eng = RefPipelineEngine(R.to_layers(ref), R.loss_fn, torch.optim.SGD(ref.parameters(), lr=LR), None, GAS, 1.0)
    losses, norms = [], []
    for step in range(STEPS):
        losses.append(float(eng.train_batch(_batches(step))))
        norms.append(float(eng.grad_norm))
    with torch.no_grad():
        ev = sum(float(R.loss_fn(eng.forward(f), l)) for f, l in _batches(99)) / GAS
    _CACHE['oracle'] = {'losses': losses, 'norms': norms, 'eval': ev}
    return _CACHE['oracle']


def _single_stage():
    if 'single' not in _CACHE:
        _CACHE['single'] = _run(1)[0]
    return _CACHE['single']


def _check(res, tag):
    want, base = _oracle(), _single_stage()
    for r in res:
        assert r['link'] == 'IpcLink', r['link']
        for got, ref in zip(r['losses'], want['losses']):
            assert abs(got - ref) / abs(ref) <= 5e-3, (tag, r['losses'], want['losses'])
        assert abs(r['eval'] - want['eval']) / abs(want['eval']) <= 5e-3, (tag, r['eval'], want['eval'])
        for got, ref in zip(r['norms'], want['norms']):
            assert abs(got - ref) / ref <= 3e-2, (tag, r['norms'], want['norms'])
        assert r['losses'] == pytest.approx(base['losses'], rel=2e-3), (tag, r['losses'], base['losses'])
        assert r['eval'] == pytest.approx(base['eval'], rel=2e-3)
    assert all(r['losses'] == res[0]['losses'] for r in res)          # the loss is broadcast to every stage


def test_single_stage_matches_the_oracle_engine():
    want, base = _oracle(), _single_stage()
    for got, ref in zip(base['losses'], want['losses']):
        assert abs(got - ref) / abs(ref) <= 5e-3, (base['losses'], want['losses'])
    assert abs(base['eval'] - want['eval']) / abs(want['eval']) <= 5e-3


@pytest.mark.parametrize('schedule', ['1f1b', 'zb'])
def test_two_stages_over_the_ipc_link_match_the_oracle_engine(schedule):
    _check(_run(2, schedule, split=[3]), f'2 stages {schedule}')


@pytest.mark.parametrize('schedule', ['1f1b', 'zb'])
def test_three_stages_over_the_ipc_link_match_the_oracle_engine(schedule):
    """a middle stage: two inbound and two outbound channels, never touches the data"""
    _check(_run(3, schedule, split=[2, 4]), f'3 stages {schedule}')
