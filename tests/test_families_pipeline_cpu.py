"""CPU / gloo: the Qwen-Image and Wan layer tuple protocols (bool key mask, int32 shape tensors, fp32 rope tables, empty
`None` placeholders, text context) through the pipeline engine — 2 stages x {1F1B, zero-bubble} against 1 stage — with
the kernel wrappers replaced by the PyTorch test doubles (the arithmetic is identical in both runs, so losses agree to
rounding of the boundary dtype)."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(HERE, 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)

GAS = 4


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _make(kind, device='cpu'):
    from synth import fill_parameters
    if kind == 'qwen_image':
        from diffusion_pipe_b200.qwen_image import QwenImagePipeline
        from oracle import qwen_ref as Q
        model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'device': device,
                                             'transformer_config': {'num_attention_heads': 2, 'num_layers': 2, 'joint_attention_dim': 64}}})
        ref = fill_parameters(Q.RefQwenImageTransformer(dim=256, heads=2, num_layers=2, joint_dim=64))
    else:
        from diffusion_pipe_b200.wan import WanPipeline
        from oracle import wan_ref as W
        cfg = {'dim': 256, 'ffn_dim': 512, 'num_heads': 2, 'num_layers': 2, 'text_dim': 64, 'text_len': 16}
        model = WanPipeline({'model': {'dtype': 'bfloat16', 'device': device, 'transformer_config': cfg}})
        ref = fill_parameters(W.RefWanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, text_len=16))
    sd = ref.state_dict()
    with torch.no_grad():
        for n, p in model.transformer.named_parameters():
            p.copy_(sd[n].to(p.dtype))
    return model


def _micro_batches(kind, model, n, seed):
    from diffusion_pipe_b200 import data_feed
    g = torch.Generator().manual_seed(seed)
    if kind == 'qwen_image':
        batch = {'latents': torch.randn(n, 16, 1, 8, 8, generator=g), 'prompt_embeds': [torch.randn(3 + (i * 2) % 5, 64, generator=g).bfloat16() for i in range(n)],      # padded to the longest: key mask
                 'mask': None}
    else:
        batch = {'latents': torch.randn(n, 16, 2, 8, 8, generator=g), 'text_embeddings': torch.randn(n, 16, 64, generator=g).bfloat16(),
                 'seq_lens': torch.full((n,), 12), 'mask': None}
    torch.manual_seed(seed)
    feats, label = model.prepare_inputs(batch)
    return data_feed.split_batch((feats, label), n)


def _worker(rank, world, port, kind, stages, schedule, outdir, gpu=False, link='dist', gas=GAS, save_params=False, lora=False):
    """gpu=False: CPU + gloo + kernel test doubles;  gpu=True (tests/test_pipeline_multigpu.py): one GPU per rank, NCCL,
    the real kernels and the requested stage link"""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from diffusion_pipe_b200 import ops
    from diffusion_pipe_b200.pipe import ManualPipelineModule, dist, initialize
    if gpu:
        import faulthandler
        faulthandler.dump_traceback_later(420, exit=True)     # a stuck worker must not outlive the test holding a GPU
        torch.cuda.set_device(rank)
        device = torch.device('cuda', rank)
        if world > 1:
            dist.init_distributed('nccl')
    else:
        import kernel_doubles
        kernel_doubles.install(_Patch(), ops)
        torch.set_num_threads(1)
        device = torch.device('cpu')
        if world > 1:
            dist.init_distributed('gloo')
    model = _make(kind, device)
    if lora:
        # adapter factors as a launcher without a common seed leaves them (train.py never seeds): every data-parallel
        # replica draws its own A at attach time (B starts at zero, models/base.py:272-303); the engine must make all of
        # them train replica 0's.  Nothing touches the factors between attach and engine construction, as in train.py.
        torch.manual_seed(1 + (rank % (world // stages)))
        model.configure_adapter({'type': 'lora', 'rank': 8, 'alpha': 8, 'dropout': 0.0})
    pm = ManualPipelineModule(layers=model.to_layers(), num_stages=stages, partition_method='uniform', manual_partition_split=None,
                              loss_fn=model.get_loss_fn(), dynamic_shape=True, device=device)
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': 1, 'gradient_accumulation_steps': gas,
                                                   'gradient_clipping': 1.0, 'steps_per_print': 0, 'stage_link': link,
                                                   'pipeline_schedule': schedule})
    params = [p for p in pm.parameters() if p.requires_grad]
    engine._configure_optimizer(lambda ps: torch.optim.SGD(ps, lr=0.05) if ps else None, params)
    losses = []
    for step in range(2):
        engine.reset_activation_shape()
        dp, dp_rank = engine.grid.get_data_parallel_world_size(), engine.grid.get_data_parallel_rank()
        # one global batch of gas * dp samples per step; replica r trains on samples r, r + dp, r + 2 dp, ...
        mbs = _micro_batches(kind, model, gas * dp, 100 + step)[dp_rank::dp]
        it = iter(mbs) if (engine.is_first_stage() or engine.is_last_stage()) else None
        losses.append(float(engine.train_batch(it)))
    engine.reset_activation_shape()        # train.py:181: the evaluation micro-batches have their own shapes (prompt padding)
    ev_it = iter(_micro_batches(kind, model, 2, 999)) if (engine.is_first_stage() or engine.is_last_stage()) else None
    ev = float(engine.eval_batch(ev_it, num_micro_batches=2))
    out = {'losses': losses, 'eval': ev, 'norm': float(engine._grad_norm), 'stage': engine.stage_id}
    if save_params:
        out['params'] = {p.original_name: p.detach().float().clone() for p in pm.parameters()}
    torch.save(out, os.path.join(outdir, f'rank{rank}.pt'))
    if world > 1:
        dist.barrier()


def _run(kind, stages, schedule, gpu=False, link='dist', dp=1, gas=GAS, save_params=False, lora=False):
    with tempfile.TemporaryDirectory() as d:
        port = _free_port()
        if dp > 1:
            world = stages * dp
            mp.spawn(_worker, args=(world, port, kind, stages, schedule, d, gpu, link, gas, save_params, lora), nprocs=world, join=True)
            return [torch.load(os.path.join(d, f'rank{r}.pt'), weights_only=False) for r in range(world)]
        if stages == 1 and not gpu:
            _worker(0, 1, port, kind, 1, schedule, d, False, link, gas, save_params, lora)
            import torch.distributed as tdist
            if tdist.is_initialized():
                tdist.destroy_process_group()
        else:
            mp.spawn(_worker, args=(stages, port, kind, stages, schedule, d, gpu, link, gas, save_params, lora), nprocs=stages, join=True)
        return [torch.load(os.path.join(d, f'rank{r}.pt'), weights_only=False) for r in range(stages)]


@pytest.mark.parametrize('kind', ['qwen_image', 'wan'])
def test_two_stage_pipeline_matches_single_stage(kind):
    base = _run(kind, 1, '1f1b')[0]
    assert all(v == v and 0 < v < 100 for v in base['losses'])
    for schedule in ('1f1b', 'zb'):
        for r in _run(kind, 2, schedule):
            assert r['losses'] == pytest.approx(base['losses'], rel=2e-3), (kind, schedule, r['losses'], base['losses'])
            assert r['eval'] == pytest.approx(base['eval'], rel=2e-3)
            assert r['norm'] == pytest.approx(base['norm'], rel=2e-2)


def test_qwen_two_stages_times_two_replicas_equals_one_pipeline_on_the_same_global_batch():
    """BASELINE.json configs[4] shape (Qwen-Image, pipeline x data parallel) on the real model code: the gradients of a
    model whose q/k/v (and Wan's k/v) gradients live in FUSED buffers are all-reduced over the replicas of each stage,
    then clipped by the global norm — losses, norm and every updated parameter equal the 2-stage run that sees the same
    8 samples per step as 8 micro-batches (zero-bubble order in both)"""
    one = _run('qwen_image', 2, 'zb', gas=8, save_params=True)
    two = _run('qwen_image', 2, 'zb', dp=2, gas=4, save_params=True)
    assert [r['stage'] for r in two] == [0, 0, 1, 1]
    for r in two:
        assert r['losses'] == pytest.approx(one[0]['losses'], rel=2e-3)
        assert r['norm'] == pytest.approx(one[0]['norm'], rel=2e-2)
        ref = one[r['stage']]['params']
        assert set(ref) == set(r['params'])
        for k, v in r['params'].items():
            assert (v - ref[k]).norm() <= 2e-2 * (ref[k].norm() + 1e-6) + 1e-4, k
    # replicas of a stage hold identical parameters after the step
    for k, v in two[0]['params'].items():
        assert torch.equal(v, two[1]['params'][k]), k


def test_lora_replicas_train_the_first_replicas_factors():
    """utils/patches.py:163-172 with an adapter: two data-parallel replicas that drew DIFFERENT LoRA factors equal one
    pipeline holding replica 0's factors on the same global batch — from the first step on (the broadcast reaches the fused
    [[W|B],[A|0]] site buffers of lora.py, not only the parameter tensors), only adapter parameters move, and both
    replicas end with identical ones"""
    one = _run('qwen_image', 1, '1f1b', gas=8, save_params=True, lora=True)[0]
    two = _run('qwen_image', 1, '1f1b', dp=2, gas=4, save_params=True, lora=True)
    trained = [k for k in one['params'] if '.lora_' in k]
    assert trained
    for r in two:
        assert r['losses'] == pytest.approx(one['losses'], rel=2e-3), (r['losses'], one['losses'])
        assert r['norm'] == pytest.approx(one['norm'], rel=2e-2)
        for k in trained:
            assert (r['params'][k] - one['params'][k]).norm() <= 2e-2 * (one['params'][k].norm() + 1e-6) + 1e-4, k
    for k, v in two[0]['params'].items():
        assert torch.equal(v, two[1]['params'][k]), k
