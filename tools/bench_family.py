"""Throughput of the other model families of BASELINE.json (configs[3] Wan2.1-14B t2v, configs[4] Qwen-Image) through the
same engine path as bench.py — not the driver's bench line (that is Flux, bench.py), a measuring tool for DESIGN.md.

    python tools/bench_family.py --model wan --steps 3 --warmup 2                       # 1 GPU, reduced depth with --layers
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_family.py --model wan --stages 8
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_family.py --model qwen_image --stages 2   # pp2 x dp4

Prints one JSON line on rank 0: samples/s device-timed (max over ranks) with resident micro-batches, the same through
pinned-host micro-batches + loss read-back (e2e), per-kernel-family shares from CUDA events, peak memory, clocks.
World size = stages x data-parallel replicas (NCCL gradient all-reduce when replicas > 1).
Algorithmic training FLOPs per sample: Wan-14B 33f 512^2: 887.8 T; Qwen-Image 1024^2, 256 text tokens: 219.4 T.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', required=True, choices=['wan', 'qwen_image'])
    ap.add_argument('--stages', type=int, default=0, help='pipeline stages (default: world size)')
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--micro-batches', type=int, default=16)
    ap.add_argument('--layers', type=int, default=0, help='transformer blocks (default: the full model: 40 / 60)')
    ap.add_argument('--res', type=int, default=0, help='pixels (default 512 for wan, 1024 for qwen_image)')
    ap.add_argument('--frames', type=int, default=33)
    ap.add_argument('--text-len', type=int, default=0, help='prompt tokens (default 512 for wan, 256 for qwen_image)')
    ap.add_argument('--schedule', default='auto', choices=['auto', '1f1b', 'zb'])
    ap.add_argument('--adapter-rank', type=int, default=0, help='train LoRA adapters of this rank instead of the full model')
    return ap.parse_args()


def block_split(n_blocks, stages):
    """[first, blocks..., last] -> stage boundaries with the extra blocks on the earliest stages (see bench.py)"""
    base, extra = divmod(n_blocks, stages)
    per_stage = [base + (1 if s < extra else 0) for s in range(stages)]
    bounds, acc = [], 1
    for s in range(stages - 1):
        acc += per_stage[s]
        bounds.append(acc)
    return bounds, per_stage


def main():
    a = parse()
    import torch
    import torch.distributed as tdist
    import bench
    from diffusion_pipe_b200 import data_feed, ops
    from diffusion_pipe_b200.pipe import ManualPipelineModule, dist, initialize

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_distributed('nccl')
    stages = a.stages or world
    assert world % stages == 0, f'world size {world} is not a multiple of --stages {stages}'
    dp = world // stages
    M = a.micro_batches

    torch.manual_seed(1234 + rank)
    if a.model == 'wan':
        from diffusion_pipe_b200.wan import WAN_T2V_14B_CONFIG, WanPipeline
        n_blocks = a.layers or WAN_T2V_14B_CONFIG['num_layers']
        res, text_len = a.res or 512, a.text_len or 512
        model = WanPipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': {'num_layers': n_blocks}}}, device=device)
        full_blocks, tflop_full = 40, 887.8
        lat_f = (a.frames - 1) // 4 + 1

        def example(g):
            return {'latents': torch.randn(16, lat_f, res // 8, res // 8, generator=g),
                    'text_embeddings': torch.randn(512, 4096, generator=g).bfloat16(), 'seq_lens': torch.tensor(text_len), 'mask': None}
        workload = f'Wan2.1-14B t2v full fine-tune bf16, {a.frames} frames {res}x{res}, {n_blocks} blocks'
    else:
        from diffusion_pipe_b200.qwen_image import QWEN_IMAGE_CONFIG, QwenImagePipeline
        n_blocks = a.layers or QWEN_IMAGE_CONFIG['num_layers']
        res, text_len = a.res or 1024, a.text_len or 256
        model = QwenImagePipeline({'model': {'dtype': 'bfloat16', 'lazy_layers': True, 'transformer_config': {'num_layers': n_blocks}}}, device=device)
        full_blocks, tflop_full = 60, 219.4

        def example(g):
            return {'latents': torch.randn(16, 1, res // 8, res // 8, generator=g),
                    'prompt_embeds': torch.randn(text_len, 3584, generator=g).bfloat16(), 'mask': None}
        workload = f'Qwen-Image full fine-tune bf16, {res}x{res}, {text_len} text tokens, {n_blocks} blocks'
    if a.adapter_rank:
        model.configure_adapter({'type': 'lora', 'rank': a.adapter_rank, 'alpha': a.adapter_rank, 'dropout': 0.0})
        workload = workload.replace('full fine-tune', f'LoRA rank {a.adapter_rank}')
    split, per_stage = block_split(n_blocks, stages)
    pm = ManualPipelineModule(layers=model.to_layers(), num_stages=stages, partition_method='manual' if stages > 1 else 'uniform',
                              manual_partition_split=split if stages > 1 else None, loss_fn=model.get_loss_fn(), dynamic_shape=True)
    schedule = ('zb' if stages > 2 else '1f1b') if a.schedule == 'auto' else a.schedule
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': 1, 'gradient_accumulation_steps': M,
                                                   'gradient_clipping': 1.0, 'steps_per_print': 0, 'pipeline_schedule': schedule,
                                                   'zb_stage_weights': [max(1, b) for b in per_stage]})
    params = [p for p in pm.parameters() if p.requires_grad]
    engine._configure_optimizer(lambda ps: torch.optim.AdamW(ps, lr=1e-5, betas=(0.9, 0.99), weight_decay=0.01, fused=True) if ps else None,
                                params)

    need_data = engine.is_first_stage() or engine.is_last_stage()

    def micro_batches(pinned):
        if not need_data:
            return None
        g = torch.Generator().manual_seed(1234 + engine.grid.get_data_parallel_rank())
        torch.manual_seed(99 + engine.grid.get_data_parallel_rank())
        batch = data_feed.BatchedDataset.collate([example(g) for _ in range(M)])
        feats, label = model.prepare_inputs(batch)
        out = []
        for f, l in data_feed.split_batch((feats, label), M):
            mv = (lambda t: t.pin_memory()) if pinned else (lambda t: t.to(device))
            out.append((tuple(mv(t) for t in f), tuple(mv(t) for t in l)))
        return out
    dev_b, host_b = micro_batches(False), micro_batches(True)
    h2d = sum(t.numel() * t.element_size() for f, l in (host_b or []) for t in ((f if engine.is_first_stage() else ()) + (l if engine.is_last_stage() else ())))

    def step(batches, read_loss):
        engine.reset_activation_shape()
        loss = engine.train_batch(iter(batches) if batches is not None else None)
        return loss.item() if read_loss else loss

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(batches, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(a.steps):
            last = step(batches, read_loss)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return ms.item(), float(last)

    for _ in range(a.warmup):
        step(dev_b, False)
    clocks = bench.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ops.PROFILE = []
    ms_dev, loss = timed(dev_b, False)
    prof, ops.PROFILE = ops.PROFILE, None
    ms_e2e, _ = timed(host_b, True)
    clk = clocks.stop() if rank == 0 else None
    torch.cuda.synchronize()
    kind_ms, kind_fl = {}, {}
    for s_, e_, f_, k_, _t in prof:
        kind_ms[k_] = kind_ms.get(k_, 0.0) + s_.elapsed_time(e_)
        kind_fl[k_] = kind_fl.get(k_, 0.0) + f_
    mem = torch.tensor([torch.cuda.max_memory_allocated(device) / 2 ** 30], device=device, dtype=torch.float64)
    h2d_t = torch.tensor([float(h2d)], device=device, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(mem, op=tdist.ReduceOp.MAX)
        tdist.all_reduce(h2d_t)
    if rank == 0:
        samples = M * dp * a.steps
        value = samples / (ms_dev / 1000.0)
        tflop = tflop_full * n_blocks / full_blocks
        gemm_ms = kind_ms.get('gemm', 0.0)
        attn_ms = sum(v for k, v in kind_ms.items() if k.startswith('attn'))
        out = {'metric': 'training samples/sec (device-timed, max over ranks)', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
               'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms_dev / a.steps, 'higher_is_better': True,
               'scaling': 'strong' if dp == 1 else 'weak in dp', 'dtype': 'bf16', 'data': 'synthetic',
               'config': {'workload': workload, 'global_batch': M * dp, 'micro_batch': 1, 'micro_batches': M,
                          'parallelism': f'pp{stages} x dp{dp}', 'pipeline_schedule': engine.pipeline_schedule,
                          'stage_link': type(engine.link).__name__, 'train_tflop_per_sample': tflop},
               'e2e': {'value': samples / (ms_e2e / 1000.0), 'unit': 'samples/s', 'h2d_bytes_per_step': int(h2d_t.item()),
                       'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / a.steps},
               'loss': loss, 'peak_mem_gib_max_rank': round(float(mem.item()), 1), 'clocks': clk,
               'step_tflops_per_gpu': tflop * value / world,
               'rank0': {'gemm_tflops': kind_fl.get('gemm', 0.0) / gemm_ms / 1e9 if gemm_ms else None,
                         'attn_tflops_algorithmic': sum(v for k, v in kind_fl.items() if k.startswith('attn')) / attn_ms / 1e9 if attn_ms else None,
                         'share_by_kernel': {k: round(v / ms_dev, 4) for k, v in sorted(kind_ms.items(), key=lambda kv: -kv[1])}}}
        print(json.dumps(out), flush=True)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == '__main__':
    main()
