#!/usr/bin/env python
"""bench.py — training samples/sec for Flux-dev 1024x1024 bf16 full fine-tune on N pipeline stages (N GPUs of one node).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU restatement of the reference path, timed on host cores

One "step" = one optimizer step of the reference's hot path (train.py:915-918): GAS micro-batches of forward+backward
through all 59 pipeline layers, gradient clipping, AdamW, on synthetic latents / text embeddings of the named shape
(BASELINE.json configs[1..2]: Flux-dev full fine-tune, bf16, 1024x1024 -> 4096 image + 512 text tokens, micro-batch 1,
16 micro-batches).  Weights are random-initialised at the real architecture (19 double + 38 single blocks, 11.9 B
parameters).  Every stage keeps its bf16 activations (no recompute): 3x forward FLOPs per step, 223.2 TFLOP/sample.

Prints ONE JSON line (rank 0).  `value` = samples/s with the micro-batches resident in HBM; `e2e` = the same through
engine.train_batch with HOST (pinned) micro-batches and a device->host read of the loss inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_TFLOP_PER_BLOCK = 1.3046          # 1.0437 GEMM + 0.2609 attention, L=4608, D=3072
TRAIN_TFLOP_PER_SAMPLE = 223.2        # 57 blocks x 1.305 x 3 (forward + backward, no recompute)
METRIC = 'training samples/sec (device-timed, max over stages) Flux-dev 1024^2 bf16'   # BASELINE.json:metric
GEMM_FRACTION = 0.80


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='dpipe', choices=['dpipe', 'reference'])
    ap.add_argument('--micro-batches', type=int, default=16)
    ap.add_argument('--micro-batch-size', type=int, default=1)
    ap.add_argument('--res', type=int, default=1024)
    ap.add_argument('--text-len', type=int, default=512)
    ap.add_argument('--layers', type=str, default='19,38', help='double,single block counts (default = Flux-dev)')
    ap.add_argument('--no-optimizer', action='store_true', help='diagnostic: skip the optimizer step (INVALID as a bench value)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-library-baseline', action='store_true',
                    help='skip the cuBLAS + SDPA timing of the restated reference blocks on the GPU (context only, N=1)')
    ap.add_argument('--schedule', default='auto', choices=['auto', '1f1b', 'zb'],
                    help="pipeline order: the reference's 1F1B or the split-backward zero-bubble order (auto: zb from 3 stages up; at 2 stages the 1F1B bubble is 1/17 of the step and measured no worse)")
    ap.add_argument('--profile-kernels', action='store_true', default=True)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle restatement of the reference blocks on host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(res, text_len, n_double, n_single, reps=1):
    """One double + one single block forward+backward at full shapes in fp32 on all host cores; returns
    (samples_per_sec extrapolated to the whole model, description, cores)."""
    import torch
    from oracle import flux_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D, H = 3072, 24
    Li, Lt = (res // 16) ** 2, text_len
    torch.manual_seed(0)
    ids = torch.zeros(Lt + Li, 3)
    ids[Lt:, 1] = torch.arange(Li) // (res // 16)
    ids[Lt:, 2] = torch.arange(Li) % (res // 16)
    cos, sin = R.flux_rope_tables(ids)
    times = {}
    for kind, cls in (('double', R.RefFluxTransformerBlock), ('single', R.RefFluxSingleTransformerBlock)):
        blk = cls(D, H)
        hid = torch.randn(1, Li, D, requires_grad=True)
        enc = torch.randn(1, Lt, D, requires_grad=True)
        temb = torch.randn(1, D, requires_grad=True)
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            e, h = blk(hid, enc, temb, (cos, sin))
            (h.sum() + e.sum()).backward()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        times[kind] = best
        del blk
    per_sample = n_double * times['double'] + n_single * times['single']
    desc = (f'oracle/flux_ref.py blocks, fp32, torch CPU, {cores} threads: 1 double ({times["double"]:.2f}s) + 1 single '
            f'({times["single"]:.2f}s) block fwd+bwd at L={Li + Lt}, D=3072, extrapolated x{n_double}/x{n_single}')
    return 1.0 / per_sample, desc, cores


def gpu_library_sample(res, text_len, n_double, n_single, device, iters=3):
    """what the reference's blocks would dispatch on this GPU — the restated blocks (oracle/flux_ref.py) under
    torch.autocast(bf16), i.e. cuBLAS GEMMs, ATen elementwise kernels and torch SDPA (flash) — one double + one single
    block forward+backward at the full shape, CUDA events, extrapolated to the whole model.  No activation checkpointing
    (the reference adds a forward recompute on top when `activation_checkpointing = true`).  Context for the bench line, not
    part of it; any failure here is reported as a string and never takes the bench line down."""
    import torch
    import torch.nn.functional as F
    from oracle import flux_ref as R
    D, H = 3072, 24
    Li, Lt = (res // 16) ** 2, text_len
    ids = torch.zeros(Lt + Li, 3, device=device)
    ids[Lt:, 1] = torch.arange(Li, device=device) // (res // 16)
    ids[Lt:, 2] = torch.arange(Li, device=device) % (res // 16)
    cos, sin = R.flux_rope_tables(ids)

    def library_sdpa(q, k, v, emulate):
        o = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3).to(torch.bfloat16), k.permute(0, 2, 1, 3).to(torch.bfloat16),
                                           v.permute(0, 2, 1, 3).to(torch.bfloat16))
        return o.permute(0, 2, 1, 3)
    saved = R.sdpa
    R.sdpa = library_sdpa
    times = {}
    try:
        for kind, cls in (('double', R.RefFluxTransformerBlock), ('single', R.RefFluxSingleTransformerBlock)):
            blk = cls(D, H).to(device)
            hid = torch.randn(1, Li, D, device=device, requires_grad=True)
            enc = torch.randn(1, Lt, D, device=device, requires_grad=True)
            temb = torch.randn(1, D, device=device, requires_grad=True)

            def once():
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    e, h = blk(hid, enc, temb, (cos, sin))
                (h.float().sum() + e.float().sum()).backward()
            once()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                once()
            e1.record()
            torch.cuda.synchronize()
            times[kind] = e0.elapsed_time(e1) / iters / 1000.0
            del blk, hid, enc, temb
    finally:
        R.sdpa = saved
    per_sample = n_double * times['double'] + n_single * times['single']
    return {'value': 1.0 / per_sample, 'unit': 'samples/s (blocks only, extrapolated)',
            'kind': 'restated reference blocks on the GPU: cuBLAS + ATen + torch SDPA under autocast(bf16), no recompute',
            'sample': f'1 double ({times["double"] * 1e3:.1f} ms) + 1 single ({times["single"] * 1e3:.1f} ms) block fwd+bwd at L={Li + Lt}, '
                      f'extrapolated x{n_double}/x{n_single}'}


def run_reference_arm(a):
    """--impl reference: the CPU restatement of the reference path (its DeepSpeed+diffusers stack cannot be installed in
    this image: deepspeed / diffusers / peft are absent from /opt/wheelhouse) on the host cores."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    n_double, n_single = (int(x) for x in a.layers.split(','))
    # one sample = 1 double + 1 single block forward+backward at the full shape (~40 s on 128 threads): the arm is bounded
    # to one untimed and at most two timed samples whatever --warmup / --steps say, so that it ends within ~2 minutes
    vals = []
    n_warm, n_timed = min(a.warmup, 1), max(1, min(a.steps, 2))
    for i in range(n_warm + n_timed):
        t0 = time.perf_counter()
        v, desc, cores = cpu_reference_sample(a.res, a.text_len, n_double, n_single)
        if i >= n_warm:
            vals.append((v, time.perf_counter() - t0))
    value = sum(v for v, _ in vals) / len(vals)
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': value,
        'unit': 'samples/s', 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': 1000.0 * a.micro_batches * a.micro_batch_size / value, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'Flux-dev full fine-tune bf16 {a.res}x{a.res} (configs[1]/[2] model), {n_double}+{n_single} blocks',
                   'global_batch': a.micro_batches * a.micro_batch_size, 'micro_batch': a.micro_batch_size,
                   'micro_batches': a.micro_batches, 'seq_len': (a.res // 16) ** 2 + a.text_len, 'parallelism': 'host cores',
                   'arithmetic': 'fp32 restatement of the reference path (oracle/flux_ref.py); the reference itself cannot be installed here'},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                         'sample': desc + f'; {len(vals)} timed sample(s) after {n_warm} untimed'},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def gemm_traffic_from_profile():
    """dram bytes (read + write) per launch of the dominant GEMM shape, from the committed ncu --set full capture"""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_gemm_traffic.json')) as f:
            return json.load(f)['dram_bytes_total']
    except Exception:
        return None


def flop_balanced_split(n_double, n_single, stages):
    """Stage boundaries over [embed, double..., single..., out]: every block costs the same FLOPs, the
    embedding rides with the first block and the output layer with the last.  When the block count does not divide, the
    EARLIEST stages take the extra block: a heavier first stage starts at t = 0 and fills its waits with deferred
    weight-gradient work, a heavier last stage delays every backward pass (simulated speed-up at 8 stages, 16
    micro-batches: 7.1x vs 6.3x, csrc/sched.cpp).  Returns (split, blocks_per_stage)."""
    nblk = n_double + n_single
    base, extra = divmod(nblk, stages)
    per_stage = [base + (1 if s < extra else 0) for s in range(stages)]
    bounds, acc = [], 1          # layer 0 is the embedding
    for s in range(stages - 1):
        acc += per_stage[s]
        bounds.append(acc)
    return bounds, per_stage


def synth_micro_batches(a, n, seed, device, pinned):
    import torch
    g = torch.Generator().manual_seed(seed)
    bs = a.micro_batch_size
    h = a.res // 8
    Li = (h // 2) ** 2
    out = []
    for _ in range(n):
        latents = torch.randn(bs, 16, h, h, generator=g)
        noise = torch.randn(bs, 16, h, h, generator=g)
        t = torch.sigmoid(torch.randn(bs, generator=g))
        te = t.view(-1, 1, 1, 1)
        x_t = ((1 - te) * latents + te * noise).view(bs, 16, h // 2, 2, h // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(bs, Li, 64)
        target = (noise - latents).view(bs, 16, h // 2, 2, h // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(bs, Li, 64)
        t5 = torch.randn(bs, a.text_len, 4096, generator=g).bfloat16()
        clip = torch.randn(bs, 768, generator=g).bfloat16()
        ids = torch.zeros(h // 2, h // 2, 3)
        ids[..., 1] += torch.arange(h // 2)[:, None]
        ids[..., 2] += torch.arange(h // 2)[None, :]
        img_ids = ids.reshape(1, Li, 3).repeat(bs, 1, 1)
        txt_ids = torch.zeros(bs, a.text_len, 3)
        guidance = torch.full((bs,), 1.0)
        img_seq_len = torch.tensor(Li).repeat(bs)
        feats = (x_t.contiguous(), t5, clip, t, img_ids, txt_ids, guidance, img_seq_len)
        label = (target.contiguous(), torch.tensor([]))
        if pinned:
            feats = tuple(x.pin_memory() for x in feats)
            label = (label[0].pin_memory(), label[1])
        else:
            feats = tuple(x.to(device) for x in feats)
            label = (label[0].to(device), label[1])
        out.append((feats, label))
    return out


def main():
    a = parse()
    if a.impl == 'reference':
        return run_reference_arm(a)

    import torch
    import torch.distributed as tdist
    from diffusion_pipe_b200 import ops
    from diffusion_pipe_b200.flux import FluxPipeline
    from diffusion_pipe_b200.pipe import ManualPipelineModule, initialize
    from diffusion_pipe_b200.pipe import dist

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit(f'--gpus {a.gpus} needs `python -m torch.distributed.run --nproc-per-node {a.gpus} bench.py ...`')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_distributed('nccl')
    n_double, n_single = (int(x) for x in a.layers.split(','))
    stages = world
    M, mbs = a.micro_batches, a.micro_batch_size

    torch.manual_seed(1234 + rank)
    model = FluxPipeline({'model': {'dtype': 'bfloat16', 'guidance': 1.0, 'lazy_layers': True,
                                    'transformer_config': {'num_layers': n_double, 'num_single_layers': n_single}}},
                         device=device)
    layers = model.to_layers()
    split, blocks_per_stage = flop_balanced_split(n_double, n_single, stages)
    pm = ManualPipelineModule(layers=layers, num_stages=stages, partition_method='manual' if stages > 1 else 'uniform',
                              manual_partition_split=split if stages > 1 else None, loss_fn=model.get_loss_fn(),
                              dynamic_shape=True)
    engine, _, _, _ = initialize(model=pm, config={'train_micro_batch_size_per_gpu': mbs, 'gradient_accumulation_steps': M,
                                                   'gradient_clipping': 1.0, 'steps_per_print': 0,
                                                   'pipeline_schedule': ('zb' if stages > 2 else '1f1b') if a.schedule == 'auto' else a.schedule,
                                                   'zb_stage_weights': [max(1, b) for b in blocks_per_stage]})
    params = [p for p in pm.parameters() if p.requires_grad]
    if not a.no_optimizer:
        engine._configure_optimizer(lambda ps: torch.optim.AdamW(ps, lr=1e-5, betas=(0.9, 0.99), weight_decay=0.01,
                                                                 fused=True), params)
    n_params = sum(p.numel() for p in params)

    need_data = engine.is_first_stage() or engine.is_last_stage()
    dev_batches = synth_micro_batches(a, M, 1234 + rank * 0, device, pinned=False) if need_data else None
    host_batches = synth_micro_batches(a, M, 1234 + rank * 0, device, pinned=True) if need_data else None

    h2d_local = 0
    if need_data:
        for feats, label in host_batches:
            if engine.is_first_stage():
                h2d_local += sum(x.numel() * x.element_size() for x in feats)
            if engine.is_last_stage():
                h2d_local += sum(x.numel() * x.element_size() for x in label)
    h2d_t = torch.tensor([h2d_local], device=device, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(h2d_t)

    def step(batches, read_loss):
        engine.reset_activation_shape()
        loss = engine.train_batch(iter(batches) if batches is not None else None)
        return loss.item() if read_loss else loss

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(batches, read_loss, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(nsteps):
            last = step(batches, read_loss)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return ms.item(), float(last)

    for _ in range(a.warmup):
        step(dev_batches, False)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ops.LAUNCHES
    if a.profile_kernels:
        ops.PROFILE = []
    ms_dev, loss_dev = timed(dev_batches, False, a.steps)
    prof = ops.PROFILE
    ops.PROFILE = None
    launches = ops.LAUNCHES - launches0
    ms_e2e, loss_e2e = timed(host_batches, True, a.steps)
    clk = clocks.stop() if rank == 0 else None

    samples_per_step = mbs * M
    value = samples_per_step * a.steps / (ms_dev / 1000.0)
    e2e_value = samples_per_step * a.steps / (ms_e2e / 1000.0)
    lt = torch.tensor([launches], device=device, dtype=torch.float64)
    mem_t = torch.tensor([torch.cuda.max_memory_allocated(device) / 2**30], device=device, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(lt)
        tdist.all_reduce(mem_t, op=tdist.ReduceOp.MAX)

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): algorithmic 2MNK per launch / CUDA-event duration ----
    roof = None
    busy_local = 0.0
    if prof:
        torch.cuda.synchronize()
        tot_ms = sum(s.elapsed_time(e) for s, e, _, _, _ in prof)
        busy_local = tot_ms / ms_dev if ms_dev else 0.0
        tot_fl = sum(f for _, _, f, _, _ in prof)
        gemm_ms = sum(s.elapsed_time(e) for s, e, _, k, _ in prof if k == 'gemm')
        gemm_fl = sum(f for _, _, f, k, _ in prof if k == 'gemm')
        attn_ms = sum(s.elapsed_time(e) for s, e, _, k, _ in prof if k.startswith('attn'))
        attn_fl = sum(f for _, _, f, k, _ in prof if k.startswith('attn'))
        n_gemm = sum(1 for p in prof if p[3] == 'gemm')
        kind_ms = {}
        for s_, e_, f_, k_, _t in prof:
            kind_ms[k_] = kind_ms.get(k_, 0.0) + s_.elapsed_time(e_)
        kind_share = {k: round(v / (ms_dev if ms_dev else 1), 4) for k, v in sorted(kind_ms.items(), key=lambda kv: -kv[1])}
        by_shape = {}
        for s_, e_, f_, k_, tag in prof:
            if k_ == 'gemm':
                d = by_shape.setdefault(tag, [0, 0.0, 0.0])
                d[0] += 1
                d[1] += s_.elapsed_time(e_)
                d[2] += f_
        breakdown = [{'MNK_aMN_bMN_epi_acc': list(t), 'n': v[0], 'ms': round(v[1], 2), 'tflops': round(v[2] / v[1] / 1e9, 1)}
                     for t, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:14]]
        try:
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(ROOT, 'gpurun_out', f'bench_gemm_breakdown_rank{rank}.json'), 'w') as f:
                json.dump(breakdown, f, indent=1)
        except OSError:
            pass
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        achieved = gemm_fl / gemm_ms / 1e9 if gemm_ms > 0 else 0.0
        roof = {'bound': 'tensor', 'kernel': 'gemm_bf16_kernel (tcgen05, all epilogues/layouts)', 'achieved': achieved,
                'peak': peak, 'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PF sustained',
                'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': gemm_traffic_from_profile(),
                'launches_timed': n_gemm, 'avg_launch_ms': gemm_ms / max(1, n_gemm),
                'share_of_step': gemm_ms / (ms_dev if ms_dev else 1),
                'attention': {'achieved': attn_fl / attn_ms / 1e9 if attn_ms > 0 else None, 'unit': 'TFLOP/s (algorithmic: 4LqLkD fwd, 2.5x bwd)',
                              'share_of_step': attn_ms / (ms_dev if ms_dev else 1)},
                'kernels_share_of_step': tot_ms / (ms_dev if ms_dev else 1), 'share_by_kernel': kind_share,
                'step_tflops': TRAIN_TFLOP_PER_SAMPLE * (n_double + n_single) / 57.0 * value / max(1, world)}

    # kernel-busy fraction of every stage (instrumented kernels / step time): separates pipeline bubbles from imbalance
    busy_t = torch.zeros(world, device=device, dtype=torch.float64)
    busy_t[rank] = busy_local
    if world > 1:
        tdist.all_reduce(busy_t)
    stage_busy = [round(float(x), 4) for x in busy_t.tolist()]

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        del dev_batches, host_batches
        v, desc, cores = cpu_reference_sample(a.res, a.text_len, n_double, n_single)
        cpu = {'value': v, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': desc}

    link_name, schedule_name = type(engine.link).__name__, engine.pipeline_schedule

    if rank == 0:
        h2d = int(h2d_t.item())
        out = {
            'metric': METRIC, 'value': value,
            'unit': 'samples/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms_dev / a.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'Flux-dev full fine-tune bf16 {a.res}x{a.res} (configs[1]/[2] model), {n_double}+{n_single} blocks, '
                                   f'{n_params * (1 if stages == 1 else stages) / 1e9:.1f}B params' + ('' if stages == 1 else ' (approx.)'),
                       'global_batch': samples_per_step, 'micro_batch': mbs, 'micro_batches': M, 'seq_len': 4096 + a.text_len,
                       'parallelism': f'pp{stages}', 'partition': 'flop-balanced manual split' if stages > 1 else 'single stage',
                       'activation_recompute': False, 'train_tflop_per_sample': TRAIN_TFLOP_PER_SAMPLE,
                       'optimizer': 'none (diagnostic)' if a.no_optimizer else 'torch.optim.AdamW(fused) bf16, clip 1.0',
                       'l2_flush': 'working set (>=24 GB of weights+grads per step) is far larger than the 126 MB L2',
                       'stage_link': link_name, 'pipeline_schedule': schedule_name},
            'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / a.steps},
            'gpu_launches': int(lt.item()),
            'loss': loss_dev,
            'peak_mem_gib_max_rank': round(float(mem_t.item()), 1),
            'stage_kernel_busy_frac': stage_busy,
            'clocks': clk,
        }
        if roof:
            out['roofline'] = roof
        if cpu:
            out['cpu_baseline'] = cpu
        # every value of `out` is a plain Python number by now: whatever the context measurement below does, the line prints
        library = None
        if world == 1 and not a.no_library_baseline:
            try:
                del engine, pm, model, layers, params
                import gc
                gc.collect()
                torch.cuda.empty_cache()
                library = gpu_library_sample(a.res, a.text_len, n_double, n_single, device)
            except Exception as exc:      # context only: never lose the bench line over it
                library = {'error': repr(exc)[:300]}
        if library:
            # part of the baseline leg: the same restated reference blocks as `cpu_baseline`, dispatched to the library kernels
            # of this GPU instead of the host cores
            (out['cpu_baseline'] if 'cpu_baseline' in out else out)['gpu_library'] = library
        print(json.dumps(out), flush=True)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == '__main__':
    main()
