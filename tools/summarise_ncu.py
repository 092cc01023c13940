"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run on the build box, no GPU needed).

    python tools/summarise_ncu.py r01
"""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else 'r01'
OUT = os.path.join(ROOT, 'profiles')
SRC = os.path.join(ROOT, 'gpurun_out')
os.makedirs(OUT, exist_ok=True)


def launch_list():
    path = os.path.join(SRC, f'launches_{R}.csv')
    if not os.path.exists(path):
        return
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rd:
        if len(row) <= vi:
            continue
        try:
            v = float(row[vi].replace(',', ''))
        except ValueError:
            continue
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row[ui], 1.0)
        name = re.sub(r'\(.*', '', re.sub(r'<.*', '', row[ki])).replace('void ', '').strip()[:60]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    with open(os.path.join(OUT, f'{R}_launches_summary.md'), 'w') as f:
        f.write(f'# {R}: ncu launch list of `python bench.py --steps 1 --warmup 1` (Flux-dev 1024^2, 1 GPU)\n\n')
        f.write('`ncu --metrics gpu__time_duration.sum --clock-control none -s 42000 -c 4000` — per-launch times are cold-cache and\n'
                'serialised: compare SHARES with bench.py\'s live CUDA-event shares, not absolutes.\n\n')
        f.write(f'{n} launches, {tot / 1e6:.1f} ms of device time (about 1.9 micro-batches of the steady state).\n\n')
        f.write('| share | total ms | launches | avg us | kernel |\n|---:|---:|---:|---:|---|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v[1] / tot < 5e-4:
                continue
            f.write(f'| {v[1] / tot * 100:.2f}% | {v[1] / 1e6:.2f} | {v[0]} | {v[1] / v[0] / 1e3:.1f} | `{k}` |\n')
    shutil.copy(path, os.path.join(OUT, f'{R}_launches.csv'))


WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__cluster_dim_x', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def full_captures():
    rows_out = []
    for fn in sorted(os.listdir(SRC)):
        if not (fn.startswith('prof_') and fn.endswith(f'_{R}.csv')) or 'elementwise' in fn:
            continue
        with open(os.path.join(SRC, fn)) as fh:
            rd = list(csv.reader(fh.read().splitlines()))
        if len(rd) < 3:
            continue
        hdr, units = rd[0], rd[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for row in rd[2:]:
            d = {'report': fn, 'kernel': row[idx['Kernel Name']][:90]}
            for w in WANT:
                if w in idx:
                    d[w] = f'{row[idx[w]]} {units[idx[w]]}'.strip()
            rows_out.append(d)
    if not rows_out:
        return
    with open(os.path.join(OUT, f'{R}_full_captures.md'), 'w') as f:
        f.write(f'# {R}: `ncu --set full --clock-control none --import-source on` captures (1 double + 1 single block at the real shapes)\n\n')
        for d in rows_out:
            f.write(f"## {d['kernel']}\n\n(report {d['report']}, kept in gpurun_out/ — too large for git)\n\n")
            for w in WANT:
                if w in d:
                    f.write(f'- `{w}`: {d[w]}\n')
            f.write('\n')


def hbm_table():
    """HBM-bound kernels of the elementwise capture: measured DRAM bytes per launch, achieved GB/s against the measured copy
    peak (MEASURED_PEAKS.json hbm_gbs), and the algorithmic bytes of the shape (1 double + 1 single block, L = 4608 = 4096
    image + 512 text tokens, D = 3072, 24 heads) where the kernel name and grid identify it."""
    import json
    fn = os.path.join(SRC, f'prof_elementwise_{R}.csv')
    if not os.path.exists(fn):
        return
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        peak = 6650.0
    with open(fn) as fh:
        rd = list(csv.reader(fh.read().splitlines()))
    if len(rd) < 3:
        return
    hdr, units = rd[0], rd[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(row, name):
        try:
            v = float(row[ix[name]].replace(',', ''))
        except (KeyError, ValueError):
            return None
        u = units[ix[name]]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3,
                 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}.get(u, 1.0)
        return v * scale
    agg = collections.OrderedDict()
    for row in rd[2:]:
        name = re.sub(r'\(.*', '', re.sub(r'<.*', '', row[ix['Kernel Name']])).replace('void ', '').replace('dpipe::', '').strip()
        grid = row[ix['launch__grid_size']] if 'launch__grid_size' in ix else '?'
        t, rb, wb = val(row, 'gpu__time_duration.sum'), val(row, 'dram__bytes_read.sum'), val(row, 'dram__bytes_write.sum')
        if not t or rb is None:
            continue
        d = agg.setdefault((name, grid), [0, 0.0, 0.0, 0.0, 0.0])
        d[0] += 1
        d[1] += t
        d[2] += rb
        d[3] += wb
        pct = val(row, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')
        d[4] += pct or 0.0
    with open(os.path.join(OUT, f'{R}_hbm_kernels.md'), 'w') as f:
        f.write(f'# {R}: the HBM-bound kernels under `ncu --set full --clock-control none` (1 double + 1 single Flux block at 1024^2)\n\n')
        f.write(f'Achieved = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum per launch, against the measured copy\n'
                f'peak of this pool ({peak:.1f} GB/s, MEASURED_PEAKS.json).  Launches of one kernel with the same grid are averaged.  Under ncu\n'
                f'every launch starts with a cold L2 and runs alone; inputs a neighbouring kernel left in the 126 MB L2 show up here as DRAM reads.\n\n')
        # algorithmic bytes (read + written, MB) of the launches that the kernel name and grid identify in the 1 + 1 block Flux
        # model: L = 4608 joint / 4096 image / 512 text rows, D = 3072 (28.3 / 25.2 / 3.1 MB per bf16 [rows, D] matrix)
        algo = {('ln_modulate_bwd_kernel', '288'): (4 * 28.3 + 7.1, 'dxn, x, dres in; dx + column partials out (single block)'),
                ('ln_modulate_bwd_kernel', '256'): (4 * 25.2 + 6.3, 'image stream of the double block'),
                ('ln_modulate_bwd_kernel', '32'): (4 * 3.1 + 0.8, 'text stream'),
                ('qknorm_rope_bwd_kernel', '648'): (5 * 28.3 + 3 * 28.3, 'dq, dk, dv, q-hat, k-hat in; token-major dqkv out (24 heads x 4608)'),
                ('qknorm_rope_bwd_kernel', '576'): (8 * 25.2, 'image stream'),
                ('qknorm_rope_bwd_kernel', '72'): (8 * 3.1, 'text stream'),
                ('gate_bwd_kernel', '256'): (3 * 25.2 + 6.3, 'dx, y in; dy + partials out'),
                ('gate_bwd_kernel', '32'): (3 * 3.1 + 0.8, 'text stream'),
                ('colsum_kernel', '864'): (113.2, 'MLP half of d linear1: [4608, 12288] bf16 in'),
                ('colsum_kernel', '768'): (100.7, '[4096, 12288] bf16 in'),
                ('colsum_kernel', '96'): (12.6, '[512, 12288] bf16 in'),
                ('mod_bwd_kernel', '144'): (56.6 + 56.6, 'W [9216, 3072] in, dW out (first micro-batch: no dW read)'),
                ('mod_bwd_kernel', '288'): (113.2 + 113.2, 'W [18432, 3072] in, dW out'),
                ('attn_bwd_delta_kernel', '13824'): (2 * 28.3, 'O and dO rows in, D out')}
        f.write('| kernel | grid | launches | avg us | DRAM read MB | DRAM write MB | algorithmic MB | GB/s | of HBM peak | ncu dram % |\n|---|---:|---:|---:|---:|---:|---|---:|---:|---:|\n')
        for (name, grid), d in agg.items():
            n, t, rb, wb, pct = d
            gbs = (rb + wb) / t / 1e9
            a = algo.get((name, str(grid)))
            atxt = f'{a[0]:.0f} ({a[1]})' if a else ''
            f.write(f'| `{name}` | {grid} | {n} | {t / n * 1e6:.1f} | {rb / n / 1e6:.1f} | {wb / n / 1e6:.1f} | {atxt} | {gbs:.0f} | {gbs / peak:.2f} | {pct / n:.0f} |\n')
        f.write('\nDRAM writes below the algorithmic output size mean the capture ended with the output still in the 126 MB L2 (write-back '
                'happens under the next kernel); reads match the algorithmic input bytes — none of these kernels re-reads its inputs.\n')


launch_list()
full_captures()
hbm_table()
print(os.listdir(OUT))
