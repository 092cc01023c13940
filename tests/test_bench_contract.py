"""CPU: the JSON-line contract of bench.py that can be exercised without a GPU — the `--impl reference` arm (the CPU
restatement of the reference path on the host cores) at a tiny shape, under a plain launch and as a non-zero rank of a
multi-process launch (which must print nothing and exit 0), plus the stage partition the GPU arm uses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ['--impl', 'reference', '--res', '128', '--text-len', '32', '--layers', '1,1', '--steps', '1', '--warmup', '0']


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + ARGS, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'samples/s' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['value'] > 0 and d['gpu_launches'] == 0 and d['data'] == 'synthetic' and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'oracle/flux_ref.py' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    sys.path.insert(0, ROOT)
    import bench
    assert d['metric'] == bench.METRIC                                 # the same metric string as the GPU arm prints


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + ARGS + ['--gpus', '2'], capture_output=True, text=True,
                       timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and not [l for l in p.stdout.splitlines() if l.startswith('{')]


def test_stage_partition_of_the_gpu_arm():
    """57 Flux blocks over N stages: boundaries in layer indices (embedding layer first), extra blocks on the EARLIEST stages"""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flop_balanced_split(19, 38, 1) == ([], [57])
    assert bench.flop_balanced_split(19, 38, 2) == ([30], [29, 28])
    assert bench.flop_balanced_split(19, 38, 4) == ([16, 30, 44], [15, 14, 14, 14])
    split, per = bench.flop_balanced_split(19, 38, 8)
    assert per == [8, 7, 7, 7, 7, 7, 7, 7] and split == [9, 16, 23, 30, 37, 44, 51] and sum(per) == 57
