"""CPU: csrc/sched.cpp (the C++ schedule planners and the balanced partitioner — host code, no CUDA) compiled with
AddressSanitizer + UndefinedBehaviorSanitizer and driven over 4000 random geometries (tests/native/sanitizer_driver.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_planners_are_clean_under_asan_and_ubsan(tmp_path):
    exe = str(tmp_path / 'asan_sched')
    cmd = ['g++', '-std=c++17', '-O1', '-g', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined',
           '-I', os.path.join(ROOT, 'include'), '-I', '/usr/local/cuda/include', '-I', os.path.join(ROOT, 'diffusion-pipe_b200', 'csrc'),
           os.path.join(ROOT, 'diffusion-pipe_b200', 'csrc', 'sched.cpp'), os.path.join(ROOT, 'tests', 'native', 'sanitizer_stub.cpp'),
           os.path.join(ROOT, 'tests', 'native', 'sanitizer_driver.cpp'), '-o', exe]
    c = subprocess.run(cmd, capture_output=True, text=True)
    if c.returncode != 0 and ('asan' in c.stderr.lower() or 'ubsan' in c.stderr.lower() or 'sanitize' in c.stderr.lower()):
        pytest.skip('no sanitizer runtime for g++ on this box')
    assert c.returncode == 0, c.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith('ok '), (r.stdout[-500:], r.stderr[-3000:])
