#!/bin/bash
# round 2, GPU call B (1 GPU): C++ stage executor on the 1-GPU link tests, attention kernels with packed fp32 math.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
echo "=== link / attention / block tests"
timeout 1200 python -m pytest tests/test_stage_link_one_gpu.py tests/test_kernels_gpu.py tests/test_flux_blocks_gpu.py tests/test_flux_e2e_gpu.py tests/test_wan_gpu.py -m gpu -q -x 2>&1 | tail -30
echo "=== attention backward"
for v in 2 3; do DPIPE_ATTN_BWD=$v timeout 300 python tools/probe_attn.py --case perfbwd:1x24x4608x4608 | grep RESULT | cut -c1-400; done
echo "=== attention forward, polynomial share 0 / 2 / 4 of 8"
for p in 0 2 4; do DPIPE_ATTN_FWD_POLY=$p timeout 300 python tools/probe_attn.py --case perf:1x24x4608x4608 | grep RESULT | cut -c1-300; done
DPIPE_ATTN_FWD_POLY=2 timeout 300 python tools/probe_attn.py --case perf:1x40x9216x9216 | grep RESULT | cut -c1-300
timeout 300 python tools/probe_attn.py --case fwd:1x2x1024x1024:peaky | grep RESULT | cut -c1-300
echo "=== bench 1 GPU"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r02b_bench1.json 2> gpurun_out/r02b_bench1.err
tail -3 gpurun_out/r02b_bench1.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d['clocks'])
print(d['roofline']['share_by_kernel'], d['roofline']['attention'], d['roofline']['achieved'])
PY
