#!/bin/bash
# round 2, GPU call D (1 GPU): dq4 kernel (Q / dO in TMEM) correctness + A/B, family smoke runs of bench.py
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
echo "=== attention tests"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_flux_blocks_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python tools/probe_attn.py 2>&1 | grep -v "^{" | tail -3
python - <<'PY'
import json
for r in json.load(open('gpurun_out/probe_attn.json')):
    print(r['case'], 'ok' if r.get('ok') else 'FAIL', {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ('rel', 'ours_tflops', 'sdpa_tflops', 'ok_all', 'exception', 'stderr')})
PY
echo "=== attention backward A/B"
for v in 3 4; do DPIPE_ATTN_BWD=$v timeout 300 python tools/probe_attn.py --case perfbwd:1x24x4608x4608 | grep RESULT | cut -c1-330; done
DPIPE_ATTN_BWD=4 timeout 300 python tools/probe_attn.py --case perfbwd:1x40x9216x9216 | grep RESULT | cut -c1-330
DPIPE_ATTN_BWD=4 timeout 300 python tools/probe_attn.py --case perfbwd:1x40x9216x512 | grep RESULT | cut -c1-330
echo "=== family smoke: wan 2 blocks, qwen 2 blocks (1 GPU)"
timeout 600 python bench.py --family wan --blocks 2 --micro-batches 2 --steps 2 --warmup 1 2>gpurun_out/r02d_wan.err | cut -c1-1500
tail -2 gpurun_out/r02d_wan.err | cut -c1-300
timeout 600 python bench.py --family qwen --blocks 2 --micro-batches 2 --steps 2 --warmup 1 2>gpurun_out/r02d_qwen.err | cut -c1-1500
tail -2 gpurun_out/r02d_qwen.err | cut -c1-300
