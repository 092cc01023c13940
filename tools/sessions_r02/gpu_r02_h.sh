#!/bin/bash
# round 2, GPU call H (1 GPU): dK/dV v3 kernel check + A/B, ncu evidence (launch list, full captures, HBM kernels), the whole
# GPU suite, and a clean 1-GPU bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
echo "=== dkv3 (DPIPE_ATTN_BWD=5): tests + A/B"
DPIPE_ATTN_BWD=5 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_flux_blocks_gpu.py -m gpu -q -x 2>&1 | tail -5
T5=$?
for v in 4 5; do DPIPE_ATTN_BWD=$v timeout 300 python tools/probe_attn.py --case perfbwd:1x24x4608x4608 | grep RESULT | cut -c1-130; done
DPIPE_ATTN_BWD=5 timeout 300 python tools/probe_attn.py --case bwd:2x3x512x384 | grep RESULT | cut -c1-400
DPIPE_ATTN_BWD=5 timeout 300 python tools/probe_attn.py --case bwd:1x2x300x200 | grep RESULT | cut -c1-400
echo "=== ncu evidence"
bash tools/profile_ncu.sh r02 2>&1 | tail -12
echo "=== whole GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "=== bench 1 GPU"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench1.json 2> gpurun_out/r02h_bench1.err
tail -2 gpurun_out/r02h_bench1.err | cut -c1-300
tail -1 gpurun_out/r02h_bench1.json | cut -c1-400
