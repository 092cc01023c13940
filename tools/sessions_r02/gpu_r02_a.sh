#!/bin/bash
# round 2, GPU call A (1 GPU): new parity tests first, then the whole GPU suite, attention-backward A/B, a short bench.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
echo "=== new tests"
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_stage_link_one_gpu.py tests/test_eval_checkpoint_gpu.py tests/test_real_shapes_gpu.py -m gpu -q -x 2>&1 | tail -40
echo "=== rest of the gpu suite"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_real_shapes_gpu.py --deselect tests/test_stage_link_one_gpu.py --deselect tests/test_eval_checkpoint_gpu.py --deselect tests/test_kernels_gpu.py 2>&1 | tail -15
echo "=== attention backward A/B"
for v in 2 3; do DPIPE_ATTN_BWD=$v timeout 300 python tools/probe_attn.py --case perfbwd:1x24x4608x4608 | grep RESULT; done
DPIPE_ATTN_BWD=3 timeout 300 python tools/probe_attn.py --case perfbwd:1x40x9216x9216 | grep RESULT
timeout 300 python tools/probe_attn.py --case perf:1x24x4608x4608 | grep RESULT
echo "=== bench 1 GPU"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r02a_bench1.json 2> gpurun_out/r02a_bench1.err
tail -3 gpurun_out/r02a_bench1.err
cat gpurun_out/r02a_bench1.json
