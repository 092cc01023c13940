#!/bin/bash
# round 2, GPU call C (1 GPU): stage executor tests with full logs; ncu source-level captures of the attention kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
echo "=== link tests"
timeout 900 python -m pytest tests/test_stage_link_one_gpu.py -m gpu -q -x > gpurun_out/r02c_link.log 2>&1
grep -n -i "error\|Traceback\|assert\|raise\|passed\|failed" gpurun_out/r02c_link.log | head -40
echo "=== ncu attention backward"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_d' -s 4 -c 2 -o gpurun_out/prof_attn_bwd_r02c -f \
    python tools/probe_attn.py --case perfbwd:1x24x4608x4608 > gpurun_out/prof_attn_bwd_r02c.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_attn_bwd_r02c.log | cut -c1-200
echo "=== ncu attention forward"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd' -s 4 -c 1 -o gpurun_out/prof_attn_fwd_r02c -f \
    python tools/probe_attn.py --case perf:1x24x4608x4608 > gpurun_out/prof_attn_fwd_r02c.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_attn_fwd_r02c.log | cut -c1-200
ls -la gpurun_out | grep r02c
