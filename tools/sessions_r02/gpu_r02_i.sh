#!/bin/bash
# round 2, GPU call I (1 GPU): ncu evidence (sized to travel back) + dK/dV v2 vs v3 A/B; the summary is printed LAST
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
bash tools/profile_ncu.sh r02 > gpurun_out/r02i_profile.log 2>&1
for v in 4 5; do DPIPE_ATTN_BWD=$v timeout 200 python tools/probe_attn.py --case perfbwd:1x24x4608x4608 | grep RESULT | cut -c1-125 >> gpurun_out/r02i_ab.log; done
DPIPE_ATTN_BWD=5 timeout 200 python tools/probe_attn.py --case bwd:2x3x512x384 | grep RESULT | cut -c1-300 >> gpurun_out/r02i_ab.log
DPIPE_ATTN_BWD=5 timeout 200 python tools/probe_attn.py --case perfbwd:1x40x9216x9216 | grep RESULT | cut -c1-125 >> gpurun_out/r02i_ab.log
echo "=== summary"
tail -9 gpurun_out/r02i_profile.log
cat gpurun_out/r02i_ab.log
ls -la gpurun_out | awk '{print $5, $9}' | tail -20
