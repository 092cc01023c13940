#!/bin/bash
# ncu evidence for one round (run under gpurun; outputs land in gpurun_out/, summaries are copied to profiles/ by
# tools/summarise_ncu.py on the build box).   usage: tools/profile_ncu.sh r01
set -u
R=${1:-r01}
mkdir -p gpurun_out
# 1) launch list of the bench command (every launch with its device time; shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 42000 -c 4000 --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline \
    > gpurun_out/launches_${R}.log 2>&1
echo "launch list rc=$?"
# 2) full captures of the top kernels on a 1+1 block model at the real shapes
for K in gemm_bf16_kernel attn_fwd_kernel attn_bwd_dkv2_kernel attn_bwd_dq2_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${K} -s 12 -c 4 \
      -o gpurun_out/prof_${K}_${R} -f python bench.py --layers 1,1 --micro-batches 1 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline \
      > gpurun_out/prof_${K}_${R}.log 2>&1
  echo "${K} rc=$?"
done
# 3) the HBM-bound kernels (LayerNorm+modulation, gated residual, q/k-norm+RoPE backward, column reductions, modulation linears,
#    loss): one full capture of each on the same 1+1 block model -> DRAM GB/s against the HBM roof
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'ln_modulate|gate_bwd|qknorm_rope_bwd|colsum|colreduce|mod_fwd|mod_bwd|mse_loss|attn_bwd_delta' -s 30 -c 40 \
    -o gpurun_out/prof_elementwise_${R} -f python bench.py --layers 1,1 --micro-batches 1 --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline \
    > gpurun_out/prof_elementwise_${R}.log 2>&1
echo "elementwise rc=$?"
ls -la gpurun_out/
