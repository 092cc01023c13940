#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU; outputs land in gpurun_out/ — kept well under gpurun's 64 MiB return
# limit: reports are exported to CSV on the box and only the attention reports travel back — and are summarised into
# profiles/ by tools/summarise_ncu.py on the build box).   usage: tools/profile_ncu.sh r02
set -u
R=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-library-baseline --instrumented-steps 0"
# 1) launch list of the bench command (every launch with its device time; shares, not absolutes): ~1.5 micro-batches of the
#    steady state of the full 57-block model (-s skips model construction, warm-up and the first micro-batches)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 2600 --csv \
    --log-file gpurun_out/launches_${R}.csv $B --steps 1 --warmup 1 > gpurun_out/launches_${R}.log 2>&1
echo "launch list rc=$?"
# 2) full captures of the tensor-core kernels on a 1+1 block model at the real shapes (3 optimizer steps run: warm-up, timed,
#    end-to-end -> 6 attention launches of each kind, ~100 GEMM launches)
for KS in gemm_bf16_kernel:40:4 attn_fwd_kernel:2:1 attn_bwd_dkv:2:1 attn_bwd_dq:2:1; do
  K=${KS%%:*}; REST=${KS#*:}; SKIP=${REST%%:*}; CNT=${REST#*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K} -s ${SKIP} -c ${CNT} \
      -o gpurun_out/prof_${K}_${R} -f $B --layers 1,1 --micro-batches 1 --steps 1 --warmup 1 > gpurun_out/prof_${K}_${R}.log 2>&1
  echo "${K} rc=$?"
  ncu -i gpurun_out/prof_${K}_${R}.ncu-rep --page raw --csv > gpurun_out/prof_${K}_${R}.csv 2>/dev/null
done
rm -f gpurun_out/prof_gemm_bf16_kernel_${R}.ncu-rep
# 3) the HBM-bound kernels: LayerNorm+modulation, gated residual, q/k-norm+RoPE backward, column reductions, modulation
#    linears, loss, gradient norm / clip: one full capture each -> DRAM bytes and GB/s against the measured copy peak
timeout 600 ncu --set full --clock-control none \
    -k regex:'ln_modulate|gate_bwd|qknorm_rope_bwd|colsum|colreduce|mod_fwd|mod_bwd|mse_loss|attn_bwd_delta|grad_sumsq|grad_scale' -s 20 -c 40 \
    -o gpurun_out/prof_elementwise_${R} -f $B --layers 1,1 --micro-batches 1 --steps 1 --warmup 1 > gpurun_out/prof_elementwise_${R}.log 2>&1
echo "elementwise rc=$?"
ncu -i gpurun_out/prof_elementwise_${R}.ncu-rep --page raw --csv > gpurun_out/prof_elementwise_${R}.csv 2>/dev/null
rm -f gpurun_out/prof_elementwise_${R}.ncu-rep
du -sh gpurun_out
