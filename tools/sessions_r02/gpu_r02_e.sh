#!/bin/bash
# round 2, GPU call E (2 GPUs): multi-GPU parity tests over NVLink (C++ executor + IpcLink, DistLink), Flux pp2,
# Qwen data-parallel x2 (NCCL gradient all-reduce overlapped with the backward pass) and Qwen pp2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "=== multi-GPU tests"
timeout 1200 python -m pytest tests/test_pipeline_multigpu.py -m gpu -q -x > gpurun_out/r02e_tests.log 2>&1
grep -n -i "error\|Traceback\|assert\|passed\|failed" gpurun_out/r02e_tests.log | head -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== flux pp2"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02e_flux_pp2.json 2> gpurun_out/r02e_flux_pp2.err
tail -2 gpurun_out/r02e_flux_pp2.err | cut -c1-300; cut -c1-2500 gpurun_out/r02e_flux_pp2.json
echo "=== qwen dp2 (20 blocks): overlapped vs serial all-reduce"
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --family qwen --pp 1 --blocks 20 --micro-batches 8 --steps 3 --warmup 2 > gpurun_out/r02e_qwen_dp2.json 2> gpurun_out/r02e_qwen_dp2.err
tail -2 gpurun_out/r02e_qwen_dp2.err | cut -c1-300; cut -c1-2500 gpurun_out/r02e_qwen_dp2.json
DPIPE_DP_OVERLAP=0 timeout 900 $TR --master-port 29513 bench.py --gpus 2 --family qwen --pp 1 --blocks 20 --micro-batches 8 --steps 3 --warmup 2 --instrumented-steps 0 > gpurun_out/r02e_qwen_dp2_serial.json 2> gpurun_out/r02e_qwen_dp2_serial.err
tail -2 gpurun_out/r02e_qwen_dp2_serial.err | cut -c1-300; cut -c1-1200 gpurun_out/r02e_qwen_dp2_serial.json
echo "=== qwen pp2 (40 blocks)"
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --family qwen --pp 2 --blocks 40 --micro-batches 8 --steps 3 --warmup 2 --instrumented-steps 0 > gpurun_out/r02e_qwen_pp2.json 2> gpurun_out/r02e_qwen_pp2.err
tail -2 gpurun_out/r02e_qwen_pp2.err | cut -c1-300; cut -c1-1200 gpurun_out/r02e_qwen_pp2.json
