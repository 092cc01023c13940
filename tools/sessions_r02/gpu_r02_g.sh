#!/bin/bash
# round 2, GPU call G (8 GPUs): BASELINE.json configs[2] (Flux pp8), configs[3] (Wan2.1-14B pp8), configs[4] (Qwen pp2 x dp4)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
clean() { for p in $(nvidia-smi --query-compute-apps=pid --format=csv,noheader | sort -u); do kill -KILL $p 2>/dev/null; done; sleep 2; }
run() {  # name, port, args...
  local name=$1 port=$2; shift 2
  echo "=== $name"
  setsid $TR --master-port $port bench.py --gpus 8 "$@" > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err &
  local pid=$!
  ( sleep 420; kill -KILL -- -$pid 2>/dev/null ) &
  local dog=$!
  wait $pid
  echo "rc=$?"; kill $dog 2>/dev/null; kill -KILL -- -$pid 2>/dev/null grep -v "^stage=\|^    \|^  loss\|OMP_NUM\|\*\*\*\*" gpurun_out/r02g_$name.err | tail -4 | cut -c1-300
  tail -1 gpurun_out/r02g_$name.json | cut -c1-600
  clean
}
run flux_pp8 29601 --steps 5 --warmup 3 --no-cpu-baseline
run wan_pp8 29602 --family wan --steps 3 --warmup 2 --instrumented-steps 1
run qwen_pp2dp4 29603 --family qwen --pp 2 --steps 3 --warmup 2 --instrumented-steps 1
run flux_pp8_1f1b 29604 --steps 3 --warmup 2 --no-cpu-baseline --schedule 1f1b --instrumented-steps 0
nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader | head -8
