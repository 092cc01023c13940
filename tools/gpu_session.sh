#!/bin/bash
# One gpurun call = one box acquisition: run a whole validation session in it and keep every log under gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh r02 tests bench launches'
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_session.sh r02 scale8'      (charged 8x: keep it short)
# Stages (any subset, in the order given):
#   tests     python -m pytest tests -m gpu            (no -x: one failing case must not hide the rest)
#   poison    the same with DPIPE_TEST_POISON_EMPTY_CUDA=1: uninitialised CUDA memory reads as NaN (tests/conftest.py)
#   bench     bench.py at N=1 (the driver's default line)
#   launches  ncu launch list of the bench command + full captures of the top kernels (tools/profile_ncu.sh)
#   wan       tools/probe_wan_block.py at the 14B shapes + bench.py --family wan on a reduced depth
#   qwen      bench.py --family qwen on a reduced depth
#   scaleN    bench.py on N = 2 / 4 / 8 GPUs, 1F1B and zero-bubble
set -u
R=${1:-r02}; shift
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-900}" "$@" > gpurun_out/${R}_${name}.log 2>&1; echo "== $name rc=$?"; tail -n 3 gpurun_out/${R}_${name}.log; }
trun() { local n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 "$@"; }
for S in "$@"; do
  case $S in
    tests)    T=1500 run tests python -m pytest tests -q -m gpu ;;
    poison)   T=1500 run tests_poison env DPIPE_TEST_POISON_EMPTY_CUDA=1 python -m pytest tests -q -m gpu ;;
    bench)    T=600 run bench1 python bench.py ;;
    launches) T=2400 run ncu bash tools/profile_ncu.sh "$R" ;;
    wan)      T=600 run wan_block python tools/probe_wan_block.py
              T=600 run wan_family python bench.py --family wan --blocks 8 --micro-batches 4 ;;
    qwen)     T=600 run qwen_family python bench.py --family qwen --blocks 12 --micro-batches 8 ;;
    scale2|scale4|scale8)
              N=${S#scale}
              for SCH in 1f1b zb; do
                T=600 run pp${N}_${SCH} python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
                    --master-port 29517 bench.py --gpus "$N" --steps 3 --warmup 3 --schedule $SCH
              done ;;
    *) echo "unknown stage $S" ;;
  esac
done
ls -la gpurun_out | tail -n 30
