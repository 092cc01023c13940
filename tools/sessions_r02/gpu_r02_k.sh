#!/bin/bash
# round 2, GPU call K (2 GPUs): Flux pp2 bench line on a clean box (the earlier 2-GPU numbers were polluted)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
setsid python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02k_flux_pp2.json 2> gpurun_out/r02k_flux_pp2.err &
pid=$!
( sleep 200; kill -KILL -- -$pid 2>/dev/null ) &
dog=$!
wait $pid; echo "rc=$?"; kill $dog 2>/dev/null
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02k_flux_pp2.json').read().strip().splitlines()[-1])
print(round(d['value'], 4), round(d['ms_per_step'], 1), round(d['e2e']['value'], 4), d['first_step_loss'], d['loss'], d['stage_kernel_busy_frac'], d['clocks']['sm_mhz'],
      d['config']['partition']['blocks_per_stage'], d['config']['partition'].get('double_over_single'), round(d['roofline']['achieved']))
PY
