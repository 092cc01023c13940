#!/bin/bash
# round 2, GPU call J (1 GPU): LayerNorm-backward / qk-norm-backward occupancy changes + dK/dV v3 as default candidate
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
exec 2>&1
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_flux_blocks_gpu.py tests/test_flux_e2e_gpu.py tests/test_wan_gpu.py tests/test_qwen_gpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r02j_t4.log
DPIPE_ATTN_BWD=5 timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_flux_blocks_gpu.py tests/test_flux_e2e_gpu.py tests/test_wan_gpu.py tests/test_qwen_gpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r02j_t5.log
timeout 300 python -m pytest "tests/test_real_shapes_gpu.py::test_flux_dev_block_at_1024px[single]" -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r02j_real.log
for v in 4 5; do
  DPIPE_ATTN_BWD=$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-library-baseline > gpurun_out/r02j_bench_v$v.json 2> gpurun_out/r02j_bench_v$v.err
done
echo "=== summary"
cat gpurun_out/r02j_t4.log gpurun_out/r02j_t5.log gpurun_out/r02j_real.log
python - <<'PY'
import json
for v in (4, 5):
    try:
        d = json.loads(open(f'gpurun_out/r02j_bench_v{v}.json').read().strip().splitlines()[-1])
        print(v, round(d['value'], 4), round(d['ms_per_step'], 1), d['clocks']['sm_mhz'], d['first_step_loss'], d['loss'],
              {k: d['roofline']['share_by_kernel'][k] for k in ('attn_bwd', 'ln_bwd', 'qknorm_bwd', 'gemm')})
    except Exception as e:
        print(v, 'ERR', repr(e)[:200])
PY
