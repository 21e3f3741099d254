#!/bin/bash
# A/B of programmatic dependent launch (IRR_PDL): parity tests with it on, bench with it on and off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pdl_tests.log; cat gpurun_out/pdl_tests.log
for v in 1 0 1 0; do
  IRR_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --cpu-baseline-steps 0 --no-pruned > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$v.json')); print('IRR_PDL=$v', round(d['value'],2), 'pairs/s', round(d['ms_per_step'],3), 'ms', 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'])" || tail -3 gpurun_out/bench_pdl$v.err
done
