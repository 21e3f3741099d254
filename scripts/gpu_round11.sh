#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "cost_volume or warp or corr" > gpurun_out/tests_corr.log 2>&1; echo "rc=$?" >> gpurun_out/tests_corr.log
timeout 600 python -m pytest tests/test_models_gpu.py -q -m gpu -x > gpurun_out/tests_models.log 2>&1; echo "rc=$?" >> gpurun_out/tests_models.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-baseline-steps 0 > gpurun_out/bench_q.json 2> gpurun_out/bench.err
tail -4 gpurun_out/tests_corr.log; tail -3 gpurun_out/tests_models.log; python -c "
import json
d=json.load(open('gpurun_out/bench_q.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
for l in d['roofline_corr_levels']: print(l['kernel'], l['B_C_H_W'], round(l['ms'],4), round(l['GBps']))
"; tail -3 gpurun_out/bench.err
