#!/bin/bash
# correlation iteration: parity tests of the correlation / warp ops, micro-benchmark, role counters
mkdir -p gpurun_out
if [ -z "$CORR_SKIP_TESTS" ]; then
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "cost_volume or correlation or warp" 2>&1 | tail -8 > gpurun_out/corr_tests.log; cat gpurun_out/corr_tests.log
timeout 300 python scripts/bench_corr.py > gpurun_out/bench_corr.log 2>&1; cat gpurun_out/bench_corr.log
fi
timeout 120 python scripts/corr_counters.py > gpurun_out/corr_counters.log 2>&1; tail -8 gpurun_out/corr_counters.log
