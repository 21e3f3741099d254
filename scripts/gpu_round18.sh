#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "3xf16" 2>&1 | tail -8 > gpurun_out/h16_tests.log; cat gpurun_out/h16_tests.log
timeout 300 python scripts/roll_counters.py > gpurun_out/roll_counters3.log 2>&1; head -7 gpurun_out/roll_counters3.log
timeout 300 python scripts/h16_counters.py > gpurun_out/h16_counters4.log 2>&1; cat gpurun_out/h16_counters4.log
IRR_CONV_ADDEND=1 IRR_CONV_ONLY=7 timeout 300 python scripts/bench_conv.py 3xf16 2>&1 | tail -2
timeout 300 python scripts/bench_conv.py 3xf16 2>&1 | tee gpurun_out/h16_bench_conv4.log
