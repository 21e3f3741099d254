#!/bin/bash
# first contact with the 3xF16 TMA-staged conv kernel
mkdir -p gpurun_out
timeout 240 python scripts/h16_debug.py > gpurun_out/h16_debug.log 2>&1; echo "debug rc=$?" >> gpurun_out/h16_debug.log
tail -40 gpurun_out/h16_debug.log
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "3xf16" 2>&1 | tail -30 > gpurun_out/h16_tests.log; cat gpurun_out/h16_tests.log
timeout 300 python scripts/bench_conv.py 3xtf32 3xf16 > gpurun_out/h16_bench_conv.log 2>&1; cat gpurun_out/h16_bench_conv.log
