#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "3xf16" 2>&1 | tail -12 > gpurun_out/h16_tests.log; cat gpurun_out/h16_tests.log
timeout 1200 python -m pytest tests/test_models_gpu.py -x -q -m gpu -k "3xf16" 2>&1 | tail -12
bash scripts/gpu_bench.sh
