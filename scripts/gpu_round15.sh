#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_models_gpu.py -x -q -m gpu -k "3xf16" -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/h16_model_tests.log; cat gpurun_out/h16_model_tests.log
IRR_DUMP_TIMES=gpurun_out/times_3xf16.json timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_3xf16.json 2> gpurun_out/bench_3xf16.err; cat gpurun_out/bench_3xf16.json; tail -5 gpurun_out/bench_3xf16.err
