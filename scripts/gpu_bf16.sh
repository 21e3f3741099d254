#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pitch_gpu.py -q -m gpu -k "bf16" > gpurun_out/bf16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/bf16_tests.log; tail -25 gpurun_out/bf16_tests.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log; tail -8 gpurun_out/tests.log
IRR_DUMP_TIMES=gpurun_out/times_cfg5_bf16.json timeout 600 python bench.py --config 5 --batch 4 --steps 10 --warmup 3 --no-torch-gpu --no-pruned --no-strong --cpu-baseline-steps 0 --weights synthetic > gpurun_out/bench_cfg5_bf16.json 2> gpurun_out/bench_cfg5_bf16.err; cut -c1-250 gpurun_out/bench_cfg5_bf16.json; tail -3 gpurun_out/bench_cfg5_bf16.err
timeout 600 python scripts/two_in_flight.py 2>&1 | tee gpurun_out/two_in_flight.txt | tail -6
