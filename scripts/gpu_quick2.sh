#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_pitch_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 900 python -m pytest tests/test_models_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python scripts/bench_conv.py 3xf16 2>&1 | sed -n 8,14p
IRR_CONV_ADDEND=1 IRR_CONV_ONLY=7 timeout 300 python scripts/bench_conv.py 3xf16 2>&1 | tail -1
timeout 120 python scripts/determinism.py 2>&1 | head -2 | cut -c1-160
IRR_DUMP_TIMES=gpurun_out/times_cfg3_q.json timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-gpu --no-pruned --no-strong --cpu-baseline-steps 0 --weights synthetic > gpurun_out/bench_cfg3_q.json 2> gpurun_out/bench_cfg3_q.err; cut -c1-250 gpurun_out/bench_cfg3_q.json; tail -3 gpurun_out/bench_cfg3_q.err
