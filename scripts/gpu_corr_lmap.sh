#!/bin/bash
mkdir -p gpurun_out
./scripts/probe/corr_loop_probe > gpurun_out/corr_loop_probe.txt 2>&1; cat gpurun_out/corr_loop_probe.txt
echo "== default"; timeout 300 python scripts/bench_corr.py 2>&1 | head -3 | tee gpurun_out/bench_corr_lmap0.txt
echo "== LMAP"; IRR_CORR_LMAP=1 timeout 300 python scripts/bench_corr.py 2>&1 | head -3 | tee gpurun_out/bench_corr_lmap1.txt
IRR_CORR_LMAP=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_pitch_gpu.py -x -q -m gpu -k "cost_volume or correlation" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_losses.py tests/test_models_gpu.py -q -m gpu -k "harness or pipelined or family" 2>&1 | tail -6
