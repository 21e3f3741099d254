#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pitch_gpu.py -q -m gpu -k "direct or warp or bf16_store" 2>&1 | tail -8
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log; tail -6 gpurun_out/tests.log
IRR_DUMP_TIMES=gpurun_out/times_cfg3_direct.json timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-gpu --no-pruned --no-strong --cpu-baseline-steps 0 --weights synthetic > gpurun_out/bench_cfg3_direct.json 2> gpurun_out/bench_cfg3_direct.err; cut -c1-250 gpurun_out/bench_cfg3_direct.json; tail -3 gpurun_out/bench_cfg3_direct.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_pitch_gpu.py -q -m gpu -x > gpurun_out/sanitize_memcheck_pitch.log 2>&1; grep -E "passed|failed|ERROR SUMMARY|Invalid|Illegal" gpurun_out/sanitize_memcheck_pitch.log | head -8
