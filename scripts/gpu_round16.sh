#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/roll_debug.log; : > $L
run() { echo "== $*" >> $L; timeout 90 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python scripts/h16_one.py 1 32 16 128 32 3 1 1
run python scripts/h16_one.py 2 32 37 256 32 3 1 1
run python scripts/h16_one.py 1 11 50 512 32 3 1 1
grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^Traceback\|^  File\|^    " $L | tail -30
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "3xf16" 2>&1 | tail -15 > gpurun_out/h16_tests.log; cat gpurun_out/h16_tests.log
timeout 300 python scripts/bench_conv.py 3xf16 > gpurun_out/h16_bench_conv3.log 2>&1; cat gpurun_out/h16_bench_conv3.log
