#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/h16_debug2.log; : > $L
run() { echo "== $*" >> $L; timeout 90 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python scripts/h16_one.py 1 32 16 32 32 3 1 1
run python scripts/h16_one.py 1 32 16 256 32 3 1 1
run python scripts/h16_one.py 1 32 16 32 32 3 1 2
run python scripts/h16_one.py 2 115 28 64 128 3 1 1
run python scripts/h16_one.py 1 64 109 256 64 3 1 1
grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^Traceback\|^  File\|^    " $L | tail -40
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "3xf16" 2>&1 | tail -30 > gpurun_out/h16_tests.log; cat gpurun_out/h16_tests.log
timeout 300 python scripts/bench_conv.py 3xtf32 3xf16 > gpurun_out/h16_bench_conv.log 2>&1; cat gpurun_out/h16_bench_conv.log
