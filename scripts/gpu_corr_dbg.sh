#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/corr_dbg.log; : > $L
run() { echo "== $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run python scripts/corr_one.py 2 32 109 256
run python scripts/corr_one.py 16 32 109 256
run compute-sanitizer --tool memcheck --print-limit 5 python scripts/corr_one.py 16 32 109 256
run python scripts/corr_one.py 2 32 40 64 fused
run compute-sanitizer --tool memcheck --print-limit 5 python scripts/corr_one.py 2 32 40 64 fused
grep -v "^Search\|^CUDA kernel\|^For debugging\|^Compile with\|^Traceback\|^  File\|^    " $L | tail -60
