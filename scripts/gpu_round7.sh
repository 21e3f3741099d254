#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/bench_conv.py 3xtf32 > gpurun_out/bench_conv.log 2>&1
IRR_DUMP_TIMES=gpurun_out/times_3xtf32.json timeout 600 python bench.py --steps 5 --warmup 3 --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/bench_3xtf32.json 2> gpurun_out/bench.err
IRR_NO_SIDE_STREAM=1 timeout 600 python bench.py --steps 5 --warmup 3 --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/bench_3xtf32_noside.json 2>> gpurun_out/bench.err
head -8 gpurun_out/bench_conv.log; python -c "
import json
for f in ['bench_3xtf32','bench_3xtf32_noside']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'])
"; tail -3 gpurun_out/bench.err
