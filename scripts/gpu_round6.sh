#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests -q -m gpu -s -x > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log
timeout 600 python scripts/bench_conv.py 3xtf32 > gpurun_out/bench_conv.log 2>&1
IRR_CONV_ADDEND=1 timeout 600 python scripts/bench_conv.py 3xtf32 > gpurun_out/bench_conv_addend.log 2>&1
IRR_DUMP_TIMES=gpurun_out/times_3xtf32.json timeout 600 python bench.py --steps 5 --warmup 3 --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/bench_3xtf32.json 2> gpurun_out/bench.err
IRR_CONV_ONLY=7 IRR_CONV_ADDEND=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/conv_tc_32_prof python scripts/bench_conv.py 3xtf32 > gpurun_out/ncu_conv7.log 2>&1
IRR_CONV_ONLY=6 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/conv_tc_thin_prof python scripts/bench_conv.py 3xtf32 > gpurun_out/ncu_conv6.log 2>&1
tail -3 gpurun_out/tests.log; cat gpurun_out/bench_conv.log; cat gpurun_out/bench_conv_addend.log | head -12; python -c "
import json
d=json.load(open('gpurun_out/bench_3xtf32.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline_conv']['achieved'])
"; tail -3 gpurun_out/bench.err
