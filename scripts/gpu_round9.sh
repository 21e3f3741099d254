#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests -q -m gpu -s -x > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log
IRR_DUMP_TIMES=gpurun_out/times_3xtf32.json timeout 600 python bench.py --steps 5 --warmup 3 --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/bench_3xtf32.json 2> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_kernel -s 4 -c 1 -o gpurun_out/corr_l4_prof python bench.py --steps 1 --warmup 3 --no-graph --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/ncu_corr.log 2>&1
tail -3 gpurun_out/tests.log; python -c "
import json
d=json.load(open('gpurun_out/bench_3xtf32.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline_conv']['achieved'], d['roofline_conv']['share_of_step'])
for l in d['roofline_corr_levels']: print(l['kernel'], l['B_C_H_W'], round(l['ms'],4), round(l['GBps']))
"; tail -3 gpurun_out/bench.err
