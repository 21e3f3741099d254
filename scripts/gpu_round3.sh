#!/bin/bash
mkdir -p gpurun_out
IRR_DUMP_TIMES=gpurun_out/times_3xtf32.json timeout 600 python bench.py --steps 3 --warmup 3 --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/bench_3xtf32_b.json 2> gpurun_out/bench.err
IRR_DUMP_TIMES=gpurun_out/times_tf32.json timeout 600 python bench.py --steps 3 --warmup 3 --math tf32 --cpu-baseline-steps 0 > gpurun_out/bench_tf32_b.json 2>> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 236 -c 3 -o gpurun_out/conv_tc_big_prof python bench.py --steps 1 --warmup 3 --no-graph --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_conv.log; tail -3 gpurun_out/bench.err
