#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt
timeout 180 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?" >> gpurun_out/tc_debug.log
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -s -k tcgen05 > gpurun_out/tests_tc.log 2>&1; echo "rc=$?" >> gpurun_out/tests_tc.log
timeout 900 python -m pytest tests -q -m gpu -s --deselect tests/test_ops_gpu.py::test_conv2d_tcgen05_vs_torch_cpu > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log
timeout 600 python scripts/bench_conv.py fp32 3xtf32 tf32 > gpurun_out/bench_conv.log 2>&1
IRR_DUMP_TIMES=gpurun_out/times_3xtf32.json timeout 600 python bench.py --steps 5 --warmup 3 --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/bench_3xtf32.json 2> gpurun_out/bench.err
IRR_CONV_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/conv_tc_565_prof python scripts/bench_conv.py 3xtf32 > gpurun_out/ncu_conv0.log 2>&1
IRR_CONV_ONLY=7 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/conv_tc_32_prof python scripts/bench_conv.py 3xtf32 > gpurun_out/ncu_conv7.log 2>&1
tail -4 gpurun_out/tc_debug.log; tail -3 gpurun_out/tests_tc.log; tail -3 gpurun_out/tests.log; cat gpurun_out/bench_conv.log; cut -c1-300 gpurun_out/bench_3xtf32.json; tail -3 gpurun_out/bench.err
