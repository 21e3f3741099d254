#!/bin/bash
# GPU call 2: tcgen05 decode/debug, full gpu test-suite, fp32 vs 3xTF32 bench, ncu captures of the level-4 correlation
# launch and of a tcgen05 conv launch.
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt
timeout 180 python scripts/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?" >> gpurun_out/tc_debug.log
timeout 900 python -m pytest tests -q -m gpu -s -x --deselect tests/test_ops_gpu.py::test_conv2d_tcgen05_vs_torch_cpu > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -s -k tcgen05 > gpurun_out/tests_tc.log 2>&1; echo "rc=$?" >> gpurun_out/tests_tc.log
timeout 600 python bench.py --steps 5 --warmup 3 --math fp32 --cpu-baseline-steps 2 > gpurun_out/bench_fp32.json 2> gpurun_out/bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/bench_3xtf32.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --math tf32 --cpu-baseline-steps 0 > gpurun_out/bench_tf32.json 2>> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_kernel -s 4 -c 1 -o gpurun_out/corr_l4_prof python bench.py --steps 1 --warmup 3 --no-graph --math fp32 --cpu-baseline-steps 0 > gpurun_out/ncu_corr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 2 -o gpurun_out/conv_tc_prof python bench.py --steps 1 --warmup 3 --no-graph --math 3xtf32 --cpu-baseline-steps 0 > gpurun_out/ncu_conv.log 2>&1
tail -12 gpurun_out/tc_debug.log; tail -4 gpurun_out/tests.log; tail -4 gpurun_out/tests_tc.log; cut -c1-400 gpurun_out/bench_fp32.json; cut -c1-400 gpurun_out/bench_3xtf32.json; tail -3 gpurun_out/bench.err
