#!/bin/bash
# Round-end evidence run: gpu test-suite, smoke, both bench arms, ncu launch list + full captures (summarised into profiles/ here).
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench.err
IRR_DUMP_TIMES=gpurun_out/times.json timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --math fp32 --cpu-baseline-steps 0 > gpurun_out/bench_fp32.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --math tf32 --cpu-baseline-steps 0 > gpurun_out/bench_tf32.json 2>> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1060 -c 360 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --cpu-baseline-steps 0 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_kernel -s 4 -c 1 -o gpurun_out/corr_l4_prof python bench.py --steps 1 --warmup 3 --no-graph --cpu-baseline-steps 0 > gpurun_out/ncu_corr.log 2>&1
IRR_CONV_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/conv_tc_565_prof python scripts/bench_conv.py 3xtf32 > gpurun_out/ncu_conv0.log 2>&1
timeout 600 python scripts/bench_conv.py fp32 3xtf32 tf32 > gpurun_out/bench_conv.log 2>&1
tail -3 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
