#!/bin/bash
# One GPU call: full gpu test-suite, smoke, a short bench, the ncu launch list and one full capture of the correlation kernel.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s 2>&1 | tail -60 > gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --steps 5 --warmup 3 --no-graph --cpu-baseline-steps 0 > gpurun_out/bench_nograph.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --cpu-baseline-steps 0 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_kernel -s 26 -c 3 -o gpurun_out/corr_prof python bench.py --steps 1 --warmup 3 --no-graph --cpu-baseline-steps 0 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/tests.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
