#!/bin/bash
# full GPU test-suite + smoke + both bench arms
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log; tail -4 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_reference.json
