#!/bin/bash
# full bench line + per-kernel time dump
mkdir -p gpurun_out
IRR_DUMP_TIMES=gpurun_out/times.json timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','e2e','clocks','roofline_conv']})"; tail -3 gpurun_out/bench.err
