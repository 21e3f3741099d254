#!/bin/bash
# row-pitch validation: new pitch tests, full gpu suite, cfg 5 (KITTI shape) and cfg 3 bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pitch_gpu.py -q -m gpu > gpurun_out/pitch_tests.log 2>&1; echo "rc=$?" >> gpurun_out/pitch_tests.log; tail -15 gpurun_out/pitch_tests.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log; tail -12 gpurun_out/tests.log
IRR_DUMP_TIMES=gpurun_out/times_cfg5.json timeout 600 python bench.py --config 5 --batch 4 --steps 10 --warmup 3 --no-torch-gpu --no-pruned --no-strong --cpu-baseline-steps 0 > gpurun_out/bench_cfg5_pitch.json 2> gpurun_out/bench_cfg5_pitch.err; cut -c1-300 gpurun_out/bench_cfg5_pitch.json; tail -3 gpurun_out/bench_cfg5_pitch.err
IRR_DUMP_TIMES=gpurun_out/times_cfg3.json timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-gpu --no-pruned --no-strong --cpu-baseline-steps 0 > gpurun_out/bench_cfg3_pitch.json 2> gpurun_out/bench_cfg3_pitch.err; cut -c1-300 gpurun_out/bench_cfg3_pitch.json; tail -3 gpurun_out/bench_cfg3_pitch.err
