#!/bin/bash
mkdir -p gpurun_out
{
echo "== default"; timeout 300 python scripts/bench_corr.py 2>&1 | head -3
echo "== IRR_CORR_PF=1"; IRR_CORR_PF=1 timeout 300 python scripts/bench_corr.py 2>&1 | head -3
echo "== NS_FUSED=4"; IRR_B200_LIB=$PWD/build_exp/libirr_b200_ns4.so timeout 300 python scripts/bench_corr.py 2>&1 | head -3
echo "== NS_FUSED=4 + PF"; IRR_CORR_PF=1 IRR_B200_LIB=$PWD/build_exp/libirr_b200_ns4.so timeout 300 python scripts/bench_corr.py 2>&1 | head -3
} | tee gpurun_out/bench_corr_exp2.txt
IRR_CORR_PF=1 IRR_B200_LIB=$PWD/build_exp/libirr_b200_ns4.so timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_pitch_gpu.py -x -q -m gpu -k "cost_volume or correlation" 2>&1 | tail -3
timeout 600 python scripts/two_in_flight.py 2>&1 | tee gpurun_out/two_in_flight.txt | tail -6
