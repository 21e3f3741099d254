#!/bin/bash
# Round-2 evidence run: gpu test-suite (+ parity reports), smoke, both bench arms, configs 4 / 5 on one GPU, micro-benchmarks,
# ncu launch list + full captures (summarised into profiles/ off-box with scripts/summarize_ncu.py).
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt gpurun_out/parity_trained.txt
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/tests.log 2>&1; echo "rc=$?" >> gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench.err
IRR_DUMP_TIMES=gpurun_out/times.json timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2>> gpurun_out/bench.err
timeout 400 python bench.py --config 5 --batch 4 --steps 10 --warmup 3 --no-torch-gpu --no-pruned --no-strong --cpu-baseline-steps 0 > gpurun_out/bench_cfg5_b4.json 2>> gpurun_out/bench.err
timeout 400 python bench.py --config 4 --batch 4 --steps 10 --warmup 3 --no-torch-gpu --no-pruned --no-strong --cpu-baseline-steps 0 > gpurun_out/bench_cfg4_b4.json 2>> gpurun_out/bench.err
timeout 300 python scripts/bench_corr.py > gpurun_out/bench_corr.log 2>&1
timeout 300 python scripts/bench_conv.py 3xf16 > gpurun_out/bench_conv.log 2>&1
timeout 120 python scripts/determinism.py > gpurun_out/determinism.txt 2>&1
if [ "${FINAL_NCU:-1}" = "1" ]; then
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/ncu_forward.py > gpurun_out/ncu_forward.log 2>&1
for v in plain fused; do
  a=""; [ $v = fused ] && a=fused
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_tma_kernel -c 1 -f -o gpurun_out/corr_tma_${v}_prof python scripts/corr_one.py 16 32 109 256 $a > gpurun_out/ncu_corr_$v.log 2>&1
done
IRR_CONV_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_h16_kernel -s 1 -c 1 -f -o gpurun_out/conv_h16_565_prof python scripts/bench_conv.py 3xf16 > gpurun_out/ncu_conv0.log 2>&1
IRR_CONV_ONLY=7 IRR_CONV_ADDEND=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 1 -c 1 -f -o gpurun_out/conv_roll_32x32_prof python scripts/bench_conv.py 3xf16 > gpurun_out/ncu_conv7.log 2>&1
IRR_CONV_ONLY=11 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_direct_kernel -s 1 -c 1 -f -o gpurun_out/conv_direct_3x16_prof python scripts/bench_conv.py fp32 > gpurun_out/ncu_conv11.log 2>&1
fi
tail -3 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_cfg5_b4.json; cut -c1-200 gpurun_out/bench_cfg4_b4.json; tail -3 gpurun_out/bench.err
