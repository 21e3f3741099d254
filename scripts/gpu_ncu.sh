#!/bin/bash
# ncu evidence: launch list of one forward + --set full captures of the dominant kernels (summarised into profiles/ off-box)
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/ncu_forward.py > gpurun_out/ncu_forward.log 2>&1
IRR_CONV_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_h16_kernel -s 1 -c 1 -f -o gpurun_out/conv_h16_565_prof python scripts/bench_conv.py 3xf16 > gpurun_out/ncu_conv0.log 2>&1
IRR_CONV_ONLY=5 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_h16_kernel -s 1 -c 1 -f -o gpurun_out/conv_h16_531x32_prof python scripts/bench_conv.py 3xf16 > gpurun_out/ncu_conv5.log 2>&1
IRR_CONV_ONLY=7 IRR_CONV_ADDEND=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 1 -c 1 -f -o gpurun_out/conv_roll_32x32_prof python scripts/bench_conv.py 3xf16 > gpurun_out/ncu_conv7.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:corr_kernel -s 4 -c 1 -f -o gpurun_out/corr_l4_prof python scripts/ncu_forward.py > gpurun_out/ncu_corr.log 2>&1
ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/launches.csv
