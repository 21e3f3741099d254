#!/bin/bash
mkdir -p gpurun_out
for v in ${CORR_NCU:-plain fused}; do
  a=""; [ $v = fused ] && a=fused
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_tma_kernel -c 1 -f -o gpurun_out/corr_tma_${v}_prof python scripts/corr_one.py 16 32 109 256 $a > gpurun_out/ncu_corr_$v.log 2>&1
done
ls -la gpurun_out/corr_tma*.ncu-rep
